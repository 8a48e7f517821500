"""tools/sass_diff.py OLD.so NEW.so -- which kernels of two builds of libwcfe.so differ, instruction for instruction (cuobjdump -sass,
addresses and encodings stripped).  Used to show that a later commit left the measured kernels untouched:

    git worktree add /tmp/oldtree <measured commit>
    (cd /tmp/oldtree/wholecellfe_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -o /tmp/libwcfe_old.so *.cu)
    python tools/sass_diff.py /tmp/libwcfe_old.so wholecellfe_b200/libwcfe.so
"""
import subprocess, re, hashlib, sys
def sass(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    funcs = {}; cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m: cur = m.group(1); funcs[cur] = []; continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            # strip the address / encoding comments, keep the instruction text
            ins = re.sub(r"/\*[0-9a-fx]+\*/", "", line).strip()
            funcs[cur].append(ins)
    return {k: hashlib.sha1("\n".join(v).encode()).hexdigest()[:12] + ":%d" % len(v) for k, v in funcs.items()}
a = sass(sys.argv[1]); b = sass(sys.argv[2])
same = sorted(k for k in a if k in b and a[k] == b[k]); diff = sorted(k for k in a if k in b and a[k] != b[k])
print("identical:", len(same), "changed:", len(diff), "only old:", sorted(set(a) - set(b)), "only new:", sorted(set(b) - set(a)))
import subprocess as sp
for k in diff: print("CHANGED", sp.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:150], a[k], b[k])
