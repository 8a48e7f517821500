"""Build wholecellfe_b200/libwcfe.so (hand-written sm_100a CUDA + the C ABI of include/wcfe.h) in-tree with nvcc."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwcfe.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared"]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "wcfe.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: tuning variants (e.g. defines=("WCFE_PATCH=128",), out="libwcfe_p128.so"), loaded via env WCFE_LIB."""
    target = os.path.join(HERE, out) if out else LIB
    if not force and not out and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", target] + sources()
    subprocess.check_call(cmd, cwd=CSRC)
    return target


if __name__ == "__main__":
    print(build(force=True, verbose=True))
