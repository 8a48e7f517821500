#!/usr/bin/env python
"""Turn one tools/final_check.sh run (gpurun_out/r02z_*) into the tracked files under profiles/: launch summary, ncu summary with source
hot spots, traffic.json, the two bench lines."""
import collections
import csv
import json
import re
import shutil
import subprocess
import sys

TAG = sys.argv[1] if len(sys.argv) > 1 else "r02z"
ROUND = "r02"
rows = list(csv.reader(open("gpurun_out/%s_launches.csv" % TAG, errors="ignore")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
head = rows[hi]; col = {h: i for i, h in enumerate(head)}
acc = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(head) or r[col["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[col["Kernel Name"]].replace("wcfe::", "")).replace("(bool)", "")
    v = float(r[col["Metric Value"]].replace(",", "")); u = r[col["Metric Unit"]]
    acc.setdefault(name, []).append(v / 1000 if u in ("nsecond", "ns") else (v if u in ("usecond", "us") else v * 1000))
chain = ["void k_cell_patch<1, 0, 0, 0>", "void k_finish_reduce<0, 0, 0>", "void k_subs_ring<0, 0, 0>"]
tot = sum(sum(acc[k]) / len(acc[k]) for k in chain if k in acc)
with open("profiles/%s_launches_summary.txt" % ROUND, "w") as out:
    out.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-other-configs\n"
              "# round 2, final tree (tools/final_check.sh + tools/refresh_profiles.py); the first 400 launches of the process, serialised under ncu: no PDL overlap.  The timed step is the chain\n"
              "# k_cell_patch<1,...> -> k_finish_reduce -> k_subs_ring (shares below are of that chain); k_boundary_finish / k_bucket_reduce as separate launches belong to\n"
              "# bench.py's per-kernel profile pass, k_predict_cell / k_cell_patch<0,...> / k_aos_to_soa / k_soa_to_aos to its host round-trip (e2e) passes\n")
    for k, v in acc.items():
        a = sum(v) / len(v)
        out.write("%-60s launches %4d  avg %9.1f us%s\n" % (k, len(v), a, "  share of the chain %.3f" % (a / tot) if k in chain else ""))
shutil.copy("gpurun_out/%s_launches.csv" % TAG, "profiles/%s_launches_bench_steps30.csv" % ROUND)
rep = "gpurun_out/%s_full.ncu-rep" % TAG
run = lambda *a: subprocess.run([sys.executable, *a], capture_output=True, text=True).stdout
with open("profiles/%s_ncu_summary.txt" % ROUND, "w") as out:
    out.write('# ncu --set full --clock-control none --import-source on -k regex:"k_cell_patch|k_finish_reduce|k_subs_ring" -s 12 -c 3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs\n'
              "# round 2 (B200, bench workload: 998784-triangle cell on 4M-node substrate, k31), final tree of round 2 (tools/final_check.sh); per-launch values of one step; times under ncu are serialised (compare shares)\n")
    out.write(run("profiles/summarize_ncu.py", rep, "--traffic", "profiles/traffic.json"))
    for kern, n in (("k_cell_patch", 30), ("k_subs_ring", 24), ("k_finish_reduce", 14)):
        out.write("\n## where %s spends its warp-time and its instructions (profiles/source_hotspots.py, same capture)\n" % kern)
        out.write("\n".join(run("profiles/source_hotspots.py", rep, kern).splitlines()[:n]) + "\n")
for src, dst in (("bench_n1", "bench_n1"), ("bench_ref", "bench_reference_arm")):
    d = json.loads(open("gpurun_out/%s_%s.json" % (TAG, src)).read().strip().splitlines()[-1])
    json.dump(d, open("profiles/%s_%s.json" % (ROUND, dst), "w"), indent=1)
    print(dst, d["value"], d.get("ms_per_step"), d["e2e"]["value"])
