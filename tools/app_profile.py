#!/usr/bin/env python
"""Where the whole reference application's step goes when its UL calls run through ul_adapter.cpp + libwcfe.so:
oracle/_ref/cellBP_gpu_k1 on the shipped example (El Topo on, two .vtk files per step) with WCFE_ADAPTER_PROFILE=1, next to the
same run of oracle/_ref/cellBP_ref_k1.  Prints wall time per step of both and the adapter's own per-entry-point breakdown."""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from wholecellfe_b200.meshio import write_ecomsol  # noqa: E402

d = np.load(os.path.join(ROOT, "tests", "golden", "example_meshes.npz"))
xc, tc, xs, ts = d["cell_xyz"], d["cell_tri"], d["subs_xyz"], d["subs_tri"]
tpl = open(os.path.join(ROOT, "tests", "golden", "reference_script_template.txt")).read()
end_t = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
with tempfile.TemporaryDirectory() as w:
    os.makedirs(os.path.join(w, "scripts"))
    write_ecomsol(os.path.join(w, "scripts", "disk.txt"), xc, tc); write_ecomsol(os.path.join(w, "scripts", "subs.txt"), xs, ts)
    open(os.path.join(w, "cell.txt"), "w").write(tpl.format(end_t=end_t, max_edge=2, improve=1, topo=1, collide=1, mesh="scripts/disk.txt"))
    open(os.path.join(w, "subs.txt"), "w").write(tpl.format(end_t=end_t, max_edge=8, improve=1, topo=1, collide=1, mesh="scripts/subs.txt"))
    for k in ("ref", "gpu"):
        exe = os.path.join(ROOT, "oracle", "_ref", "cellBP_%s_k1" % k)
        out = os.path.join(w, "out_" + k); os.makedirs(out)
        env = dict(os.environ, WCFE_ADAPTER_PROFILE="1")
        t0 = time.perf_counter()
        r = subprocess.run([exe, "cell.txt", "subs.txt", out], cwd=w, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env, timeout=1200)
        wall = time.perf_counter() - t0
        steps = round(end_t / 0.005)
        print("%s: %.2f s wall, %d steps, %.3f ms per step (process start-up included)" % (k, wall, steps, 1e3 * wall / steps))
        print("\n".join(l for l in r.stderr.splitlines() if "adapter" in l or l.startswith("  ")))
