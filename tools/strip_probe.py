"""Diagnostic: one strip of the partitioned bench workload as a plain single-GPU problem (no halo exchange), per-kernel times.
Separates "what does a strip cost" from "what does the exchange cost".  python tools/strip_probe.py [world] [halo]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wholecellfe_b200 as W
from wholecellfe_b200.meshgen import disk_mesh, square_mesh
from wholecellfe_b200.partition import strip_partition, connect_plan

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
halo = len(sys.argv) > 2 and sys.argv[2] == "halo"
xc, tc = disk_mesh(1000000); xs, ts = square_mesh(2000)
parts = strip_partition(xs, ts, world)
only0 = len(sys.argv) > 3 and sys.argv[3] == "first"
for r in ([0, world // 2] if world > 2 and not only0 else [0]):
    part = parts[r]
    g = W.ULSim("k31"); g.upload_mesh(W.CELL, xc, tc); g.upload_mesh(W.SUBS, xs[part.local_nodes], part.tri_local); g.set_rng(1, 1)
    if halo:      # connect the strip to ITSELF as every neighbour: pushes and counters exercised, rows land in its own ghost rows
        info = g.halo_info()
        nb, slot, rows, en, ep, er, gh = connect_plan(parts, r)
        er = np.minimum(er, len(part.local_nodes) - 1)
        g.halo_connect([info] * len(nb), list(range(len(nb))), [int((ep == k).sum()) for k in range(len(nb))], en, ep, er, gh)
    t = g.step_n(20, 2e-5, 0.0)
    t = g.step_n(300, 2e-5, t)
    ms = g.last_step_ms() / 300
    g.set_profile(True); t = g.step_n(10, 2e-5, t); kms, kcnt = g.kernel_ms(); g.set_profile(False)
    print("world", world, "strip", r, "halo" if halo else "plain", "local nodes", len(part.local_nodes), "ms/step %.4f" % ms,
          "per kernel", {i: round(kms[i] / kcnt[i], 4) for i in range(6) if kcnt[i]}, g.path_info(), flush=True)
    g.close()
