#!/usr/bin/env python
"""Workload for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every production instantiation of the step once,
on meshes small enough for the sanitizer's ~50x slow-down.

    compute-sanitizer --tool racecheck python tools/sanitize_run.py [pdl0]

  1  the reference's shipped example (config 1), k1 / k31 / kMD, glibc replay tape: k_cell_patch<PREDICT>, k_boundary_finish,
     k_bucket_reduce<WARP> (coarse substrate), k_subs_ring, PDL chain (or fork/join with `pdl0`)
  2  the adapter sequence on the same meshes: predict / commit / update (k_cell_patch<!PREDICT>), round trip through the host
  3  a 20k-triangle kMD cell on a finer substrate, Philox: k_bucket_reduce<thread per triangle>
     (1 and 3: the cell is displaced by most of a substrate triangle once, so that the persistent anchor membership is exercised:
     entries die, buckets are compacted, triangles listed)
  4  two substrate strips in one process (HALO instantiation of k_subs_ring, peer stores + in-kernel waits, cell counter)
  5  a batch of 4 replicas (BATCH instantiations)
  6  a non-planar state (general kernels k_predict_cell, k_bond, k_cell_node, k_subs_node)
  7  three strips in one process with the CELL partitioned too (PART instantiations of k_cell_patch / k_finish_reduce, pushes into
     the other contexts' buffers, epoch-counter waits), strips and sectors
  8  a moving substrate (k_predict_subs, displacement interpolation in k_bond, grid rebuilds forced by a 4-nm padding)
Nothing is compared here -- tests/ does that; the sanitizer's report is the result."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

if "pdl0" in sys.argv:
    os.environ["WCFE_PDL"] = "0"
import wholecellfe_b200 as W  # noqa: E402
from wholecellfe_b200.meshgen import disk_mesh, square_mesh  # noqa: E402
from wholecellfe_b200.partition import InProcessStrips  # noqa: E402
from wholecellfe_b200.ensemble import sweep_grid, EnsembleRunner  # noqa: E402

d = np.load(os.path.join(ROOT, "tests", "golden", "example_meshes.npz"))
xc, tc, xs, ts = d["cell_xyz"], d["cell_tri"], d["subs_xyz"], d["subs_tri"]
n = 6
for var in ("k1", "k31", "kMD"):
    g = W.ULSim(var); g.upload_mesh(W.CELL, xc, tc); g.upload_mesh(W.SUBS, xs, ts); g.srand(1)
    t = g.step_n(n, 0.005, 0.0)
    for _ in range(3):
        t = g.step(0.005, t)                      # predict / commit / update
    t, _, _ = g.step_n_roundtrip(3, 0.005, t)
    x = g.get(W.CELL, "x"); x[:, 0] += 3.7; x[:, 1] -= 2.9          # every anchor leaves its triangle: dead entries, compaction, listing
    g.set(W.CELL, "x", x)
    t = g.step_n(3, 0.005, t)
    g.get(W.CELL, "x"); g.close()
print("1-2 done", flush=True)
xc3, tc3 = disk_mesh(20000); xs3, ts3 = square_mesh(160)
g = W.ULSim("kMD"); g.upload_mesh(W.CELL, xc3, tc3); g.upload_mesh(W.SUBS, xs3, ts3); g.set_rng(1, 1)
t = g.step_n(n, 5e-5, 0.0)
x = g.get(W.CELL, "x"); x[:, 0] += 0.41; x[:, 1] += 0.33            # the same with one thread per bucket
g.set(W.CELL, "x", x)
g.step_n(3, 5e-5, t); g.get(W.SUBS, "ea"); g.close()
print("3 done", flush=True)
run = InProcessStrips("k31", (xc3, tc3), (xs3, ts3), [0, 0])
run.step_n(n, 5e-5)
run.close()
print("4 done", flush=True)
ens = EnsembleRunner(sweep_grid(4), (xc, tc), (xs, ts))
ens.advance(n, 0.005); ens.close()
print("5 done", flush=True)
g = W.ULSim("k1"); g.upload_mesh(W.CELL, xc, tc); g.upload_mesh(W.SUBS, xs, ts); g.srand(1)
ea = np.zeros((len(xc), 6)); ea[:, 2] = 1e-3
g.set(W.CELL, "ea", ea)
g.step_n(n, 0.005, 0.0); g.close()
print("6 done", flush=True)
for shape in ("strips", "sectors"):
    run = InProcessStrips("k31", (xc3, tc3), (xs3, ts3), [0, 0, 0], cell="partitioned", shape=shape)
    run.step_n(n, 5e-5)
    run.close()
print("7 done", flush=True)
os.environ["WCFE_GRID_PAD"] = "0.004"
p = W.default_params("k1"); p.subs_scale = 1.0
g = W.ULSim("k1", params=p); g.upload_mesh(W.CELL, xc, tc); g.upload_mesh(W.SUBS, xs, ts); g.srand(1)
V = np.zeros((len(xs), 3)); V[:, 0] = 2.0 * xs[:, 1] / 50; V[:, 1] = 1.2 * xs[:, 0] / 50
g.set(W.SUBS, "V", V)
g.step_n(12, 1e-3, 0.0)
print("8 done, grid rebuilt", g.grid_refreshes(), "times", flush=True); g.close()
