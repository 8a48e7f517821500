"""Diagnostic: in-process strips (fused peer halo) against the single-GPU run; reports where the two differ."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wholecellfe_b200 as W
from wholecellfe_b200.meshgen import disk_mesh, square_mesh
from wholecellfe_b200.partition import InProcessStrips

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
xc, tc = disk_mesh(6000); xs, ts = square_mesh(120)
rng = np.random.default_rng(11); ea0 = np.zeros((len(xs), 6)); ea0[:, [0, 1, 3]] = 1e-3 * rng.normal(size=(len(xs), 3))
run = InProcessStrips("k31", (xc, tc), (xs, ts), [0] * world, rng=(1, 5)); run.set_subs_field("ea", ea0)
run.step_n(steps, 2e-4)
g = W.ULSim("k31"); g.upload_mesh(W.CELL, xc, tc); g.upload_mesh(W.SUBS, xs, ts); g.set_rng(1, 5); g.set(W.SUBS, "ea", ea0)
g.step_n(steps, 2e-4, 0.0)
print("path", g.path_info())
for f in ("ea", "V", "A"):
    ref = g.get(W.SUBS, f); got = np.full_like(ref, np.nan)
    for r in range(world):
        ids, vals = run.owned_field(r, f); got[ids] = vals
    bad = np.nonzero((got != ref).any(axis=1))[0]
    print(f, "mismatching nodes", len(bad), "max abs diff", float(np.abs(got - ref).max()), "scale", float(np.abs(ref).max()))
    if len(bad):
        print("   x range of bad nodes", xs[bad, 0].min(), xs[bad, 0].max(), " y range", xs[bad, 1].min(), xs[bad, 1].max(), " first ids", bad[:8])
for r, sim in enumerate(run.sims):
    print("cell x equal on strip", r, bool(np.array_equal(sim.get(W.CELL, "x"), g.get(W.CELL, "x"))))
