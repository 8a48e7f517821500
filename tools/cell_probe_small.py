"""one context, partitioned-cell instantiation with no peers and every 8th patch: a single-wave launch of k_cell_patch for ncu"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wholecellfe_b200 as W
from wholecellfe_b200.meshgen import disk_mesh, square_mesh
xc, tc = disk_mesh(1000000); xs, ts = square_mesh(2000)
npatch = (len(xc) + 127) // 128
cx = np.minimum.reduceat(xc[:, 0], np.arange(0, len(xc), 128))
sel = np.nonzero((cx > -2) & (cx < 1))[0]
g = W.ULSim("k31"); g.upload_mesh(W.CELL, xc, tc); g.upload_mesh(W.SUBS, xs, ts); g.set_rng(1, 1)
g.cell_halo_info(); g.cell_halo_connect([], [], sel.astype(np.int32), np.ones(len(sel), np.int32), [], [], [], 5.0)
t = g.step_n(30, 2e-5, 0.0)
print("patches", len(sel), "ms/step", g.last_step_ms() / 30)
g.close()
