"""Diagnostic: how does k_cell_patch's duration depend on the number of patches a rank computes?  One context, the partitioned-cell
instantiation with NO peers and a patch list of the first / a strided / a geometric fraction of the patches (physics wrong: halo
rows go stale -- timing only).  python tools/cell_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wholecellfe_b200 as W
from wholecellfe_b200.meshgen import disk_mesh, square_mesh

xc, tc = disk_mesh(1000000); xs, ts = square_mesh(2000)
npatch = (len(xc) + 127) // 128
cx = np.minimum.reduceat(xc[:, 0], np.arange(0, len(xc), 128))
for name, sel in (("all", np.arange(npatch)), ("x<0 half", np.nonzero(cx < 0.05)[0]), ("x<-5 (1/5)", np.nonzero(cx < -5)[0]), ("strip -2..1 (1/8 of a run of 8)", np.nonzero((cx > -2) & (cx < 1))[0]),
                  ("every 2nd", np.arange(0, npatch, 2)), ("every 8th", np.arange(0, npatch, 8))):
    g = W.ULSim("k31"); g.upload_mesh(W.CELL, xc, tc); g.upload_mesh(W.SUBS, xs, ts); g.set_rng(1, 1)
    if name != "all":
        g.cell_halo_info()
        g.cell_halo_connect([], [], sel.astype(np.int32), np.ones(len(sel), np.int32), [], [], [], 5.0)
    t = g.step_n(10, 2e-5, 0.0)
    try:
        t = g.step_n(100, 2e-5, t)
        ms = g.last_step_ms() / 100
        g.set_profile(True); t = g.step_n(10, 2e-5, t); kms, kcnt = g.kernel_ms(); g.set_profile(False)
        print("%-34s patches %5d  ms/step %.4f  per kernel %s" % (name, len(sel), ms, {i: round(kms[i] / kcnt[i], 4) for i in range(6) if kcnt[i]}), flush=True)
    except Exception as e:
        print(name, "failed:", e, flush=True)
    g.close()
