"""Larger meshes than the oracle can follow for long: a refined kMD case against the grid-accelerated oracle (BASELINE
configs[2] in miniature) and size-independent properties at the full bench size (configs[3]: 1M-triangle cell on a
4M-node substrate).  GPU only."""
import numpy as np
import pytest
from conftest import rel_err
import wholecellfe_b200 as W
from wholecellfe_b200.meshgen import disk_mesh, square_mesh
from oracle.oracle import OracleSim

pytestmark = pytest.mark.gpu
FLOOR = {"ea": 1e-3, "fa_u": 1e-4, "V": 1e-3, "A": 1.0, "cC": 1e-3, "x": 1.0}


def test_refined_kmd_mesh_step_by_step_against_grid_oracle():
    xc, tc = disk_mesh(20000)
    xs, ts = square_mesh(161)
    g = W.ULSim("kMD"); g.upload_mesh(W.CELL, xc, tc); g.upload_mesh(W.SUBS, xs, ts); g.srand(1)
    o = OracleSim("kMD", xc, tc, xs, ts, use_grid=True); o.srand(1)
    assert g.path_info()["planar"]
    dt = 2e-4; t = 0.0
    for k in range(6):
        for f in ["x", "V", "A", "ea", "cC", "fa_u"]:
            g.set(W.CELL, f, o.get(0, f))
        for f in ["V", "A", "ea"]:
            g.set(W.SUBS, f, o.get(1, f))
        assert g.path_info()["planar"]             # the oracle's state (1e-16 noise in e_zz) stays on the production kernels
        g.step(dt, t); o.step(dt, t); t += dt
        for f in ["x", "V", "A", "ea", "cC", "fa_u"]:
            ref = o.get(0, f)
            assert rel_err(g.get(W.CELL, f), ref, max(np.abs(ref).max(), FLOOR[f])) < 1e-10, (k, f)
        for w in ["loc_force_cnt", "loc_force_first", "ap_fire", "fa_reset", "cc_clamp"]:
            assert np.array_equal(g.discrete(w), o.discrete(w)), (k, w)


@pytest.fixture(scope="module")
def big():
    xc, tc = disk_mesh(1000000)
    xs, ts = square_mesh(2000)
    return xc, tc, xs, ts


def test_full_size_properties(big):
    xc, tc, xs, ts = big
    dt = 2e-5
    outs = []
    for rep in range(3):
        g = W.ULSim("k31"); g.upload_mesh(W.CELL, xc, tc); g.upload_mesh(W.SUBS, xs, ts); g.srand(1)
        assert g.path_info()["planar"]
        if rep == 2:
            g.set_debug(True)                      # force-field capture (a differently compiled kernel instance)
        g.step_n(20, dt, 0.0)
        outs.append([g.get(W.CELL, f) for f in ["x", "V", "A", "ea", "cC", "fa_u"]] + [g.get(W.SUBS, f) for f in ["V", "A", "ea"]])
        if rep == 2:
            # every anchor is found, and what the integrins pull on the cell they push on the substrate (UL:1259 vs
            # UL:1414-1422): the two force fields cancel -- counting an anchor that lies exactly on a shared substrate edge
            # once per containing triangle, as the reference's exhaustive loop does (no `break`, UL:1340-1451)
            cnt = g.discrete("loc_force_cnt")
            assert cnt.min() >= 1 and cnt.max() <= 6
            Fcell = g.force_field("F_FA")["F"]
            Fc = (Fcell * cnt[:, None]).sum(axis=0); Fs = g.force_field("F_FA_subs")["F"].sum(axis=0)
            scale = np.abs(Fcell).sum()
            assert scale > 0 and np.abs(Fc + Fs).max() < 1e-10 * scale
            # lumped masses add up to the area (thick = rho = 1) and the total area is that of the disk
            d = g.diagnostics()
            assert abs(g.force_field("Res_cell")["MM"].sum() - d[0]) < 1e-9 * d[0]
            assert abs(d[0] - np.pi * 100.0) < 0.2
            assert abs(g.force_field("Res_subs")["MM"].sum() - 1e4) < 1e-6
            assert np.isfinite(outs[-1][0]).all()
        g.close()
    for u, v in zip(outs[0], outs[1]):
        assert np.array_equal(u, v)                # bit-reproducible run to run at full size
    for u, v in zip(outs[0][:6], outs[2][:6]):
        assert rel_err(u, v, max(np.abs(v).max(), 1e-3)) < 1e-9    # debug capture does not change the physics


def test_full_size_planar_path_equals_general_path(big):
    xc, tc, xs, ts = big
    res = []
    for general in (False, True):
        g = W.ULSim("k31"); g.upload_mesh(W.CELL, xc, tc); g.upload_mesh(W.SUBS, xs, ts); g.set_rng(1, 7)   # Philox
        g.set_force_general(general)
        g.step_n(5, 2e-5, 0.0)
        res.append({f: g.get(W.CELL, f) for f in ["x", "V", "A", "ea", "cC", "fa_u"]})
        res[-1]["loc"] = g.discrete("loc_force_first")
        g.close()
    for f in ["x", "V", "A", "ea", "cC", "fa_u"]:
        assert rel_err(res[0][f], res[1][f], max(np.abs(res[1][f]).max(), FLOOR[f])) < 1e-10, f
    assert np.array_equal(res[0]["loc"], res[1]["loc"])
