"""The literal drop-in: oracle/_ref/libuladapter_*.so is the reference-side driver (oracle/ref_wrap.cpp: std::vector<Vec3d>,
NonDestructiveTriMesh, Eigen::VectorXd, main.cpp's call sequence) linked against wholecellfe_b200/csrc/ul_adapter.cpp +
libwcfe.so INSTEAD of the reference's UL_growth_explicit.o.  Built in the container that has the reference headers
(`make -C oracle adapter`); skipped where the prebuilt file is absent.  GPU only."""
import numpy as np
import pytest
from conftest import rel_err
from oracle.oracle import OracleSim
from oracle.refrun import RefSim, adapter_available

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not adapter_available("k1"), reason="oracle/_ref/libuladapter_*.so not built here")]
FLOOR = {"ea": 1e-3, "fa_u": 1e-4, "V": 1e-3, "A": 1.0, "cC": 1e-3, "x": 1.0}
SUBS_FLOOR = {"V": 0.1, "A": 10.0, "ea": 1e-3}


@pytest.mark.parametrize("variant,nsteps", [("k1", 40), ("k31", 12)])
def test_reference_call_sequence_through_the_adapter(meshes, variant, nsteps):
    xc, tc, xs, ts = meshes
    a = RefSim(variant, xc, tc, xs, ts, kind="adapter"); a.srand(1)
    o = OracleSim(variant, xc, tc, xs, ts); o.srand(1)
    t = 0.0
    for k in range(nsteps):
        a.step(0.005, t); o.step(0.005, t); t += 0.005
        for f in ["x", "V", "A", "ea", "cC", "fa_u"]:
            ref = o.get(0, f)
            assert rel_err(a.get(0, f), ref, max(np.abs(ref).max(), FLOOR[f])) < 1e-10, (variant, k, f)
        for f in ["V", "A", "ea"]:
            ref = o.get(1, f)
            assert rel_err(a.get(1, f), ref, max(np.abs(ref).max(), SUBS_FLOOR[f])) < 1e-9, (variant, k, f)
    assert np.array_equal(a.get(1, "x"), xs)          # frozen substrate (main.cpp:299)


def test_adapter_kmd_single_steps_from_oracle_state(meshes):
    """kMD is discontinuous at zero extension: compare step by step from the oracle's state"""
    xc, tc, xs, ts = meshes
    a = RefSim("kMD", xc, tc, xs, ts, kind="adapter"); a.srand(1)
    o = OracleSim("kMD", xc, tc, xs, ts); o.srand(1)
    t = 0.0
    for k in range(30):
        for f in ["x", "V", "A", "ea", "cC", "fa_u"]:
            a.set(0, f, o.get(0, f))
        for f in ["V", "A", "ea"]:
            a.set(1, f, o.get(1, f))
        a.step(0.005, t); o.step(0.005, t); t += 0.005
        for f in ["x", "V", "A", "ea", "cC", "fa_u"]:
            ref = o.get(0, f)
            assert rel_err(a.get(0, f), ref, max(np.abs(ref).max(), FLOOR[f])) < 1e-10, (k, f)


@pytest.mark.parametrize("frac_cell,frac_subs", [(0.8, 1.0), (0.4, 1.0), (1.0, 0.5), (0.7, 0.6)])
def test_adapter_when_the_host_integrator_shortens_the_step(meshes, frac_cell, frac_subs):
    """El Topo's integrate() may accept less than the requested dt (collision_safety on): main.cpp then forms the substrate's --
    and, in the redo branch, the cell's -- V_hlf with the SHORTENED dt while updateForces_Explicit still receives ori_dt
    (main.cpp:244-361).  The same driver sequence, once against the reference's own UL_growth_explicit.o and once against the
    adapter + libwcfe.so, k1, 12 steps with every third one shortened: the velocity update V = V_hlf + (dt_real - dt/2) A
    (UL:497, :622) must use the caller's V_hlf, not one rebuilt from ori_dt."""
    from oracle.refrun import ref_available
    if not ref_available("k1"):
        pytest.skip("oracle/_ref/libulref_k1.so not built here")
    xc, tc, xs, ts = meshes
    a = RefSim("k1", xc, tc, xs, ts, kind="adapter"); a.srand(1)
    r = RefSim("k1", xc, tc, xs, ts, kind="ref"); r.srand(1)
    t = 0.0
    for k in range(12):
        if k % 3 == 2:
            a.step_short(0.005, t, frac_cell, frac_subs); r.step_short(0.005, t, frac_cell, frac_subs)
            t += 0.005 * frac_cell * frac_subs
        else:
            a.step(0.005, t); r.step(0.005, t); t += 0.005
        for f in ["x", "V", "A", "ea", "cC", "fa_u"]:
            ref = r.get(0, f)
            assert rel_err(a.get(0, f), ref, max(np.abs(ref).max(), FLOOR[f])) < 1e-10, (k, f)
        for f in ["V", "A", "ea"]:
            ref = r.get(1, f)
            assert rel_err(a.get(1, f), ref, max(np.abs(ref).max(), SUBS_FLOOR[f])) < 1e-9, (k, f)
