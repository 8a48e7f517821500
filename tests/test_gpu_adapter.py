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


DRIVER_NAMES = ["explicit", "biaxial", "uniaxial", "offbiaxial", "bulgetest"]


@pytest.mark.parametrize("name", DRIVER_NAMES)
def test_every_driver_on_the_device_against_reference_vectors(drivers, name):
    """k_driver through wcfe_driver_velocities_kind, and through the adapter's own driverVelocities_* symbols, against the reference's
    five drivers (UL:103-345; three of them commented out in main.cpp) on nodes on / one ulp off every threshold: bit for bit."""
    import wholecellfe_b200 as W
    from oracle.refrun import driver as rdriver
    d, expect = drivers
    g = W.ULSim("k1", device=0)
    n0 = g.kernel_launches()
    for it, t in enumerate(d["times"]):
        for surface in (W.CELL, W.SUBS):
            got = g.driver_velocities(surface, d["X"], d["V"], d["A"], float(d["dt"]), float(t), kind=name)
            assert np.array_equal(got, expect(name, it)), (name, t, surface)
        via_adapter = rdriver("k1", name, d["X"], d["V"], d["A"], float(d["dt"]), float(t), lib_kind="adapter")
        assert np.array_equal(via_adapter, expect(name, it)), (name, t, "adapter")
    assert g.kernel_launches() - n0 == 2 * len(d["times"])
    with pytest.raises(W.WcfeError):
        g.driver_velocities(W.SUBS, d["X"], d["V"], d["A"], 0.005, 0.0, kind=7)
    g.close()


@pytest.mark.parametrize("name,kind", [("uniaxial", 2), ("offbiaxial", 3), ("bulgetest", 4)])
def test_substrate_step_with_a_dormant_driver_through_the_adapter(meshes, name, kind):
    """A maintainer who swaps the substrate's driver at main.cpp:291-295 for one of the commented-out alternatives: the same call
    sequence against the reference's own UL_growth_explicit.o and against the adapter + libwcfe.so, k1, 8 steps.  The substrate's
    V = V_hlf + (dt_real - dt/2) A (UL:499) must come out with THAT driver's V_hlf (imposed 0.1 / pins), not _Biaxial's."""
    from oracle.refrun import ref_available
    if not ref_available("k1"):
        pytest.skip("oracle/_ref/libulref_k1.so not built here")
    xc, tc, xs, ts = meshes
    a = RefSim("k1", xc, tc, xs, ts, kind="adapter"); a.srand(1)
    r = RefSim("k1", xc, tc, xs, ts, kind="ref"); r.srand(1)
    rng = np.random.default_rng(4)
    v0 = 1e-2 * rng.normal(size=(len(xs), 3)); v0[:, 2] = 0.0        # a substrate at rest would hide the pins
    for s in (a, r):
        s.set(1, "V", v0)
    t = 0.0
    for k in range(8):
        a.step_driver(0.005, t, kind); r.step_driver(0.005, t, kind); t += 0.005
        for f in ["x", "V", "A", "ea", "cC", "fa_u"]:
            ref = r.get(0, f)
            assert rel_err(a.get(0, f), ref, max(np.abs(ref).max(), FLOOR[f])) < 1e-10, (name, k, f)
        for f in ["V", "A", "ea"]:
            ref = r.get(1, f)
            assert rel_err(a.get(1, f), ref, max(np.abs(ref).max(), SUBS_FLOOR[f])) < 1e-9, (name, k, f)
    assert np.array_equal(a.get(1, "x"), xs)
    # the comparison is not vacuous: the shipped driver leaves another substrate velocity (run AFTER r: the reference draws from
    # libc's one global rand() stream)
    b = RefSim("k1", xc, tc, xs, ts, kind="ref"); b.srand(1)
    b.set(1, "V", v0)
    t = 0.0
    for k in range(8):
        b.step(0.005, t); t += 0.005
    assert np.abs(r.get(1, "V") - b.get(1, "V")).max() > 1e-2, name
    assert rel_err(r.get(0, "V"), b.get(0, "V")) < 1e-6       # ... and the same cell (the frozen substrate's velocity does not reach it)


def test_dormant_driver_rules_at_the_c_abi(meshes):
    """include/wcfe.h: a substrate driver other than _Biaxial is honoured by the next substrate update on the planar AND the general
    kernels alike and lasts for one update; the cell's update refuses after a driver other than _Explicit (its drag force reads
    V_hlf, UL:571-573) until the shipped driver is called again."""
    import wholecellfe_b200 as W
    xc, tc, xs, ts = meshes
    rng = np.random.default_rng(8)
    v0 = 1e-2 * rng.normal(size=(len(xs), 3)); v0[:, 2] = 0.0
    out = {}
    for general in (0, 1):
        g = W.ULSim("k1", device=0)
        g.upload_mesh(W.CELL, xc, tc); g.upload_mesh(W.SUBS, xs, ts); g.srand(1)
        g.set_force_general(general)
        g.set(W.SUBS, "V", v0)
        assert bool(g.path_info()["planar"]) == (general == 0)
        t = 0.0
        for k in range(3):
            g.predict(W.CELL, 0.005, t, download=False); g.commit(W.CELL)
            vh = g.driver_velocities(W.SUBS, xs, g.get(W.SUBS, "V"), g.get(W.SUBS, "A"), 0.005, t, kind="uniaxial")
            g.update_cell(0.005, 0.005, t); g.update_subs(0.005, 0.005, t); t += 0.005
            A = g.get(W.SUBS, "A")
            assert np.abs(g.get(W.SUBS, "V") - (vh + 0.0025 * A)).max() <= 1e-15 * max(1.0, np.abs(A).max())
        out[general] = (g.get(W.SUBS, "V"), g.get(W.SUBS, "A"), g.get(W.CELL, "V"))
        # the choice lasted for one update: a plain step afterwards is _Biaxial's again
        vb = g.get(W.SUBS, "V"); ab = g.get(W.SUBS, "A")
        g.step(0.005, t)
        free = (np.abs(xs[:, 0]) <= 49.0) & (np.abs(xs[:, 1]) <= 49.0)
        A2 = g.get(W.SUBS, "A")
        assert np.abs(g.get(W.SUBS, "V")[free] - (vb + 0.0025 * ab + 0.0025 * A2)[free]).max() <= 1e-12 * max(1.0, np.abs(A2).max())
        # the cell
        g.predict(W.CELL, 0.005, t, download=False); g.commit(W.CELL)
        g.driver_velocities(W.CELL, xc, g.get(W.CELL, "V"), g.get(W.CELL, "A"), 0.005, t, kind="offbiaxial")
        with pytest.raises(W.WcfeError) as e:
            g.update_cell(0.005, 0.005, t)
        assert e.value.code == -3
        g.driver_velocities(W.CELL, xc, g.get(W.CELL, "V"), g.get(W.CELL, "A"), 0.005, t)       # the shipped driver again
        g.update_cell(0.005, 0.005, t); g.update_subs(0.005, 0.005, t)
        g.close()
    for q in range(3):
        s = max(np.abs(out[1][q]).max(), 1e-3)
        assert np.abs(out[0][q] - out[1][q]).max() / s < 1e-9, q
