"""The C restatement (oracle/ul_oracle.c) against golden vectors generated from the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
from conftest import rel_err
from oracle.oracle import OracleSim, eval_element, glibc_rand_stream

CELL_FIELDS = ["x", "V", "A", "ea", "cC", "fa_u"]
SUBS_FIELDS = ["V", "A", "ea"]


def test_glibc_rand_restatement(traj):
    ref = traj["rand_srand1"]
    assert np.array_equal(glibc_rand_stream(1, len(ref)), ref)


def test_element_known_answers(elements):
    e = elements
    worst = 0.0
    for i in range(len(e["X"])):
        Re, Rea, MMe = eval_element("k1", int(e["mat"][i]), e["X"][i].ravel(), e["x"][i].ravel(), e["ea"][i].ravel(), 0.005,
                                    e["centroid"][i], float(e["time"][i]))
        worst = max(worst, rel_err(Re, e["Re"][i]), rel_err(Rea, e["Re_ea"][i]), rel_err(MMe, e["MMe"][i]))
    assert worst < 1e-12


@pytest.mark.parametrize("variant,marks,tol", [("k1", [1, 2, 10, 50, 200, 800], 2e-11), ("k31", [1, 2, 5, 10], 1e-12), ("kMD", [1, 2, 5, 10], 1e-11)])
def test_trajectory_against_reference(meshes, traj, variant, marks, tol):
    xc, tc, xs, ts = meshes
    sim = OracleSim(variant, xc, tc, xs, ts)
    sim.srand(1)
    dt = float(traj["dt"]); t = 0.0
    for k in range(1, max(marks) + 1):
        sim.step(dt, t); t += dt
        if k in marks:
            for f in CELL_FIELDS:
                assert rel_err(sim.get(0, f), traj["%s_step%d_c_%s" % (variant, k, f)]) < tol, (variant, k, f)
            for f in SUBS_FIELDS:
                ref = traj["%s_step%d_s_%s" % (variant, k, f)]
                # substrate fields at rest are cancellation noise of mu=1e6 (quirk q7): scale by the cell's field
                scale = max(np.abs(ref).max(), np.abs(traj["%s_step%d_c_%s" % (variant, k, f)]).max() * 1e-3, 1e-12)
                assert rel_err(sim.get(1, f), ref, scale) < 1e-9, (variant, k, f)


@pytest.mark.parametrize("variant,k0", [("k1", 700), ("k31", 300), ("kMD", 300)])
def test_single_step_from_developed_state(meshes, traj, variant, k0):
    """k31 / kMD are chaotic (bond-rupture avalanches amplify 1 ulp by ~10x every 4 steps), so far from t=0 the
    comparison is one step from the reference's own state."""
    xc, tc, xs, ts = meshes
    sim = OracleSim(variant, xc, tc, xs, ts)
    pre = "%s_pre%d_" % (variant, k0); post = "%s_post%d_" % (variant, k0)
    for f in CELL_FIELDS:
        sim.set(0, f, traj[pre + "c_" + f])
    for f in SUBS_FIELDS:
        sim.set(1, f, traj[pre + "s_" + f])
    sim.srand(1)
    sim.rng_discard(int(traj[pre + "rand_consumed"]))
    sim.step(float(traj["dt"]), float(traj[pre + "t"]))
    for f in CELL_FIELDS:
        assert rel_err(sim.get(0, f), traj[post + "c_" + f]) < 1e-12, (variant, f)


def test_grid_point_location_equals_brute_force(meshes, traj):
    xc, tc, xs, ts = meshes
    a = OracleSim("k31", xc, tc, xs, ts, use_grid=False); b = OracleSim("k31", xc, tc, xs, ts, use_grid=True)
    t = 0.0
    for k in range(40):
        a.step(0.005, t); b.step(0.005, t); t += 0.005
    for f in CELL_FIELDS:
        assert np.array_equal(a.get(0, f), b.get(0, f))
    for w in ["loc_bond", "loc_force_cnt", "loc_force_first"]:
        assert np.array_equal(a.discrete(w), b.discrete(w))
    assert np.array_equal(a.get(1, "A"), b.get(1, "A"))


@pytest.mark.parametrize("name", ["explicit", "biaxial", "uniaxial", "offbiaxial", "bulgetest"])
def test_every_driver_against_reference_vectors(drivers, name):
    """ulo_driver_kind against the reference's five driverVelocities_* (UL:103-345) on nodes lying on, and one ulp either side of,
    every threshold, at times on and around every switch of the time windows: bit for bit."""
    from oracle.oracle import driver
    d, expect = drivers
    for it, t in enumerate(d["times"]):
        got = driver("k1", name, d["X"], d["V"], d["A"], float(d["dt"]), float(t))
        assert np.array_equal(got, expect(name, it)), (name, t, int((got != expect(name, it)).sum()))


def test_dormant_fields_against_reference_vectors(meshes, dormant):
    """The restatement of the reference's dormant physics -- diffusion residuals (UL:915-925), pressure load (UL:1765-1816), sphere
    penalty (UL:1178-1194) -- against what the reference's own functions return on a bent, unevenly loaded cell."""
    xc, tc, xs, ts = meshes
    g = dormant
    o = OracleSim("k1", xc, tc, xs, ts)
    for f in ["x", "X_prev", "ea", "cA", "cC"]:                      # x before X_prev: setting x alone also resets X_prev
        o.set(0, f, g["c_" + f])
    for name in ["Res_Adif", "Res_Cdif"]:
        for q, dt in enumerate(g["dts"]):
            assert rel_err(o.force_field(name, float(dt), 0.7)["F"], g[name][q]) < 1e-13, (name, dt)
    for q, t in enumerate(g["times"]):
        assert rel_err(o.force_field("F_pressure", 0.005, float(t))["F"], g["F_pressure"][q], max(np.abs(g["F_pressure"][q]).max(), 1.0)) < 1e-13, t
    assert np.abs(g["F_pressure"][-1]).max() == 0.0 and np.abs(g["F_pressure"][1]).max() > 1e4      # the schedule's branches are all there
    assert rel_err(o.force_field("F_sphere", 0.005, 0.7)["F"], g["F_sphere"]) < 1e-14
