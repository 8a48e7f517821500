"""The C restatement against the UNMODIFIED reference translation units (oracle/_ref, built by `make -C oracle ref`
where /root/reference exists).  Skipped on boxes without oracle/_ref."""
import numpy as np
import pytest
from conftest import rel_err
from oracle.oracle import OracleSim, eval_element
from oracle.refrun import RefSim, ref_available, ref_eval_element

pytestmark = pytest.mark.skipif(not ref_available("k1"), reason="oracle/_ref not built here")


@pytest.mark.parametrize("variant,nsteps", [("k1", 40), ("k31", 12), ("kMD", 12)])
def test_first_steps_all_fields(meshes, variant, nsteps):
    """k31 / kMD: only the first dozen steps -- after that bond-rupture avalanches amplify rounding differences
    by ~10x every 4 steps (measured between this restatement and the reference itself), see DESIGN.md."""
    xc, tc, xs, ts = meshes
    r = RefSim(variant, xc, tc, xs, ts); r.srand(1)
    o = OracleSim(variant, xc, tc, xs, ts); o.srand(1)
    t = 0.0
    for k in range(nsteps):
        r.step(0.005, t); o.step(0.005, t); t += 0.005
        for f in ["x", "V", "A", "ea", "cC", "fa_u"]:
            assert rel_err(o.get(0, f), r.get(0, f)) < 1e-10, (variant, k, f)
    # the in-step force fields of the restatement agree with the reference's functions on the same state
    for name in ["F_FA", "F_bend", "F_area", "Res_cell", "Res_subs", "F_FA_subs"]:
        a = o.force_field(name, 0.005, t); b = r.force_field(name, 0.005, t)
        scale = max(np.abs(b["F"]).max(), 1.0)
        assert rel_err(a["F"], b["F"], scale) < 1e-10, name


@pytest.mark.parametrize("variant,scale", [("k1", 1.0), ("k31", 0.5)])
def test_moving_substrate_restatement_against_the_reference(meshes, variant, scale):
    """The substrate's predictor without the literal `0*` of main.cpp:299 (x_new = x + scale * dt * V_hlf): the restatement's
    moving-substrate branch -- displacement interpolation at the anchors in the start-of-step coordinates, last match wins
    (UL:2026-2065), general substrate element (UL:988-1174) -- against the unmodified reference functions driven through the same
    sequence, substrate started with velocity and strain so that it moves at once."""
    xc, tc, xs, ts = meshes
    r = RefSim(variant, xc, tc, xs, ts); r.srand(1)
    o = OracleSim(variant, xc, tc, xs, ts); o.srand(1); o.params.subs_scale = scale
    rng = np.random.default_rng(23)
    # a smooth shear / stretch field plus a little noise (um/s); dt = 1e-3 keeps the mu = 1e6 substrate's explicit step stable
    V = np.zeros((len(xs), 3)); V[:, 0] = 2.0 * xs[:, 1] / 50 + 0.02 * rng.normal(size=len(xs)); V[:, 1] = 1.2 * xs[:, 0] / 50 + 0.02 * rng.normal(size=len(xs))
    ea = np.zeros((len(xs), 6)); ea[:, [0, 1, 3]] = 1e-3 * rng.normal(size=(len(xs), 3))
    for f, v in (("V", V), ("ea", ea)):
        r.set(1, f, v); o.set(1, f, v)
    t = 0.0
    x0 = r.get(1, "x").copy()
    for k in range(10):
        r.step(0.001, t, subs_scale=scale); o.step(0.001, t); t += 0.001
        for f in ["x", "V", "A", "ea", "cC", "fa_u"]:
            assert rel_err(o.get(0, f), r.get(0, f)) < 1e-10, (variant, k, f)
        for f in ["x", "V", "A", "ea"]:
            b = r.get(1, f)
            assert rel_err(o.get(1, f), b, np.abs(b).max()) < 1e-10, (variant, k, f)
    assert np.abs(r.get(1, "x") - x0).max() > 5e-3 and np.isfinite(r.get(1, "A")).all()


def test_random_elements_3d_and_planar():
    rng = np.random.default_rng(7)
    base = np.array([[0, 0, 0], [1.0, 0, 0], [0.4, 0.9, 0]])
    for trial in range(200):
        X = base + 0.15 * rng.normal(size=(3, 3)); x = X + 0.02 * rng.normal(size=(3, 3)); ea = 0.04 * rng.normal(size=(3, 6))
        if trial % 2 == 0:
            X[:, 2] = 0; x[:, 2] = 0; ea[:, [2, 4, 5]] = 0
        mat = 1 + trial % 2
        cen = rng.normal(size=3)
        a = eval_element("k1", mat, X.ravel(), x.ravel(), ea.ravel(), 0.005, cen, 1.3)
        b = ref_eval_element("k1", mat, X.ravel(), x.ravel(), ea.ravel(), np.full(3, 100.0), np.zeros(3), 0.005, cen, 1.3)
        for u, v in zip(a, b):
            assert rel_err(u, v) < 1e-12


@pytest.mark.parametrize("variant", ["k1", "k31", "kMD"])
def test_every_driver_restatement_against_the_reference_functions(variant):
    """The five driverVelocities_* of each variant's own translation unit (the three copies are identical, UL:103-345) against
    ulo_driver_kind on fresh random nodes and times, bit for bit."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_drivers import inputs
    from oracle.oracle import driver as odriver
    from oracle.refrun import driver as rdriver, DRIVERS
    X, V, A = inputs(seed={"k1": 5, "k31": 6, "kMD": 7}[variant], n=1500)
    rng = np.random.default_rng(99)
    for t in list(rng.uniform(0, 9, 6)) + [1.0, 2.0, 3.0, 8.0]:
        dt = float(rng.uniform(1e-4, 1e-2))
        for name in DRIVERS:
            assert np.array_equal(odriver(variant, name, X, V, A, dt, float(t)), rdriver(variant, name, X, V, A, dt, float(t))), (variant, name, t)


@pytest.mark.parametrize("variant", ["k1", "kMD"])
def test_dormant_fields_restatement_against_the_reference_functions(meshes, variant):
    """Diffusion residuals, pressure load and sphere penalty (calls commented out in the reference, UL:617-618, :577, :551): the
    restatement against the reference's own functions on a state bent out of plane by the same recipe on both sides."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden_dormant import bent_state
    xc, tc, xs, ts = meshes
    r = RefSim(variant, xc, tc, xs, ts); bent_state(r, xc)
    o = OracleSim(variant, xc, tc, xs, ts); bent_state(o, xc)
    assert np.abs(r.get(0, "X_prev")[:, 2]).max() > 0.1
    for name, dt, t in [("Res_Adif", 0.005, 0.7), ("Res_Adif", 0.5, 0.7), ("Res_Cdif", 0.5, 0.7), ("F_pressure", 0.005, 0.4), ("F_pressure", 0.005, 1.5),
                        ("F_pressure", 0.005, 2.6), ("F_sphere", 0.005, 0.0)]:
        a = o.force_field(name, dt, t)["F"]; b = r.force_field(name, dt, t)["F"]
        assert rel_err(a, b) < 1e-12, (variant, name, dt, t)
