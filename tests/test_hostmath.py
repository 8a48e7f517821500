"""The kernels' per-element / per-node math (wholecellfe_b200/csrc/wcfe_math.cuh) compiled for the HOST as a test
build (tests/hostmath/hostmath.cpp) and checked against the oracle.  Lets the CPU-only round check catch algebra
errors before GPU time is spent; the product never loads this library."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from conftest import rel_err
from oracle.oracle import eval_element, lib as olib

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)


def d(a):
    return a.ctypes.data_as(_dp)


@pytest.fixture(scope="module")
def hm():
    so = os.path.join(HERE, "hostmath", "libhostmath.so")
    src = os.path.join(HERE, "hostmath", "hostmath.cpp")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++", src, "-o", so])
    L = C.CDLL(so)
    L.hm_philox.restype = C.c_uint
    return L


def test_closed_form_element_matches_literal_restatement(hm, elements):
    e = elements
    worst = 0.0
    for i in range(len(e["X"])):
        mat = int(e["mat"][i]); time = float(e["time"][i])
        mu = 1000.0 if mat == 1 else 1e6
        tact = (time * 200 if time < 2 else 400) if mat == 1 else 0.0
        ea = e["ea"][i]
        e6 = np.ascontiguousarray(1 / 3 * ea[0] + 1 / 3 * ea[1] + 1 / 3 * ea[2])
        r1 = np.zeros(3); r2 = np.zeros(3); eas = np.zeros(6); A0 = C.c_double(); ac = C.c_double()
        hm.hm_elem_eval(d(np.ascontiguousarray(e["X"][i])), d(np.ascontiguousarray(e["x"][i])), d(e6), C.c_double(mu), C.c_double(tact * 0.5),
                        C.c_double(1.0), C.c_double(1.0), d(r1), d(r2), d(eas), C.byref(A0), C.byref(ac))
        Re = e["Re"][i].reshape(3, 3)      # golden: the reference's own evalElementRe
        sc = np.abs(Re).max()
        worst = max(worst, np.abs(Re[1] - r1).max() / sc, np.abs(Re[2] - r2).max() / sc, np.abs(Re[0] + r1 + r2).max() / sc)
        mine = np.array([(1 / 3) * (eas - ea[k]) * A0.value for k in range(3)])
        worst = max(worst, np.abs(e["Re_ea"][i].reshape(3, 6) - mine).max() / (np.abs(eas).max() * A0.value))
        worst = max(worst, abs(e["MMe"][i][0] - A0.value / 3) / e["MMe"][i][0])
    assert worst < 1e-11


def test_point_in_triangle_bit_exact_with_oracle(hm):
    """Same predicate, same rounding: decisions and sub-areas must be identical, including on-edge points."""
    import ctypes
    O = olib()
    rng = np.random.default_rng(3)
    n1 = np.array([0.0, 0.0, 0.0]); n2 = np.array([5.5, 0.25, 0.0]); n3 = np.array([1.0, 6.0, 0.0])
    pts = [n1.copy(), n2.copy(), n3.copy(), 0.5 * (n1 + n2), 0.5 * (n2 + n3), (n1 + n2 + n3) / 3, np.array([7.0, 7.0, 0.0])]
    pts += [np.array([rng.uniform(-1, 7), rng.uniform(-1, 7), 0.0]) for _ in range(500)]
    pts += [n1 + s * (n2 - n1) for s in rng.uniform(0, 1, 100)]           # on an edge up to rounding
    # oracle predicate through a one-triangle simulation is awkward; compare with numpy restating the same IEEE ops
    def ref(pt):
        e1 = n2 - n1; e2 = n3 - n1; ep = pt - n1
        cr = lambda a, b: np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])
        nrm = lambda v: np.sqrt(v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]))
        At = 0.5 * nrm(cr(e1, e2)); a3 = cr(e1, ep); a2 = cr(ep, e2)
        A3 = 0.5 * nrm(a3) if a3[2] >= 0 else -1.0
        A2 = 0.5 * nrm(a2) if a2[2] >= 0 else -1.0
        A1 = At - A2 - A3
        return (A1 >= 0 and A2 >= 0 and A3 >= 0), np.array([A1, A2, A3]), At
    for pt in pts:
        A = np.zeros(3); At = C.c_double()
        got = hm.hm_point_in_triangle(d(n1), d(n2), d(n3), d(np.ascontiguousarray(pt)), d(A), C.byref(At))
        want, Aw, Atw = ref(pt)
        assert bool(got) == bool(want)
        assert np.array_equal(A, Aw) and At.value == Atw


def test_planar_point_in_triangle_is_bit_identical_to_the_3d_form(hm):
    """point_in_triangle_2d (no square roots: |z| instead of sqrt(0 + (0 + z*z))) against the reference-order 3-D predicate on
    planar inputs: same decision, same sub-areas to the last bit -- random triangles of both orientations and very different
    scales, interior, exterior, on-edge and on-vertex points, exact zeros."""
    rng = np.random.default_rng(17)
    n_checked = 0
    for trial in range(400):
        scale = 10.0 ** rng.uniform(-3, 2)
        n = np.zeros((3, 3)); n[:, :2] = rng.uniform(-1, 1, size=(3, 2)) * scale + rng.uniform(-50, 50, size=2)
        if trial % 7 == 0:
            n[:, :2] = np.round(n[:, :2], 1)                              # grid-aligned vertices: exact zeros and ties
        pts = [n[0].copy(), n[1].copy(), n[2].copy(), 0.5 * (n[0] + n[1]), 0.5 * (n[1] + n[2]), 0.5 * (n[0] + n[2]), n.mean(axis=0)]
        for _ in range(20):
            w = rng.dirichlet([1, 1, 1]) if rng.random() < 0.6 else rng.uniform(-0.5, 1.5, size=3)
            pts.append(w @ n / w.sum())
        pts += [n[0] + s * (n[1] - n[0]) for s in rng.uniform(0, 1, 5)]
        for pt in pts:
            pt = np.ascontiguousarray(pt); pt[2] = 0.0
            A3d = np.zeros(3); At3d = C.c_double(); A2d = np.zeros(3); At2d = C.c_double()
            a = hm.hm_point_in_triangle(d(n[0].copy()), d(n[1].copy()), d(n[2].copy()), d(pt), d(A3d), C.byref(At3d))
            b = hm.hm_point_in_triangle_2d(d(n[0].copy()), d(n[1].copy()), d(n[2].copy()), d(pt), d(A2d), C.byref(At2d))
            assert a == b
            assert A3d.tobytes() == A2d.tobytes() and At3d.value == At2d.value
            n_checked += 1
    assert n_checked > 10000


def test_philox_counter_rng_is_a_pure_function(hm):
    a = [hm.hm_philox(1, s, o) for s in range(4) for o in range(8)]
    b = [hm.hm_philox(1, s, o) for s in range(4) for o in range(8)]
    assert a == b and len(set(a)) == len(a)
    assert hm.hm_philox(2, 0, 0) != hm.hm_philox(1, 0, 0)
    u = np.array([hm.hm_philox(5, s, o) for s in range(64) for o in range(64)], dtype=np.float64) / 2**32
    assert abs(u.mean() - 0.5) < 0.02 and abs(u.var() - 1 / 12) < 0.01


def test_frozen_substrate_element_matches_reference_at_nonzero_strain(hm):
    """elem_eval_frozen (S = mu (a^-1 - det a I), the arithmetic of k_subs_ring / k_subs_patch) against the reference's own
    substrate overload of evalElementRe (UL:988-1174) on golden vectors with x == X and planar strain 1e-6 .. 5e-2, both
    orientations; also the C restatement.  Error scale: the element's own |Re| plus the reference's cancellation floor
    (~1e-15 mu L: its long route leaves |Res_subs| ~ 1e-7 on a strain-free substrate)."""
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "elements_frozen.npz"))
    mu = 1e6
    worst = worst_o = worst_ea = 0.0
    for i in range(len(g["X"])):
        X = g["X"][i]; ea = g["ea"][i]
        X6 = np.ascontiguousarray(X[:, :2]); e6 = 1 / 3 * ea[0] + 1 / 3 * ea[1] + 1 / 3 * ea[2]
        e3 = np.ascontiguousarray(e6[[0, 1, 3]])
        r1 = np.zeros(2); r2 = np.zeros(2); A0 = C.c_double()
        hm.hm_elem_eval_frozen(d(X6), d(e3), C.c_double(mu), C.c_double(1.0), d(r1), d(r2), C.byref(A0))
        Re = g["Re"][i].reshape(3, 3)
        L = np.abs(X[1:] - X[0]).max()
        sc = np.abs(Re).max() + 1e-15 * mu * L / 1e-11                      # |Re| + the cancellation floor, expressed at the 1e-11 gate
        worst = max(worst, np.abs(Re[1, :2] - r1).max() / sc, np.abs(Re[2, :2] - r2).max() / sc, np.abs(Re[0, :2] + r1 + r2).max() / sc)
        assert np.abs(Re[:, 2]).max() == 0.0                                # the reference's out-of-plane residual is exactly 0 for a planar state
        # e_s == averaged e for F == I: Re_ea = (e_avg - ea_i) A0 / 3 (UL:1170), MM = A0 / 3
        mine = np.array([(1 / 3) * (e6 - ea[k]) * A0.value for k in range(3)])
        # (the reference forms e_s = (I - b^-1)/2 as a difference from 1: an ABSOLUTE rounding error of ~2e-16 whatever the strain)
        sc_ea = (np.abs(ea).max() + 1e-4) * A0.value
        worst_ea = max(worst_ea, np.abs(g["Re_ea"][i].reshape(3, 6) - mine).max() / sc_ea)
        assert abs(g["MMe"][i][0] - A0.value / 3) <= 1e-14 * A0.value
        Ro, Reao, MMo = eval_element("k1", 2, X, X, ea, 0.005, np.zeros(3), float(g["time"][i]))
        worst_o = max(worst_o, np.abs(Ro - g["Re"][i]).max() / sc, np.abs(Reao - g["Re_ea"][i]).max() / sc_ea)
    assert worst < 1e-11 and worst_ea < 1e-11 and worst_o < 1e-11, (worst, worst_ea, worst_o)


def test_planar_element_matches_reference_vectors(hm, elements):
    """elem_eval_planar (what k_cell_patch evaluates; sqrt-free pressure, metric inverse from the current area) against the
    reference's own evalElementRe on the planar half of the golden element vectors (both materials)."""
    e = elements
    worst = 0.0; n = 0
    for i in range(len(e["X"])):
        if np.abs(e["X"][i][:, 2]).max() != 0 or np.abs(e["x"][i][:, 2]).max() != 0 or np.abs(e["ea"][i][:, [2, 4, 5]]).max() != 0:
            continue
        n += 1
        mat = int(e["mat"][i]); time = float(e["time"][i])
        mu = 1000.0 if mat == 1 else 1e6
        tact = (time * 200 if time < 2 else 400) if mat == 1 else 0.0
        ea = e["ea"][i]
        e6 = 1 / 3 * ea[0] + 1 / 3 * ea[1] + 1 / 3 * ea[2]
        e3 = np.ascontiguousarray(e6[[0, 1, 3]])
        X6 = np.ascontiguousarray(e["X"][i][:, :2]); x6 = np.ascontiguousarray(e["x"][i][:, :2])
        r1 = np.zeros(2); r2 = np.zeros(2); eas = np.zeros(3); A0 = C.c_double(); ac = C.c_double()
        hm.hm_elem_eval_planar(d(X6), d(x6), d(e3), C.c_double(mu), C.c_double(tact * 0.5), C.c_double(1.0), d(r1), d(r2), d(eas), C.byref(A0), C.byref(ac))
        Re = e["Re"][i].reshape(3, 3)
        sc = np.abs(Re).max()
        worst = max(worst, np.abs(Re[1, :2] - r1).max() / sc, np.abs(Re[2, :2] - r2).max() / sc, np.abs(Re[0, :2] + r1 + r2).max() / sc)
        eas6 = np.zeros(6); eas6[[0, 1, 3]] = eas
        mine = np.array([(1 / 3) * (eas6 - ea[k]) * A0.value for k in range(3)])
        ref = e["Re_ea"][i].reshape(3, 6)
        worst = max(worst, np.abs(ref[:, [0, 1, 3]] - mine[:, [0, 1, 3]]).max() / (np.abs(eas).max() * A0.value))
        assert abs(e["MMe"][i][0] - A0.value / 3) <= 1e-14 * A0.value
    assert n >= 40 and worst < 1e-11, (n, worst)


def test_node_table_rows_are_a_permutation_with_fewer_bank_conflicts(hm):
    """Host logic of build_patches (csrc/wcfe_rows.h): the row numbering of a substrate patch's node table must be a permutation
    of [0, NL) whatever the fans look like, and on the bench's own pattern -- a 16 x 8 block of a union-jack grid in Morton order,
    valences 4 and 8, one ring of halo nodes -- it must cost fewer shared-memory wavefronts than numbering the rows like the nodes."""
    def morton(x, y):
        m = 0
        for b in range(8):
            m |= ((x >> b) & 1) << (2 * b) | ((y >> b) & 1) << (2 * b + 1)
        return m
    W_, H_ = 16, 8
    own = sorted(((x, y) for x in range(1, W_ + 1) for y in range(1, H_ + 1)), key=lambda p: morton(p[0] - 1, p[1] - 1))
    P = 128
    assert len(own) == P
    local = {p: i for i, p in enumerate(own)}
    halo = sorted({(x + dx, y + dy) for (x, y) in own for dx in (-1, 0, 1) for dy in (-1, 0, 1)} - set(own))
    for i, p in enumerate(halo):
        local[p] = P + i
    NL = P + (len(halo) + 3) // 4 * 4
    ring8 = [(-1, -1), (0, -1), (1, -1), (1, 0), (1, 1), (0, 1), (-1, 1), (-1, 0)]
    ring4 = [(0, -1), (1, 0), (0, 1), (-1, 0)]
    jk, rptr = [], [0]
    for (x, y) in own:
        ring = ring8 if (x + y) % 2 == 0 else ring4
        for q in range(len(ring)):
            a, b = ring[q], ring[(q + 1) % len(ring)]
            jk += [local[(x + a[0], y + a[1])], local[(x + b[0], y + b[1])]]
        rptr.append(len(jk) // 2)
    jk = np.asarray(jk, dtype=np.uint16); rptr = np.asarray(rptr, dtype=np.uint16)
    row_of = np.zeros(NL, dtype=np.uint16); ident = C.c_longlong()
    u16 = C.POINTER(C.c_ushort)
    hm.hm_ring_table_rows.restype = C.c_longlong
    got = hm.hm_ring_table_rows(jk.ctypes.data_as(u16), len(jk), rptr.ctypes.data_as(u16), P, NL, row_of.ctypes.data_as(u16), C.byref(ident))
    assert sorted(row_of.tolist()) == list(range(NL))
    assert got < ident.value, (got, ident.value)      # 236 against 263 here; on the GPU 27.2 M against 29.0 M wavefronts per launch (profiles/)
    # degenerate input: no fans at all -> identity
    z = np.zeros(P + 1, dtype=np.uint16); e = np.zeros(2, dtype=np.uint16)
    hm.hm_ring_table_rows(e.ctypes.data_as(u16), 0, z.ctypes.data_as(u16), P, NL, row_of.ctypes.data_as(u16), C.byref(ident))
    assert row_of.tolist() == list(range(NL))


def test_integrin_force_laws_match_the_reference_force_field(hm, meshes):
    """integrin_force of wcfe_math.cuh (host build) -- the linear spring of k1 / k31 (UL:1259) and the piecewise-linear MD table of kMD
    (kMD UL:1246-1265, incl. its jump at zero extension and the last segment's extrapolation) -- against the oracle's calculateForce_FA
    (pinned to the reference) on the shipped cell mesh: F_i = -c_C f(fa_u_i) Asum_i rho, with Asum_i taken from the linear law."""
    from oracle.oracle import OracleSim
    import wholecellfe_b200 as W
    xc, tc, xs, ts = meshes
    rng = np.random.default_rng(12)
    n = len(xc)
    mag = 10.0 ** rng.uniform(-9, np.log10(0.03), size=n)                     # from far below the first knot to beyond the last (17e-3)
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]; u *= mag[:, None]
    cC = rng.uniform(0.1, 1.0, size=n)
    F = {}
    for v in ("k1", "kMD"):
        o = OracleSim(v, xc, tc, xs, ts)
        o.set(0, "fa_u", u); o.set(0, "cC", cC)
        F[v] = o.force_field("F_FA", 0.005, 0.0)["F"]
    asum_rho = -F["k1"][:, 0] / (cC * 1000.0 * u[:, 0])                          # Asum_i * rho_i_max
    p = W.default_params("kMD")                                                  # the table the kernels receive (kMD UL:1208-1214)
    px = np.array(list(p.md_px)); py = np.array(list(p.md_py)); kk = np.array(list(p.md_k))
    worst = 0.0
    for i in range(n):
        f = np.zeros(3)
        hm.hm_integrin_force(1, C.c_double(-1.0), d(px), d(py), d(kk), d(np.ascontiguousarray(u[i])), d(f))
        want = -F["kMD"][i] / (cC[i] * asum_rho[i])
        worst = max(worst, np.abs(f - want).max() / np.abs(want).max())
        g = np.zeros(3)
        hm.hm_integrin_force(0, C.c_double(31000.0), d(px), d(py), d(kk), d(np.ascontiguousarray(u[i])), d(g))
        assert np.array_equal(g, 31000.0 * u[i])
    assert worst < 1e-11, worst
    assert mag.min() < 1e-8 and mag.max() > 0.02


@pytest.mark.parametrize("variant,law,kf,kb", [("k1", 0, 1000.0, 1000.0), ("k31", 0, 31000.0, 31000.0), ("kMD", 1, -1.0, -1.0)])
def test_bond_update_matches_the_reference_step_node_by_node(hm, meshes, variant, law, kf, kb):
    """bond_update of wcfe_math.cuh (host build): updateFA's anchor drift, Bell off-rate (2-D norm for k1 / k31, the MD force for kMD),
    forward-Euler c_C, clamp and survivor-weighted anchor / reset branch (UL:2072-2111) -- fed, node by node, with what the oracle's
    step saw (fa_u and c_C before, cell displacement x - X of the step) and compared with what the step left; the reset and clamp
    masks must be the oracle's.  Steps 1-3 (bonds forming from rest) and 38-40 of the shipped example (ruptures under load)."""
    from oracle.oracle import OracleSim
    import wholecellfe_b200 as W
    xc, tc, xs, ts = meshes
    p = W.default_params(variant)
    px = np.array(list(p.md_px)); py = np.array(list(p.md_py)); kk = np.array(list(p.md_k))
    o = OracleSim(variant, xc, tc, xs, ts); o.srand(1)
    t = 0.0; dt = 0.005
    worst = 0.0; resets = 0
    zero = np.zeros(3)
    for k in range(40):
        fa0 = o.get(0, "fa_u"); c0 = o.get(0, "cC")
        o.step(dt, t); t += dt
        if not (k < 3 or k >= 37):
            continue
        cell_u = o.get(0, "x") - o.get(0, "X_prev")
        fa1 = o.get(0, "fa_u"); c1 = o.get(0, "cC")
        reset = o.discrete("fa_reset"); clamp = o.discrete("cc_clamp")
        for i in range(len(xc)):
            fa = np.ascontiguousarray(fa0[i].copy()); cc = C.c_double(c0[i])
            flags = hm.hm_bond_update(law, C.c_double(kf), C.c_double(kb), d(px), d(py), d(kk), C.c_double(dt), d(np.ascontiguousarray(cell_u[i])), d(zero), d(fa), C.byref(cc))
            assert (flags & 1) == reset[i] and ((flags >> 1) & 1) == clamp[i], (variant, k, i)
            worst = max(worst, abs(cc.value - c1[i]) / max(abs(c1[i]), 1e-3), np.abs(fa - fa1[i]).max() / max(np.abs(fa1).max(), 1e-6))
        resets += int(reset.sum())
    assert worst < 1e-12, (variant, worst)
    assert resets > 0                         # step 1 from rest: every node takes the reset branch (c_C was 0)


def test_boundary_math_matches_the_reference_force_fields(hm, meshes):
    """bnd_normal / bnd_curvature / ap_trips of wcfe_math.cuh (host build) against the oracle's calculateBendingForce2D (UL:1550-1639),
    calculateAreaConstraintForce (UL:1670-1763) and the trip count of calculateForce_AP's loop (UL:1530) on a developed state of the
    shipped cell: F_bend = k_bend kappa n, F_area = -k_area (A - A0) / 2 L n at every boundary vertex."""
    from oracle.oracle import OracleSim
    xc, tc, xs, ts = meshes
    o = OracleSim("k1", xc, tc, xs, ts); o.srand(1)
    t = 0.0
    for _ in range(60):
        o.step(0.005, t); t += 0.005
    x = o.get(0, "x")
    fb = o.force_field("F_bend", 0.005, t); fa = o.force_field("F_area", 0.005, t)
    area = fb["area"][0]
    p = o.params
    nb = 0; worst = 0.0
    sb = np.abs(fb["F"]).max(); sa = np.abs(fa["F"]).max()
    for i in range(len(xc)):
        nbr = o.boundary_neighbours(i)
        if nbr is None:
            assert not fb["F"][i].any() and not fa["F"][i].any()
            continue
        nb += 1
        nrm = np.zeros(3); tl = C.c_double(); kap = C.c_double(); trips = C.c_double()
        hm.hm_boundary(d(np.ascontiguousarray(x[i])), d(nbr[0]), d(nbr[1]), C.c_double(p.ap_density), d(nrm), C.byref(tl), C.byref(kap), C.byref(trips))
        worst = max(worst, np.abs(p.k_bend * kap.value * nrm - fb["F"][i]).max() / sb,
                    np.abs(-p.k_area * (area - p.area_0) / 2 * tl.value * nrm - fa["F"][i]).max() / sa)
        n_it = 0
        while n_it < p.ap_density * tl.value / 2:           # UL:1530, literally
            n_it += 1
        assert trips.value == n_it
    assert nb >= 40 and worst < 1e-12, (nb, worst)


def test_nodal_update_matches_the_reference_step(hm, meshes):
    """node_update / node_update_q / node_update_r of wcfe_math.cuh (host build) against the oracle's cell update (UL:592-628): the net
    force assembled from the step's own force fields (drag on V_hlf, bending, actin, integrins, area constraint, minus the residual),
    then A = (f - damp M rho V) / (M rho) and V = V_hlf + (dt_real - dt/2) A, for every node and component of steps 20-22 of k1.
    The quotient form is bit-identical to an IEEE division; the reciprocal forms stay within 2 ulps."""
    from oracle.oracle import OracleSim
    xc, tc, xs, ts = meshes
    o = OracleSim("k1", xc, tc, xs, ts); o.srand(1)
    p = o.params
    t = 0.0; dt = 0.005
    worst = [0.0, 0.0, 0.0]
    for k in range(23):
        V0 = o.get(0, "V"); A0 = o.get(0, "A")
        o.step(dt, t); t += dt
        if k < 20:
            continue
        Vh = V0 + (dt / 2.0) * A0; Vh[:, 2] = 0.0                                   # UL:118-120
        ff = {n: o.step_force_field(n) for n in ["Res_cell", "F_FA", "F_AP", "F_bend", "F_area"]}
        f = -p.drag * Vh + ff["F_bend"]["F"] + ff["F_AP"]["F"] + ff["F_FA"]["F"] + ff["F_area"]["F"] - ff["Res_cell"]["F"]   # UL:568-592
        MM = ff["Res_cell"]["MM"]
        V1 = o.get(0, "V"); A1 = o.get(0, "A")
        out = np.zeros(6)
        for i in range(len(xc)):
            for c in range(3):
                hm.hm_node_update(C.c_double(f[i, c]), C.c_double(MM[i]), C.c_double(p.damp), C.c_double(p.rho), C.c_double(dt), C.c_double(dt),
                                  C.c_double(Vh[i, c]), C.c_double(V0[i, c]), C.c_double(A0[i, c]), d(out))
                assert out[0] == V1[i, c] and out[1] == A1[i, c], (k, i, c)          # the division-free quotient is an IEEE division
                for q in (1, 2):
                    sa = max(abs(A1[i, c]), 1e-300)
                    worst[q] = max(worst[q], abs(out[2 * q + 1] - A1[i, c]) / sa)
    assert worst[1] == 0.0 and worst[2] < 5e-16, worst
