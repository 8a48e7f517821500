"""tests/golden/make_golden.py -- regenerate the golden fixtures from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and oracle/_ref, built by `make -C oracle ref`):
    python tests/golden/make_golden.py
Everything written here comes from the reference's own code (oracle/_ref/libulref_*.so: UL_growth_explicit.cpp
+ nondestructivetrimesh.cpp compiled as they lie) or from the reference's shipped input meshes; the C
restatement (oracle/ul_oracle.c) and the CUDA path are checked AGAINST these files, never used to make them.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.refrun import RefSim, ref_eval_element, _load   # noqa: E402
from wholecellfe_b200.meshio import read_ecomsol            # noqa: E402

REF = "/root/reference"
SCRIPTS = REF + "/wholeCellFE_k1/cellBP/scripts/"
DT = 0.005                       # scripts/cell_example.txt: sim_dt 0.005
CELL_FIELDS = ["x", "V", "A", "ea", "cC", "fa_u"]
SUBS_FIELDS = ["V", "A", "ea"]


def snapshot(sim):
    d = {}
    for f in CELL_FIELDS:
        d["c_" + f] = sim.get(0, f)
    for f in SUBS_FIELDS:
        d["s_" + f] = sim.get(1, f)
    return d


def main():
    xc, tc = read_ecomsol(SCRIPTS + "disk_20um_diam_edited.txt")
    xs, ts = read_ecomsol(SCRIPTS + "sq_subs_100um.txt")
    np.savez_compressed(os.path.join(HERE, "example_meshes.npz"), cell_xyz=xc, cell_tri=tc, subs_xyz=xs, subs_tri=ts)
    x2, t2 = read_ecomsol(SCRIPTS + "disk_20um_diam_N1025.txt")
    np.savez_compressed(os.path.join(HERE, "disk_N1025.npz"), cell_xyz=x2, cell_tri=t2)

    # glibc rand() after srand(1): the stream calculateForce_AP consumes (UL_growth_explicit.cpp:1525)
    L = _load("k1")
    L.ulref_srand(1)
    rand_stream = np.array([L.ulref_rand() for _ in range(256)], dtype=np.int64)

    # trajectories: config 1 (k1, 800 steps) and the first steps of k31 / kMD (before rupture chaos sets in)
    out = {"rand_srand1": rand_stream, "dt": DT}
    plan = {"k1": [1, 2, 10, 50, 200, 800], "k31": [1, 2, 5, 10], "kMD": [1, 2, 5, 10]}
    forced = {"k31": 300, "kMD": 300, "k1": 700}   # single-step pairs from a developed state
    for var, marks in plan.items():
        sim = RefSim(var, xc, tc, xs, ts)
        sim.srand(1)
        t = 0.0
        nb = int(sim.boundary_flags(0).sum())
        last = max(max(marks), forced[var] + 1)
        for k in range(1, last + 1):
            if k == forced[var] + 1:
                pre = snapshot(sim)
                pre.update({"s_x": sim.get(1, "x")})
                for key, v in pre.items():
                    out["%s_pre%d_%s" % (var, forced[var], key)] = v
                out["%s_pre%d_t" % (var, forced[var])] = t
                out["%s_pre%d_rand_consumed" % (var, forced[var])] = nb * (k - 1)
            sim.step(DT, t)
            t += DT
            if k == forced[var] + 1:
                for key, v in snapshot(sim).items():
                    out["%s_post%d_%s" % (var, forced[var], key)] = v
            if k in marks:
                for key, v in snapshot(sim).items():
                    out["%s_step%d_%s" % (var, k, key)] = v
                out["%s_step%d_t" % (var, k)] = t
        out[var + "_boundary_flags"] = sim.boundary_flags(0)
        del sim
    np.savez_compressed(os.path.join(HERE, "trajectories.npz"), **out)

    # element known-answer vectors (evalElementRe, both overloads), planar and fully 3-D
    rng = np.random.default_rng(20261017)
    n = 96
    X = rng.normal(size=(n, 3, 3)); x = X + 0.05 * rng.normal(size=(n, 3, 3)); ea = 0.05 * rng.normal(size=(n, 3, 6))
    X[: n // 2, :, 2] = 0; x[: n // 2, :, 2] = 0; ea[: n // 2, :, [2, 4, 5]] = 0
    # keep triangles well shaped: start from an equilateral-ish base
    base = np.array([[0, 0, 0], [1.0, 0, 0], [0.4, 0.9, 0]])
    X = base[None] + 0.2 * X
    x = X + 0.03 * (x - X) / 0.05
    X[: n // 2, :, 2] = 0; x[: n // 2, :, 2] = 0
    mat = 1 + (np.arange(n) % 2)
    time = rng.uniform(0, 3, size=n)
    cen = rng.normal(size=(n, 3)) * 5
    Re = np.zeros((n, 9)); Rea = np.zeros((n, 18)); MMe = np.zeros((n, 3))
    for i in range(n):
        Re[i], Rea[i], MMe[i] = ref_eval_element("k1", int(mat[i]), X[i].ravel(), x[i].ravel(), ea[i].ravel(),
                                                 np.full(3, 100.0), np.zeros(3), DT, cen[i], float(time[i]))
    np.savez_compressed(os.path.join(HERE, "elements.npz"), X=X, x=x, ea=ea, mat=mat, time=time, centroid=cen, Re=Re, Re_ea=Rea, MMe=MMe)
    # .vtk frames written by the reference's own writers (cellBP/iomesh.cpp) from the state at the end of frame 0 (8 steps)
    import tempfile
    vt = {}
    for var in ("k1", "kMD"):
        sim = RefSim(var, xc, tc, xs, ts)
        sim.srand(1)
        t = 0.0
        for _ in range(8):
            sim.step(DT, t)
            t += DT
        with tempfile.TemporaryDirectory() as d:
            sim.write_vtk(0, t, os.path.join(d, "c.vtk")); sim.write_vtk(1, t, os.path.join(d, "s.vtk"), 0)
            vt[var + "_cell_vtk"] = np.frombuffer(open(os.path.join(d, "c.vtk"), "rb").read(), dtype=np.uint8)
            vt[var + "_subs_vtk"] = np.frombuffer(open(os.path.join(d, "s.vtk"), "rb").read(), dtype=np.uint8)
        vt[var + "_t"] = t
        for f in ("x", "cA", "cC", "fa_u", "ea"):
            vt[var + "_c_" + f] = sim.get(0, f)
        vt[var + "_s_ea"] = sim.get(1, "ea")
    np.savez_compressed(os.path.join(HERE, "vtk_frames.npz"), **vt)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
