"""tests/golden/make_golden_dormant.py -- known-answer vectors for the reference's dormant physics.

Run in the build container only (needs oracle/_ref, built by `make -C oracle ref`):
    python tests/golden/make_golden_dormant.py
Functions the reference defines and never calls, evaluated by the reference itself (oracle/_ref/libulref_k1.so) on a state of the
shipped example that was bent out of its plane (the pressure load's vertex normals and the sphere penalty are 3-D) and given uneven
concentrations (the diffusion gradients):
  Res_Adif, Res_Cdif   outputs of calculateRes(cell) (UL:632-707; element part UL:833-842, :915-925) that no statement reads (UL:617-618)
  F_pressure           calculatePressureForce UL:1765-1816 with the lumped mass of calculateRes, at times on every branch of its schedule
  F_sphere             calculateConstraintForce UL:1178-1194
The C restatement (oracle/ul_oracle.c) and the device kernel (k_dormant, through wcfe_dormant_field) are checked AGAINST this file.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.refrun import RefSim   # noqa: E402

TIMES = [0.7, 1.0, 1.5, 2.0, 2.4, 3.0, 3.5]
DTS = [0.005, 0.5]


def bent_state(sim, xc):
    """10 steps of the shipped example, then out-of-plane noise on x and uneven c_A, then 2 more steps (so that the start-of-step
    positions X are out of plane too); the same recipe drives RefSim, OracleSim and -- through the oracle -- the device tests"""
    sim.srand(1)
    t = 0.0
    for _ in range(10):
        sim.step(0.005, t); t += 0.005
    rng = np.random.default_rng(3)
    x = sim.get(0, "x"); x[:, 2] = 0.3 * rng.normal(size=len(xc))
    sim.set(0, "x", x); sim.set(0, "cA", 100 + 10 * rng.normal(size=len(xc)))
    for _ in range(2):
        sim.step(0.005, t); t += 0.005
    return t


def main():
    d = np.load(os.path.join(HERE, "example_meshes.npz"))
    xc, tc, xs, ts = d["cell_xyz"], d["cell_tri"], d["subs_xyz"], d["subs_tri"]
    r = RefSim("k1", xc, tc, xs, ts)
    bent_state(r, xc)
    out = {"times": np.array(TIMES), "dts": np.array(DTS)}
    for f in ["x", "X_prev", "ea", "cA", "cC"]:
        out["c_" + f] = r.get(0, f)
    for name in ["Res_Adif", "Res_Cdif"]:
        out[name] = np.stack([r.force_field(name, dt, 0.7)["F"] for dt in DTS])
    out["F_pressure"] = np.stack([r.force_field("F_pressure", 0.005, t)["F"] for t in TIMES])
    out["F_sphere"] = r.force_field("F_sphere", 0.005, 0.7)["F"]
    np.savez_compressed(os.path.join(HERE, "dormant.npz"), **out)
    print("wrote dormant.npz: |z| up to %.3f, max |Res_Adif| %.4g (dt term %.3g), |Res_Cdif| %.4g, |F_pressure| %s, |F_sphere| %.4g" % (
        np.abs(out["c_x"][:, 2]).max(), np.abs(out["Res_Adif"]).max(), np.abs(out["Res_Adif"][1] - out["Res_Adif"][0]).max(),
        np.abs(out["Res_Cdif"]).max(), np.round(np.abs(out["F_pressure"]).max(axis=(1, 2)), 1), np.abs(out["F_sphere"]).max()))


if __name__ == "__main__":
    main()
