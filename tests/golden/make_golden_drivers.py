"""tests/golden/make_golden_drivers.py -- known-answer vectors for every driverVelocities_* the reference defines.

Run in the build container only (needs oracle/_ref, built by `make -C oracle ref`):
    python tests/golden/make_golden_drivers.py
Every V_hlf (stored as a code per component: free value V + dt/2 A, zero, or the imposed 0.1; checked here to reproduce the
reference's output bit for bit) comes from the reference's own functions (UL_growth_explicit.cpp:103-345: _Explicit, _Biaxial -- the two main.cpp
calls -- and _Uniaxial, _OffBiaxial, _BulgeTest, which main.cpp:217-220, :291-295 keep commented out; the three variants' copies
are identical) through oracle/_ref/libulref_k1.so.  The C restatement (oracle/ul_oracle.c ulo_driver_kind) and the device kernel
(k_driver, through wcfe_driver_velocities_kind) are checked AGAINST this file, bit for bit.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.refrun import driver, DRIVERS   # noqa: E402

TIMES = [0.0, 0.5, 0.99, 1.0, 1.5, 2.0, 2.5, 3.0, 5.0, 8.0, 9.0]


def inputs(seed=20261019, n=2048):
    """Nodes of a unit sheet (the geometry the three dormant drivers are written for) and of the +-50 um substrate (_Biaxial's), with a
    share of them exactly ON a threshold or one ulp to either side of it: every comparison in the drivers is strict."""
    rng = np.random.default_rng(seed)
    X = np.empty((n, 3))
    X[:, 0] = rng.uniform(-0.1, 1.3, n); X[:, 1] = rng.uniform(-0.1, 1.1, n)
    X[:, 2] = rng.choice([0.0, 0.0005, -0.0005, 0.001, -0.001, 0.002, -0.002], n)
    big = rng.random(n) < 0.25
    X[big, 0] = rng.uniform(-52, 52, big.sum()); X[big, 1] = rng.uniform(-52, 52, big.sum())
    edge0 = [-0.0499, 0.9499, 0.9499 + 2.00 * 0.1, -0.04, 0.94, 0.94 + 0.99 * 0.1, 0.001, 0.999, 49.0, -49.0] + [0.9499 + 0.1 * t for t in TIMES] + [0.94 + 0.1 * t for t in TIMES]
    edge1 = [0.99, 0.01, 0.001, 0.999, 49.0, -49.0]
    for q in range(0, n, 5):
        v = edge0[(q // 5) % len(edge0)]
        X[q, 0] = [v, np.nextafter(v, np.inf), np.nextafter(v, -np.inf)][(q // 5) % 3]
    for q in range(2, n, 7):
        v = edge1[(q // 7) % len(edge1)]
        X[q, 1] = [v, np.nextafter(v, np.inf), np.nextafter(v, -np.inf)][(q // 7) % 3]
    V = rng.normal(size=(n, 3)); A = 40.0 * rng.normal(size=(n, 3))
    return X, V, A


def decode(code, V, A, dt):
    """V_hlf from the stored codes: every component a driver returns is the free value V + dt/2 A (0), zero (1) or 0.1 (2)"""
    free = V + (dt / 2.0) * A
    return np.where(code == 0, free, np.where(code == 1, 0.0, 0.1))


def main():
    X, V, A = inputs()
    dt = 0.005
    out = {"X": X, "V": V, "A": A, "times": np.array(TIMES), "dt": np.array(dt)}
    free = V + (dt / 2.0) * A
    for name in DRIVERS:
        vh = np.stack([driver("k1", name, X, V, A, dt, t) for t in TIMES])
        code = np.where(vh == free[None], 0, np.where(vh == 0.0, 1, np.where(vh == 0.1, 2, 3))).astype(np.int8)
        assert code.max() <= 2 and np.array_equal(decode(code, V[None], A[None], dt), vh), name     # the codes hold the reference's output bit for bit
        out["code_" + name] = code
        print("%-11s constrained components at t = 0: x %4d  y %4d  z %4d; driven (0.1): %4d; at t = %g: x %4d" % (
            name, (code[0, :, 0] != 0).sum(), (code[0, :, 1] != 0).sum(), (code[0, :, 2] != 0).sum(), (code[0] == 2).sum(), TIMES[-1], (code[-1, :, 0] != 0).sum()))
    np.savez_compressed(os.path.join(HERE, "drivers.npz"), **out)


if __name__ == "__main__":
    main()
