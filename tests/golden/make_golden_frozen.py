"""tests/golden/make_golden_frozen.py -- known-answer vectors for the FROZEN-substrate element (x == X, strain != 0).

Run in the build container only (needs /root/reference and oracle/_ref, built by `make -C oracle ref`):
    python tests/golden/make_golden_frozen.py
Every number comes from the reference's own substrate overload of evalElementRe (UL_growth_explicit.cpp:988-1174,
mat == 2) through oracle/_ref/libulref_k1.so.  The frozen-substrate closed forms of the CUDA path
(wcfe_math.cuh elem_eval_frozen, frozen_fan_entry) and the C restatement are checked AGAINST this file.
A substrate that starts strain-free stays strain-free (main.cpp:299 freezes it), so none of the trajectory fixtures
exercises this arithmetic at a non-trivial input: these vectors do.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.refrun import ref_eval_element   # noqa: E402


def main():
    rng = np.random.default_rng(20261018)
    n = 128
    base = np.array([[0, 0, 0], [1.0, 0, 0], [0.4, 0.9, 0]])
    X = base[None] * rng.uniform(0.05, 6.0, size=(n, 1, 1)) + rng.uniform(-50, 50, size=(n, 1, 3))
    X += 0.15 * rng.normal(size=(n, 3, 3)) * np.abs(X[:, 1:2, :] - X[:, 0:1, :]).max(axis=2, keepdims=True)
    X[:, :, 2] = 0.0
    flip = rng.random(n) < 0.5                      # both orientations (the union-jack substrate has both)
    X[flip] = X[flip][:, [0, 2, 1], :]
    mag = 10.0 ** rng.uniform(-6, np.log10(5e-2), size=(n, 1, 1))
    ea = mag * rng.normal(size=(n, 3, 6))
    ea[:, :, [2, 4, 5]] = 0.0                       # planar strain: xx, yy, xy only
    time = rng.uniform(0, 9, size=n)
    Re = np.zeros((n, 9)); Rea = np.zeros((n, 18)); MMe = np.zeros((n, 3))
    for i in range(n):
        Re[i], Rea[i], MMe[i] = ref_eval_element("k1", 2, X[i].ravel(), X[i].ravel(), ea[i].ravel(), np.zeros(3), np.zeros(3),
                                                 0.005, np.zeros(3), float(time[i]))
    np.savez_compressed(os.path.join(HERE, "elements_frozen.npz"), X=X, ea=ea, time=time, Re=Re, Re_ea=Rea, MMe=MMe)
    print("wrote elements_frozen.npz: max |Re| %.3e, min |Re| %.3e" % (np.abs(Re).max(), np.abs(Re).max(axis=1).min()))


if __name__ == "__main__":
    main()
