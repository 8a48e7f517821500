import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def meshes():
    d = np.load(os.path.join(GOLDEN, "example_meshes.npz"))
    return d["cell_xyz"], d["cell_tri"], d["subs_xyz"], d["subs_tri"]


@pytest.fixture(scope="session")
def traj():
    return np.load(os.path.join(GOLDEN, "trajectories.npz"))


@pytest.fixture(scope="session")
def elements():
    return np.load(os.path.join(GOLDEN, "elements.npz"))


def rel_err(a, b, scale=None):
    """max |a-b| relative to the field's own scale (never element-wise: SURVEY.md quirk q7)."""
    a = np.asarray(a); b = np.asarray(b)
    s = max(np.abs(b).max(), 1e-300) if scale is None else scale
    return float(np.abs(a - b).max() / s)


@pytest.fixture(scope="session")
def drivers():
    """tests/golden/drivers.npz (make_golden_drivers.py): inputs and, per driver and time, the reference's V_hlf as codes"""
    d = np.load(os.path.join(GOLDEN, "drivers.npz"))
    free = d["V"] + (float(d["dt"]) / 2.0) * d["A"]

    def expect(name, it):
        code = d["code_" + name][it]
        return np.where(code == 0, free, np.where(code == 1, 0.0, 0.1))
    return d, expect


@pytest.fixture(scope="session")
def dormant():
    """tests/golden/dormant.npz (make_golden_dormant.py): a bent cell state and the reference's dormant fields on it"""
    return np.load(os.path.join(GOLDEN, "dormant.npz"))
