"""wholecellfe_b200 -- the explicit updated-Lagrangian step of wholeCellFE's cellBP as hand-written sm_100a CUDA.

Only what the hot path needs: csrc/ (kernels + the C ABI of include/wcfe.h), the ctypes binding, the host-side
mirror of the reference's call sequence (ulstep.ULSim), mesh I/O in the reference's format and synthetic mesh
generators.  Importing this package does not import the oracle.
"""
from ._lib import load, LIB_PATH, CELL, SUBS   # noqa: F401
from .ulstep import ULSim, WcfeError, default_params   # noqa: F401
