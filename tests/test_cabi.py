"""C-ABI library: loads, exports every symbol include/wcfe.h declares, refuses to compute without a GPU.  CPU only."""
import ctypes as C
import os
import re
import numpy as np
import pytest
import wholecellfe_b200 as W
from wholecellfe_b200 import _lib as B
from wholecellfe_b200.build import build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build()
    return B.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "wcfe.h")).read()
    declared = sorted(set(re.findall(r"\b(wcfe_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found in include/wcfe.h"
    assert sorted(B.SYMBOLS) == declared
    for s in declared:
        assert hasattr(lib, s), s


def test_default_params_are_the_reference_literals(lib):
    from oracle.oracle import Params as OP, lib as olib
    for name, vid in B.VARIANTS.items():
        p = W.default_params(name)
        o = OP(); olib().ulo_default_params(C.byref(o), vid)
        for f, _ in OP._fields_:
            assert getattr(p, f) == getattr(o, f), (name, f)
        assert p.c_A_init == 100 and p.thetag_dot == 0.0
        assert list(p.md_px)[5] == 17.001e-03 and list(p.md_k)[3] == -7908.39043
    assert W.default_params("k1").kint_force == 1000 and W.default_params("kMD").kint_bond == -1 and W.default_params("kMD").law == 1


def test_bad_arguments_do_not_need_a_gpu(lib):
    p = B.Params()
    assert lib.wcfe_default_params(C.byref(p), 7) == B.ERR_ARG
    assert lib.wcfe_default_params(None, 0) == B.ERR_ARG
    assert lib.wcfe_destroy(None) == B.ERR_ARG
    assert lib.wcfe_last_error(None) == b"null context"
    assert lib.wcfe_kernel_launches(None) == 0


def test_no_cpu_fallback(lib):
    """Without a CUDA device the product must fail loudly, not compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    p = W.default_params("k1")
    assert lib.wcfe_create(C.byref(h), 0, C.byref(p)) == B.ERR_CUDA
    with pytest.raises(W.WcfeError):
        W.ULSim("k1")


def test_product_does_not_import_the_oracle():
    import subprocess, sys
    code = "import sys; import wholecellfe_b200, wholecellfe_b200.meshgen, wholecellfe_b200.meshio; print([m for m in sys.modules if m.startswith('oracle')])"
    out = subprocess.check_output([sys.executable, "-c", code], cwd=ROOT).decode().strip()
    assert out == "[]"
    for dirpath, _, files in os.walk(os.path.join(ROOT, "wholecellfe_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                for needle in ("ul_oracle.h", "libuloracle", "from oracle", "import oracle"):
                    assert needle not in txt, (f, needle)


def test_environment_switches_are_documented():
    """every WCFE_* variable the library, the adapter or the Python loader reads appears in INTEGRATION.md's table, and vice versa"""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    read = set()
    for f in glob.glob(os.path.join(root, "wholecellfe_b200", "csrc", "*.cu")) + glob.glob(os.path.join(root, "wholecellfe_b200", "csrc", "*.cpp")):
        read |= set(re.findall(r'getenv\("(WCFE_[A-Z0-9_]+)"', open(f).read()))
    read |= set(re.findall(r'environ\.get\("(WCFE_[A-Z0-9_]+)"', open(os.path.join(root, "wholecellfe_b200", "_lib.py")).read()))
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    table = doc[doc.index("Environment switches"):]
    documented = set(re.findall(r"`(WCFE_[A-Z0-9_]+)", table))
    assert read == documented, (sorted(read - documented), sorted(documented - read))


def test_header_constants_match_the_python_tables():
    """the enumerations of include/wcfe.h and the tables the ctypes binding / the oracle bindings use are the same numbers"""
    from oracle import oracle as O, refrun as R
    hdr = open(os.path.join(ROOT, "include", "wcfe.h")).read()
    defs = {k: int(v) for k, v in re.findall(r"#define\s+(WCFE_[A-Z0-9_]+)\s+\(?(-?\d+)\)?", hdr)}
    for name, v in B.DRIVERS.items():
        assert defs["WCFE_DRV_" + name.upper()] == v == O.DRIVERS[name] == R.DRIVERS[name], name
    for name, v in B.DORMANT.items():
        assert defs["WCFE_DORMANT_" + name.upper().replace("F_", "")] == v, name
    for name, v in B.FIELDS.items():
        key = {"x": "X", "V": "V", "A": "A", "ea": "EA", "cA": "CA", "cB": "CB", "cC": "CC", "fa_u": "FAU", "X_prev": "XPREV"}[name]
        assert defs["WCFE_F_" + key] == v, name
    assert defs["WCFE_ERR_ARG"] == B.ERR_ARG and defs["WCFE_ERR_CUDA"] == B.ERR_CUDA and defs["WCFE_CELL"] == B.CELL and defs["WCFE_SUBS"] == B.SUBS
