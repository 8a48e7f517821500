"""oracle/refrun.py -- TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_ref/libulref_{k1,k31,kMD}.so (the unmodified reference UL path, see
oracle/ref_wrap.cpp and oracle/Makefile).  Only tests/, tests/golden/make_golden.py and bench.py's
`--impl reference` / cpu_baseline leg may import this.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FIELDS = {"x": 0, "V": 1, "A": 2, "ea": 3, "cA": 4, "cB": 5, "cC": 6, "fa_u": 7, "X_prev": 8, "V_hlf": 9}
WIDTH = {"x": 3, "V": 3, "A": 3, "ea": 6, "cA": 1, "cB": 1, "cC": 1, "fa_u": 3, "X_prev": 3, "V_hlf": 3}
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def ref_available(variant="k1"):
    return os.path.exists(os.path.join(HERE, "_ref", "libulref_%s.so" % variant))


def adapter_available(variant="k1"):
    """oracle/_ref/libuladapter_*.so: this same driver linked against ul_adapter.cpp + libwcfe.so (the GPU drop-in)."""
    return os.path.exists(os.path.join(HERE, "_ref", "libuladapter_%s.so" % variant))


def _load(variant, kind="ref"):
    lib = C.CDLL(os.path.join(HERE, "_ref", "lib%s_%s.so" % ("ulref" if kind == "ref" else "uladapter", variant)))
    if kind != "ref":
        lib.ulref_create.restype = C.c_void_p
        lib.ulref_create.argtypes = [C.c_int, _dp, C.c_int, _ip, C.c_int, _dp, C.c_int, _ip]
        lib.ulref_destroy.argtypes = [C.c_void_p]
        lib.ulref_srand.argtypes = [C.c_uint]
        lib.ulref_default_kint.restype = C.c_double
        lib.ulref_get.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
        lib.ulref_set.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
        lib.ulref_boundary_flags.argtypes = [C.c_void_p, C.c_int, _ip]
        lib.ulref_step.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]
        lib.ulref_step_short.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
        lib.ulref_step_driver.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int]
        lib.ulref_driver.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, _dp]
        return lib
    lib.ulref_create.restype = C.c_void_p
    lib.ulref_create.argtypes = [C.c_int, _dp, C.c_int, _ip, C.c_int, _dp, C.c_int, _ip]
    lib.ulref_destroy.argtypes = [C.c_void_p]
    lib.ulref_srand.argtypes = [C.c_uint]
    lib.ulref_rand.restype = C.c_int
    lib.ulref_default_kint.restype = C.c_double
    lib.ulref_get.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    lib.ulref_set.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    lib.ulref_onering.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip, C.c_int]
    lib.ulref_boundary_flags.argtypes = [C.c_void_p, C.c_int, _ip]
    lib.ulref_step.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double]
    lib.ulref_step_short.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]
    lib.ulref_step_driver.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int]
    lib.ulref_driver.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, _dp]
    lib.ulref_eval_element.argtypes = [C.c_int, _dp, _dp, _dp, _dp, _dp, C.c_double, _dp, C.c_double, _dp, _dp, _dp]
    lib.ulref_force_field.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp, _dp]
    lib.ulref_write_vtk_cell.argtypes = [C.c_void_p, C.c_double, C.c_char_p]
    lib.ulref_write_vtk_subs.argtypes = [C.c_void_p, C.c_double, C.c_char_p, C.c_int]
    return lib


def _d(a):
    return a.ctypes.data_as(_dp)


DRIVERS = {"explicit": 0, "biaxial": 1, "uniaxial": 2, "offbiaxial": 3, "bulgetest": 4}


def driver(variant, kind, X, V, A, dt, t_tot, lib_kind="ref"):
    """driverVelocities_<kind> of the reference (UL:103-345) on flat arrays; lib_kind "adapter": the drop-in's symbols (GPU)."""
    lib = _load(variant, lib_kind)
    X = np.ascontiguousarray(X, dtype=np.float64); V = np.ascontiguousarray(V, dtype=np.float64); A = np.ascontiguousarray(A, dtype=np.float64)
    out = np.empty_like(X)
    rc = lib.ulref_driver(DRIVERS[kind] if isinstance(kind, str) else kind, len(X), _d(X), _d(V), _d(A), dt, t_tot, _d(out))
    assert rc == 0
    return out


class RefSim:
    """The reference UL step (El Topo off) on one cell + one substrate mesh."""

    def __init__(self, variant, xyz_c, tri_c, xyz_s, tri_s, kind="ref"):
        self.variant = variant
        self.lib = _load(variant, kind)
        xyz_c = np.ascontiguousarray(xyz_c, dtype=np.float64); xyz_s = np.ascontiguousarray(xyz_s, dtype=np.float64)
        tri_c = np.ascontiguousarray(tri_c, dtype=np.int32); tri_s = np.ascontiguousarray(tri_s, dtype=np.int32)
        self.nn = (len(xyz_c), len(xyz_s)); self.ne = (len(tri_c), len(tri_s))
        self.h = self.lib.ulref_create(len(xyz_c), _d(xyz_c), len(tri_c), tri_c.ctypes.data_as(_ip),
                                       len(xyz_s), _d(xyz_s), len(tri_s), tri_s.ctypes.data_as(_ip))
        self.kint = self.lib.ulref_default_kint()          # what main.cpp passes to updateFA
        self.kint_force = 1000.0 if variant == "k1" else 31000.0   # hard-coded in updateForces_Explicit

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.ulref_destroy(self.h); self.h = None

    def srand(self, seed):
        self.lib.ulref_srand(seed)

    def get(self, surface, field):
        n = self.nn[surface]; w = WIDTH[field]
        out = np.empty((n, w) if w > 1 else (n,), dtype=np.float64)
        rc = self.lib.ulref_get(self.h, surface, FIELDS[field], _d(out)); assert rc == 0
        return out

    def set(self, surface, field, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64); assert arr.size == self.nn[surface] * WIDTH[field]
        rc = self.lib.ulref_set(self.h, surface, FIELDS[field], _d(arr)); assert rc == 0

    def step(self, dt, t_curr, subs_scale=0.0):
        self.lib.ulref_step(self.h, dt, t_curr, self.kint, subs_scale)

    def step_short(self, dt, t_curr, frac_cell, frac_subs=1.0):
        """one step in which the host integrator accepts only a fraction of the requested dt (main.cpp:244-335)"""
        self.lib.ulref_step_short(self.h, dt, t_curr, self.kint, frac_cell, frac_subs)

    def step_driver(self, dt, t_curr, subs_kind):
        """one step with the substrate's driver swapped for an alternative of main.cpp:291-295 (2 Uniaxial, 3 OffBiaxial, 4 BulgeTest)"""
        self.lib.ulref_step_driver(self.h, dt, t_curr, self.kint, subs_kind)

    def write_vtk(self, surface, time, path, fa_source=0):
        """the reference's own writer (cellBP/iomesh.cpp)"""
        if surface == 0:
            self.lib.ulref_write_vtk_cell(self.h, time, path.encode())
        else:
            self.lib.ulref_write_vtk_subs(self.h, time, path.encode(), fa_source)

    def boundary_flags(self, surface):
        out = np.empty(self.nn[surface], dtype=np.int32)
        self.lib.ulref_boundary_flags(self.h, surface, out.ctypes.data_as(_ip)); return out

    def onering(self, surface, v):
        out = np.empty(64, dtype=np.int32)
        n = self.lib.ulref_onering(self.h, surface, v, out.ctypes.data_as(_ip), 64); return out[:n].copy()

    def force_field(self, which, dt, t_curr):
        names = {"Res_cell": 0, "F_FA": 1, "F_AP": 2, "F_bend": 3, "F_area": 4, "Res_subs": 5, "F_FA_subs": 6,
                 "Res_Adif": 7, "Res_Cdif": 8, "F_pressure": 9, "F_sphere": 10}      # 7-10: dormant in the reference (calls commented out)
        w = names[which]; surf = 1 if w in (5, 6) else 0; n = self.nn[surf]
        out = np.zeros((n, 3)); mm = np.zeros(n); rea = np.zeros((n, 6)); aux = np.zeros(4)
        rc = self.lib.ulref_force_field(self.h, w, dt, t_curr, self.kint_force, _d(out), _d(mm), _d(rea), _d(aux)); assert rc == 0
        if w in (7, 8):
            return {"F": out.ravel()[:n].copy(), "area": aux}                       # one scalar per node
        return {"F": out, "MM": mm, "Res_ea": rea, "area": aux}


def ref_eval_element(variant, mat, X, x, ea, cA, cC, dt, centroid, time):
    lib = _load(variant)
    X = np.ascontiguousarray(X, dtype=np.float64); x = np.ascontiguousarray(x, dtype=np.float64)
    ea = np.ascontiguousarray(ea, dtype=np.float64); cA = np.ascontiguousarray(cA, dtype=np.float64)
    cC = np.ascontiguousarray(cC, dtype=np.float64); cen = np.ascontiguousarray(centroid, dtype=np.float64)
    Re = np.zeros(9); Rea = np.zeros(18); MMe = np.zeros(3)
    lib.ulref_eval_element(mat, _d(X), _d(x), _d(ea), _d(cA), _d(cC), dt, _d(cen), time, _d(Re), _d(Rea), _d(MMe))
    return Re, Rea, MMe
