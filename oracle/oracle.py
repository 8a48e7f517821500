"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/libuloracle.so (the plain-C restatement, oracle/ul_oracle.c).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libuloracle.so")
FIELDS = {"x": 0, "V": 1, "A": 2, "ea": 3, "cA": 4, "cB": 5, "cC": 6, "fa_u": 7, "X_prev": 8, "V_hlf": 9}
WIDTH = {"x": 3, "V": 3, "A": 3, "ea": 6, "cA": 1, "cB": 1, "cC": 1, "fa_u": 3, "X_prev": 3, "V_hlf": 3}
DISCRETE = {"loc_bond": 0, "loc_force_cnt": 1, "loc_force_first": 2, "ap_fire": 3, "fa_reset": 4, "cc_clamp": 5}
VARIANTS = {"k1": 0, "k31": 1, "kMD": 2}
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class Params(C.Structure):
    _fields_ = [("law", C.c_int), ("kint_force", C.c_double), ("kint_bond", C.c_double), ("subs_scale", C.c_double),
                ("mu_cell", C.c_double), ("mu_subs", C.c_double),
                ("t_act_rate", C.c_double), ("t_act_max", C.c_double), ("t_act_t1", C.c_double),
                ("thick", C.c_double), ("kappa", C.c_double),
                ("damp", C.c_double), ("rho", C.c_double), ("drag", C.c_double),
                ("rho_i_max", C.c_double), ("k_bend", C.c_double), ("k_area", C.c_double), ("area_0", C.c_double),
                ("ap_rate", C.c_double), ("ap_quantum", C.c_double), ("ap_density", C.c_double),
                ("k_on", C.c_double), ("ka", C.c_double), ("kb", C.c_double), ("Fa", C.c_double), ("Fb", C.c_double),
                ("pin_limit", C.c_double), ("t_i", C.c_double), ("t_f", C.c_double)]


def build(force=False):
    """Compile oracle/libuloracle.so (gcc) if missing or stale."""
    src = os.path.join(HERE, "ul_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(src), os.path.getmtime(src[:-1] + "h")):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.ulo_create.restype = C.c_void_p
        L.ulo_create.argtypes = [C.c_int, C.c_int, _dp, C.c_int, _ip, C.c_int, _dp, C.c_int, _ip]
        L.ulo_destroy.argtypes = [C.c_void_p]
        L.ulo_params_of.restype = C.POINTER(Params); L.ulo_params_of.argtypes = [C.c_void_p]
        L.ulo_set_use_grid.argtypes = [C.c_void_p, C.c_int]
        L.ulo_seed.argtypes = [C.c_void_p, C.c_uint]
        L.ulo_rng_discard.argtypes = [C.c_void_p, C.c_long]
        L.ulo_get.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
        L.ulo_set.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
        L.ulo_get_discrete.argtypes = [C.c_void_p, C.c_int, _ip]
        L.ulo_get_diag.argtypes = [C.c_void_p, _dp]
        L.ulo_step.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.ulo_force_field.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, _dp, _dp, _dp, _dp]
        L.ulo_step_force_field.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp]
        L.ulo_eval_element.argtypes = [C.POINTER(Params), C.c_int, _dp, _dp, _dp, C.c_double, _dp, C.c_double, _dp, _dp, _dp]
        L.ulo_default_params.argtypes = [C.POINTER(Params), C.c_int]
        L.ulo_srand.argtypes = [C.c_void_p, C.c_uint]
        L.ulo_rand.argtypes = [C.c_void_p]; L.ulo_rand.restype = C.c_int
        L.ulo_driver_kind.argtypes = [C.POINTER(Params), C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, _dp]
        L.ulo_boundary_neighbours.argtypes = [C.c_void_p, C.c_int, _dp, _dp]
        L.ulo_mesh_build.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def glibc_rand_stream(seed, n):
    """First n outputs of glibc rand() after srand(seed), from the oracle's restatement."""
    L = lib()
    buf = C.create_string_buffer(4 * 31 + 8)
    L.ulo_srand(buf, seed)
    return np.array([L.ulo_rand(buf) for _ in range(n)], dtype=np.int64)


DRIVERS = {"explicit": 0, "biaxial": 1, "uniaxial": 2, "offbiaxial": 3, "bulgetest": 4}


def driver(variant, kind, X, V, A, dt, t_tot):
    """The restatement of driverVelocities_<kind> (UL:103-345) on flat AoS arrays."""
    L = lib()
    p = Params(); L.ulo_default_params(C.byref(p), VARIANTS[variant])
    X = np.ascontiguousarray(X, dtype=np.float64); V = np.ascontiguousarray(V, dtype=np.float64); A = np.ascontiguousarray(A, dtype=np.float64)
    out = np.empty_like(X)
    assert L.ulo_driver_kind(C.byref(p), DRIVERS[kind] if isinstance(kind, str) else kind, len(X), _d(X), _d(V), _d(A), dt, t_tot, _d(out)) == 0
    return out


class OracleSim:
    def __init__(self, variant, xyz_c, tri_c, xyz_s, tri_s, use_grid=False):
        self.L = lib()
        xyz_c = np.ascontiguousarray(xyz_c, dtype=np.float64); xyz_s = np.ascontiguousarray(xyz_s, dtype=np.float64)
        tri_c = np.ascontiguousarray(tri_c, dtype=np.int32); tri_s = np.ascontiguousarray(tri_s, dtype=np.int32)
        self.nn = (len(xyz_c), len(xyz_s)); self.ne = (len(tri_c), len(tri_s))
        self.h = self.L.ulo_create(VARIANTS[variant], len(xyz_c), _d(xyz_c), len(tri_c), tri_c.ctypes.data_as(_ip),
                                   len(xyz_s), _d(xyz_s), len(tri_s), tri_s.ctypes.data_as(_ip))
        self.params = self.L.ulo_params_of(self.h).contents
        self.L.ulo_set_use_grid(self.h, 1 if use_grid else 0)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ulo_destroy(self.h); self.h = None

    def srand(self, seed):
        self.L.ulo_seed(self.h, seed)

    def rng_discard(self, n):
        self.L.ulo_rng_discard(self.h, int(n))

    def get(self, surface, field):
        n = self.nn[surface]; w = WIDTH[field]
        out = np.empty((n, w) if w > 1 else (n,), dtype=np.float64)
        assert self.L.ulo_get(self.h, surface, FIELDS[field], _d(out)) == 0
        return out

    def set(self, surface, field, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64); assert arr.size == self.nn[surface] * WIDTH[field]
        assert self.L.ulo_set(self.h, surface, FIELDS[field], _d(arr)) == 0

    def boundary_neighbours(self, ni):
        """(j1, j2): positions of the two boundary neighbours of cell vertex ni (UL:1479-1507), or None for an interior vertex"""
        j1 = np.zeros(3); j2 = np.zeros(3)
        return (j1, j2) if self.L.ulo_boundary_neighbours(self.h, int(ni), _d(j1), _d(j2)) else None

    def discrete(self, which):
        out = np.empty(self.nn[0], dtype=np.int32)
        assert self.L.ulo_get_discrete(self.h, DISCRETE[which], out.ctypes.data_as(_ip)) == 0
        return out

    def diag(self):
        out = np.zeros(8); self.L.ulo_get_diag(self.h, _d(out)); return out

    def step(self, dt, t_curr):
        self.L.ulo_step(self.h, dt, t_curr)

    def force_field(self, which, dt, t_curr):
        names = {"Res_cell": 0, "F_FA": 1, "F_AP": 2, "F_bend": 3, "F_area": 4, "Res_subs": 5, "F_FA_subs": 6,
                 "Res_Adif": 7, "Res_Cdif": 8, "F_pressure": 9, "F_sphere": 10}      # 7-10: dormant in the reference (calls commented out)
        w = names[which]; surf = 1 if w in (5, 6) else 0; n = self.nn[surf]
        out = np.zeros((n, 3)); mm = np.zeros(n); rea = np.zeros((n, 6)); aux = np.zeros(4)
        assert self.L.ulo_force_field(self.h, w, dt, t_curr, _d(out), _d(mm), _d(rea), _d(aux)) == 0
        if w in (7, 8):
            return {"F": out.ravel()[:n].copy(), "area": aux}                       # one scalar per node
        return {"F": out, "MM": mm, "Res_ea": rea, "area": aux}


    def step_force_field(self, which):
        """Force field as it was inside the last step() (same names as force_field)."""
        names = {"Res_cell": 0, "F_FA": 1, "F_AP": 2, "F_bend": 3, "F_area": 4, "Res_subs": 5, "F_FA_subs": 6}
        w = names[which]; surf = 1 if w >= 5 else 0; n = self.nn[surf]
        out = np.zeros((n, 3)); mm = np.zeros(n); rea = np.zeros((n, 6))
        assert self.L.ulo_step_force_field(self.h, w, _d(out), _d(mm), _d(rea)) == 0
        return {"F": out, "MM": mm, "Res_ea": rea}


def eval_element(variant, mat, X, x, ea, dt, centroid, time):
    L = lib(); p = Params(); L.ulo_default_params(C.byref(p), VARIANTS[variant])
    X = np.ascontiguousarray(X, dtype=np.float64); x = np.ascontiguousarray(x, dtype=np.float64)
    ea = np.ascontiguousarray(ea, dtype=np.float64); cen = np.ascontiguousarray(centroid, dtype=np.float64)
    Re = np.zeros(9); Rea = np.zeros(18); MMe = np.zeros(3)
    L.ulo_eval_element(C.byref(p), mat, _d(X), _d(x), _d(ea), dt, _d(cen), time, _d(Re), _d(Rea), _d(MMe))
    return Re, Rea, MMe
