"""Host-side mirror of the reference's UL step interface, on top of the C ABI (include/wcfe.h).

`ULSim` owns one cell + one substrate on one GPU and exposes the reference's call sequence
(cellBP/main.cpp:168-418): predict (driverVelocities_* + inline predictor), commit (positions the host
integrator accepted), update (updateFA + the two updateForces_Explicit overloads), plus the El-Topo-free
`step_n` fast path.  Arrays cross this interface in the reference's own array-of-structs layout.
"""
import ctypes as C
import numpy as np
from . import _lib as B


class WcfeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("wcfe error %d: %s" % (code, msg))
        self.code = code


def default_params(variant="k1"):
    p = B.Params()
    L = B.load()
    rc = L.wcfe_default_params(C.byref(p), B.VARIANTS[variant])
    if rc:
        raise WcfeError(rc, "default_params")
    return p


def _d(a):
    return a.ctypes.data_as(B._dp)


class ULSim:
    def __init__(self, variant="k1", device=0, params=None, stream=None, batch=None):
        """batch: a list of Params -> one batched context holding len(batch) independent replicas (wcfe_create_batch)."""
        self.L = B.load()
        self.variant = variant
        self.h = C.c_void_p()
        if batch is not None:
            arr = (B.Params * len(batch))(*batch)
            self.params = arr[0]
            self.replicas = len(batch)
            rc = self.L.wcfe_create_batch(C.byref(self.h), device, arr, len(batch))
        else:
            self.params = params if params is not None else default_params(variant)
            self.replicas = 1
            rc = self.L.wcfe_create(C.byref(self.h), device, C.byref(self.params))
        if rc:
            self.h = None
            why = self.L.wcfe_last_error(None).decode()
            raise WcfeError(rc, "wcfe_create(device %d) failed: %s%s" % (device, why, " -- the product has no CPU fallback" if rc == B.ERR_CUDA else ""))
        if stream is not None:
            # torch reports the legacy default stream as handle 0, which the C ABI reads as "use the context's own stream";
            # cudaStreamLegacy (0x1) names the default stream explicitly
            self._ck(self.L.wcfe_set_stream(self.h, C.c_void_p(stream if stream != 0 else 1)))
        self.nn = [0, 0]
        self.ne = [0, 0]

    def close(self):
        if getattr(self, "h", None):
            self.L.wcfe_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _ck(self, rc):
        if rc:
            raise WcfeError(rc, self.L.wcfe_last_error(self.h).decode())

    def select(self, replica):
        """Batched context: the replica that set/get/params/discrete/diagnostics address from now on."""
        self._ck(self.L.wcfe_select_replica(self.h, int(replica)))

    def set_replica_seed(self, replica, seed):
        self._ck(self.L.wcfe_set_replica_seed(self.h, int(replica), int(seed)))

    def set_params(self, params):
        self._ck(self.L.wcfe_set_params(self.h, C.byref(params)))
        self.params = params

    def upload_mesh(self, surface, xyz, tri):
        xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        tri = np.ascontiguousarray(tri, dtype=np.int32)
        if xyz.ndim != 2 or xyz.shape[1] != 3 or tri.ndim != 2 or tri.shape[1] != 3:
            raise WcfeError(B.ERR_ARG, "xyz must be [N,3] and tri [M,3]")
        self._ck(self.L.wcfe_upload_mesh(self.h, surface, len(xyz), _d(xyz), len(tri), tri.ctypes.data_as(B._ip)))
        self.nn[surface] = len(xyz)
        self.ne[surface] = len(tri)

    def set(self, surface, field, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        if arr.size != self.nn[surface] * B.WIDTH[field]:
            raise WcfeError(B.ERR_ARG, "field %s: wrong size" % field)
        self._ck(self.L.wcfe_upload_state(self.h, surface, B.FIELDS[field], _d(arr)))

    def get(self, surface, field):
        n = self.nn[surface]
        w = B.WIDTH[field]
        out = np.empty((n, w) if w > 1 else (n,), dtype=np.float64)
        self._ck(self.L.wcfe_download_state(self.h, surface, B.FIELDS[field], _d(out)))
        return out

    def set_rng(self, mode, seed):
        self._ck(self.L.wcfe_set_rng(self.h, mode, seed))

    def srand(self, seed):
        self.set_rng(B.RNG_GLIBC_REPLAY, seed)

    def rng_skip(self, n_draws):
        self._ck(self.L.wcfe_rng_skip(self.h, int(n_draws)))

    def set_profile(self, on):
        self._ck(self.L.wcfe_set_profile(self.h, 1 if on else 0))

    def kernel_ms(self):
        ms = np.zeros(8)
        cnt = (C.c_longlong * 8)()
        self._ck(self.L.wcfe_get_kernel_ms(self.h, _d(ms), cnt))
        return ms, np.array(list(cnt))

    def set_force_general(self, on):
        self._ck(self.L.wcfe_set_force_general(self.h, 1 if on else 0))

    def path_info(self):
        out = np.zeros(8)
        self._ck(self.L.wcfe_get_path_info(self.h, _d(out)))
        return {"planar": bool(out[0]), "cell_patch_elem_ratio": out[1], "subs_patch_elem_ratio": out[2], "cell_max_halo": int(out[3]),
                "cell_max_elem": int(out[4]), "subs_max_halo": int(out[5]),
                "subs_kernel": {2: "k_subs_ring", 0: "k_subs_patch"}.get(int(out[6]), "?"), "bucket_cap": int(out[7])}

    def set_debug(self, on):
        self._ck(self.L.wcfe_set_debug(self.h, 1 if on else 0))

    def predict(self, surface, dt, t_curr, download=True):
        out = np.empty((self.nn[surface], 3)) if download else None
        self._ck(self.L.wcfe_predict(self.h, surface, dt, t_curr, _d(out) if download else None))
        return out

    def commit(self, surface, xyz=None):
        if xyz is not None:
            xyz = np.ascontiguousarray(xyz, dtype=np.float64)
        self._ck(self.L.wcfe_commit(self.h, surface, _d(xyz) if xyz is not None else None))

    def update(self, dt_req, dt_real, t_curr):
        self._ck(self.L.wcfe_update(self.h, dt_req, dt_real, t_curr))

    def update_cell(self, dt_req, dt_real, t_curr):
        """updateFA + updateForces_Explicit(cell) (main.cpp:354, :360)"""
        self._ck(self.L.wcfe_update_cell(self.h, dt_req, dt_real, t_curr))

    def update_subs(self, dt_req, dt_real, t_curr):
        """updateForces_Explicit(substrate) (main.cpp:361); follows update_cell of the same step"""
        self._ck(self.L.wcfe_update_subs(self.h, dt_req, dt_real, t_curr))

    def mark_positions_committed(self):
        self._ck(self.L.wcfe_mark_positions_committed(self.h))

    def driver_velocities(self, surface, X, V, A, dt, t_curr, kind=None):
        """driverVelocities_* (UL:103-345) on host-owned AoS arrays, computed on the device; kind: one of B.DRIVERS (default: the driver
        main.cpp calls for that surface).  The next update of `surface` uses this V_hlf (see include/wcfe.h)."""
        X = np.ascontiguousarray(X, dtype=np.float64); V = np.ascontiguousarray(V, dtype=np.float64); A = np.ascontiguousarray(A, dtype=np.float64)
        out = np.empty_like(X)
        k = surface if kind is None else (B.DRIVERS[kind] if isinstance(kind, str) else int(kind))
        self._ck(self.L.wcfe_driver_velocities_kind(self.h, surface, k, len(X), _d(X), _d(V), _d(A), dt, t_curr, _d(out)))
        return out

    def step(self, dt, t_curr):
        """One El-Topo-free step; same signature as the oracle's step()."""
        t = C.c_double(t_curr)
        self._ck(self.L.wcfe_step_n(self.h, 1, dt, C.byref(t)))
        return t.value

    def step_n(self, n, dt, t0):
        t = C.c_double(t0)
        self._ck(self.L.wcfe_step_n(self.h, n, dt, C.byref(t)))
        return t.value

    def step_n_async(self, n, dt, t0):
        """Queue n steps and return immediately (sync() waits and checks for device-side errors)."""
        t = C.c_double(t0)
        self._ck(self.L.wcfe_step_n_async(self.h, n, dt, C.byref(t)))
        return t.value

    def step_n_roundtrip(self, n, dt, t0, both_surfaces=False):
        """n steps with the positions handed to the host and back every step; both_surfaces: also the (frozen) substrate's, serial."""
        self._ck(self.L.wcfe_set_roundtrip_mode(self.h, 1 if both_surfaces else 0))
        t = C.c_double(t0)
        a = C.c_longlong()
        b = C.c_longlong()
        self._ck(self.L.wcfe_step_n_roundtrip(self.h, n, dt, C.byref(t), C.byref(a), C.byref(b)))
        return t.value, a.value, b.value

    def diagnostics(self):
        out = np.zeros(8)
        self._ck(self.L.wcfe_get_diagnostics(self.h, _d(out)))
        return out

    def discrete(self, which):
        out = np.empty(self.nn[0], dtype=np.int32)
        self._ck(self.L.wcfe_get_discrete(self.h, B.DISCRETE[which], out.ctypes.data_as(B._ip)))
        return out

    def boundary_flags(self, surface):
        out = np.empty(self.nn[surface], dtype=np.int32)
        self._ck(self.L.wcfe_get_boundary_flags(self.h, surface, out.ctypes.data_as(B._ip)))
        return out

    def force_field(self, which):
        w = B.FORCE_FIELDS[which]
        surf = 1 if w >= 5 else 0
        n = self.nn[surf]
        out = np.zeros((n, 3))
        mm = np.zeros(n)
        rea = np.zeros((n, 6))
        self._ck(self.L.wcfe_force_field(self.h, w, 0.0, 0.0, _d(out), _d(mm), _d(rea)))
        return {"F": out, "MM": mm, "Res_ea": rea}

    def dormant_field(self, which, dt, t_curr):
        """Res_Adif / Res_Cdif ([nn]) or F_pressure / F_sphere ([nn, 3]) of the cell on the current device state: the reference's
        dormant physics (include/wcfe.h wcfe_dormant_field)."""
        w = B.DORMANT[which]
        out = np.empty(self.nn[B.CELL] if w < 2 else (self.nn[B.CELL], 3))
        self._ck(self.L.wcfe_dormant_field(self.h, w, dt, t_curr, _d(out)))
        return out

    def device_ptr(self, surface, field):
        """(address, pitch) of a SoA state field on the device: component q of node i lives at address + 8*(q*pitch + i)."""
        ptr = C.c_void_p()
        pitch = C.c_longlong()
        self._ck(self.L.wcfe_device_ptr(self.h, surface, B.FIELDS[field], C.byref(ptr), C.byref(pitch)))
        return ptr.value, pitch.value

    # ---- partitioned substrate: peer-memory halo exchange fused into the substrate kernel (include/wcfe.h) ----
    def halo_info(self):
        """bytes of this context's wcfe_halo_info (addresses + CUDA IPC handles of its strain buffers and arrival counters)"""
        info = B.HaloInfo()
        self._ck(self.L.wcfe_halo_info_get(self.h, C.byref(info)))
        return bytes(info)

    def halo_connect(self, peer_infos, my_slot_at_peer, rows_from_peer, export_node, export_peer, export_row, ghost_node):
        n = len(peer_infos)
        arr = (B.HaloInfo * max(1, n))()
        for q, b in enumerate(peer_infos):
            C.memmove(C.byref(arr[q]), b, C.sizeof(B.HaloInfo))
        ia = lambda v: np.ascontiguousarray(v, dtype=np.int32)
        slot, rows, en, ep, er, gn = ia(my_slot_at_peer), ia(rows_from_peer), ia(export_node), ia(export_peer), ia(export_row), ia(ghost_node)
        ip = lambda a: a.ctypes.data_as(B._ip)
        self._ck(self.L.wcfe_halo_connect(self.h, n, arr, ip(slot), ip(rows), len(en), ip(en), ip(ep), ip(er), len(gn), ip(gn)))

    def halo_disconnect(self):
        self._ck(self.L.wcfe_halo_disconnect(self.h))

    def cell_halo_info(self):
        """bytes of this context's wcfe_cell_halo_info (addresses + CUDA IPC handles of its cell buffers and epoch counters)"""
        info = B.CellHaloInfo()
        self._ck(self.L.wcfe_cell_halo_info_get(self.h, C.byref(info)))
        return bytes(info)

    def cell_halo_connect(self, peer_infos, my_slot_at_peer, patch_list, patch_owned, export_node, export_peer, import_node, margin):
        n = len(peer_infos)
        arr = (B.CellHaloInfo * max(1, n))()
        for q, b in enumerate(peer_infos):
            C.memmove(C.byref(arr[q]), b, C.sizeof(B.CellHaloInfo))
        ia = lambda v: np.ascontiguousarray(v, dtype=np.int32)
        slot, pl, po, en, ep, im = ia(my_slot_at_peer), ia(patch_list), ia(patch_owned), ia(export_node), ia(export_peer), ia(import_node)
        ip = lambda a: a.ctypes.data_as(B._ip)
        self._ck(self.L.wcfe_cell_halo_connect(self.h, n, arr, ip(slot), len(pl), ip(pl), ip(po), len(en), ip(en), ip(ep), len(im), ip(im), float(margin)))

    def cell_halo_disconnect(self):
        self._ck(self.L.wcfe_cell_halo_disconnect(self.h))

    def grid_refreshes(self):
        """moving substrate: rebuilds of the point-location grid so far"""
        return int(self.L.wcfe_grid_refreshes(self.h))

    def kernel_launches(self):
        return int(self.L.wcfe_kernel_launches(self.h))

    def last_step_ms(self):
        return float(self.L.wcfe_last_step_ms(self.h))

    def sync(self):
        self._ck(self.L.wcfe_sync(self.h))
