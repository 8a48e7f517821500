"""ctypes binding of libwcfe.so (include/wcfe.h).  Fails loudly when the CUDA library is missing."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WCFE_LIB") or os.path.join(HERE, "libwcfe.so")   # WCFE_LIB: a tuning variant built by build.py
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)

OK, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_CAPACITY, ERR_NUMERIC, ERR_PEER = 0, -1, -2, -3, -4, -5, -6
HALO_MAX_PEERS = 8
CELL, SUBS = 0, 1
VARIANTS = {"k1": 0, "k31": 1, "kMD": 2}
FIELDS = {"x": 0, "V": 1, "A": 2, "ea": 3, "cA": 4, "cB": 5, "cC": 6, "fa_u": 7, "X_prev": 8}
WIDTH = {"x": 3, "V": 3, "A": 3, "ea": 6, "cA": 1, "cB": 1, "cC": 1, "fa_u": 3, "X_prev": 3}
DISCRETE = {"loc_bond": 0, "loc_force_cnt": 1, "loc_force_first": 2, "ap_fire": 3, "fa_reset": 4, "cc_clamp": 5}
FORCE_FIELDS = {"Res_cell": 0, "F_FA": 1, "F_AP": 2, "F_bend": 3, "F_area": 4, "Res_subs": 5, "F_FA_subs": 6}
RNG_GLIBC_REPLAY, RNG_PHILOX = 0, 1
DORMANT = {"Res_Adif": 0, "Res_Cdif": 1, "F_pressure": 2, "F_sphere": 3}      # WCFE_DORMANT_*
DRIVERS = {"explicit": 0, "biaxial": 1, "uniaxial": 2, "offbiaxial": 3, "bulgetest": 4}      # WCFE_DRV_*

# every symbol include/wcfe.h declares
SYMBOLS = ["wcfe_default_params", "wcfe_create", "wcfe_destroy", "wcfe_set_params", "wcfe_set_stream", "wcfe_last_error",
           "wcfe_upload_mesh", "wcfe_upload_state", "wcfe_download_state", "wcfe_upload_states", "wcfe_download_states", "wcfe_mesh_sizes", "wcfe_set_rng", "wcfe_set_debug",
           "wcfe_rng_skip", "wcfe_set_profile", "wcfe_get_kernel_ms", "wcfe_set_force_general", "wcfe_get_path_info",
           "wcfe_predict", "wcfe_commit", "wcfe_update", "wcfe_update_cell", "wcfe_update_subs", "wcfe_driver_velocities", "wcfe_driver_velocities_kind",
           "wcfe_mark_positions_committed", "wcfe_step_n", "wcfe_step_n_async", "wcfe_step_n_roundtrip", "wcfe_get_diagnostics",
           "wcfe_get_discrete", "wcfe_get_boundary_flags", "wcfe_kernel_launches", "wcfe_last_step_ms", "wcfe_force_field", "wcfe_dormant_field", "wcfe_device_ptr", "wcfe_sync", "wcfe_selftest_quotient", "wcfe_grid_refreshes",
           "wcfe_create_batch", "wcfe_select_replica", "wcfe_replicas", "wcfe_set_replica_seed", "wcfe_set_roundtrip_mode",
           "wcfe_halo_info_get", "wcfe_halo_connect", "wcfe_halo_disconnect", "wcfe_cell_halo_info_get", "wcfe_cell_halo_connect", "wcfe_cell_halo_disconnect",
           "wcfe_strip_partition", "wcfe_strip_sizes", "wcfe_strip_mesh", "wcfe_strip_halo", "wcfe_strip_free", "wcfe_strip_owner", "wcfe_key_partition", "wcfe_partition_weights", "wcfe_sweep_cell", "wcfe_ensemble_shard", "wcfe_morton_order",
           "wcfe_cell_partition", "wcfe_cell_plan_sizes", "wcfe_cell_plan_lists", "wcfe_cell_plan_free"]


class Params(C.Structure):
    _fields_ = [("law", C.c_int), ("kint_force", C.c_double), ("kint_bond", C.c_double), ("subs_scale", C.c_double),
                ("mu_cell", C.c_double), ("mu_subs", C.c_double),
                ("t_act_rate", C.c_double), ("t_act_max", C.c_double), ("t_act_t1", C.c_double),
                ("thick", C.c_double), ("kappa", C.c_double), ("thetag_dot", C.c_double),
                ("damp", C.c_double), ("rho", C.c_double), ("drag", C.c_double),
                ("rho_i_max", C.c_double), ("k_bend", C.c_double), ("k_area", C.c_double), ("area_0", C.c_double),
                ("ap_rate", C.c_double), ("ap_quantum", C.c_double), ("ap_density", C.c_double),
                ("k_on", C.c_double), ("ka", C.c_double), ("kb", C.c_double), ("Fa", C.c_double), ("Fb", C.c_double),
                ("pin_limit", C.c_double), ("t_i", C.c_double), ("t_f", C.c_double),
                ("md_px", C.c_double * 6), ("md_py", C.c_double * 6), ("md_k", C.c_double * 6),
                ("c_A_init", C.c_double)]


class HaloInfo(C.Structure):
    """wcfe_halo_info: what a neighbouring rank needs to reach this context's strain buffers (ship it as bytes)"""
    _fields_ = [("pid", C.c_longlong), ("device", C.c_int), ("ea_cur", C.c_int), ("np", C.c_longlong),
                ("ptr", C.c_ulonglong * 3), ("offset", C.c_ulonglong * 3), ("ipc", (C.c_ubyte * 64) * 3)]


class CellHaloInfo(C.Structure):
    """wcfe_cell_halo_info: what another rank needs to reach this context's cell buffers (ship it as bytes)"""
    _fields_ = [("pid", C.c_longlong), ("device", C.c_int), ("reserved", C.c_int), ("np", C.c_longlong),
                ("ptr", C.c_ulonglong * 11), ("offset", C.c_ulonglong * 11), ("ipc", (C.c_ubyte * 64) * 11)]


_lib = None


def load():
    """Load libwcfe.so.  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libwcfe.so is not built (%s missing): run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `python -m wholecellfe_b200.build`; there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.wcfe_default_params.argtypes = [C.POINTER(Params), C.c_int]
    L.wcfe_create.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(Params)]
    L.wcfe_destroy.argtypes = [vp]
    L.wcfe_create_batch.argtypes = [C.POINTER(vp), C.c_int, C.POINTER(Params), C.c_int]
    L.wcfe_select_replica.argtypes = [vp, C.c_int]
    L.wcfe_replicas.argtypes = [vp]
    L.wcfe_set_replica_seed.argtypes = [vp, C.c_int, C.c_uint]
    L.wcfe_set_params.argtypes = [vp, C.POINTER(Params)]
    L.wcfe_set_stream.argtypes = [vp, vp]
    L.wcfe_last_error.argtypes = [vp]; L.wcfe_last_error.restype = C.c_char_p
    L.wcfe_upload_mesh.argtypes = [vp, C.c_int, C.c_int, _dp, C.c_int, _ip]
    L.wcfe_upload_state.argtypes = [vp, C.c_int, C.c_int, _dp]
    L.wcfe_download_state.argtypes = [vp, C.c_int, C.c_int, _dp]
    L.wcfe_upload_states.argtypes = [vp, C.c_int, C.c_int, _ip, C.POINTER(_dp)]
    L.wcfe_download_states.argtypes = [vp, C.c_int, C.c_int, _ip, C.POINTER(_dp)]
    L.wcfe_mesh_sizes.argtypes = [vp, C.c_int, _ip, _ip]
    L.wcfe_set_rng.argtypes = [vp, C.c_int, C.c_uint]
    L.wcfe_set_debug.argtypes = [vp, C.c_int]
    L.wcfe_rng_skip.argtypes = [vp, C.c_longlong]
    L.wcfe_set_profile.argtypes = [vp, C.c_int]
    L.wcfe_set_force_general.argtypes = [vp, C.c_int]
    L.wcfe_get_path_info.argtypes = [vp, _dp]
    L.wcfe_get_kernel_ms.argtypes = [vp, _dp, C.POINTER(C.c_longlong)]
    L.wcfe_predict.argtypes = [vp, C.c_int, C.c_double, C.c_double, _dp]
    L.wcfe_commit.argtypes = [vp, C.c_int, _dp]
    L.wcfe_update.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    L.wcfe_update_cell.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    L.wcfe_update_subs.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    L.wcfe_driver_velocities.argtypes = [vp, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, _dp]
    L.wcfe_driver_velocities_kind.argtypes = [vp, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, _dp]
    L.wcfe_mark_positions_committed.argtypes = [vp]
    L.wcfe_step_n.argtypes = [vp, C.c_int, C.c_double, _dp]
    L.wcfe_step_n_async.argtypes = [vp, C.c_int, C.c_double, _dp]
    L.wcfe_step_n_roundtrip.argtypes = [vp, C.c_int, C.c_double, _dp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    L.wcfe_set_roundtrip_mode.argtypes = [vp, C.c_int]
    L.wcfe_get_diagnostics.argtypes = [vp, _dp]
    L.wcfe_get_discrete.argtypes = [vp, C.c_int, _ip]
    L.wcfe_get_boundary_flags.argtypes = [vp, C.c_int, _ip]
    L.wcfe_kernel_launches.argtypes = [vp]; L.wcfe_kernel_launches.restype = C.c_longlong
    L.wcfe_last_step_ms.argtypes = [vp]; L.wcfe_last_step_ms.restype = C.c_double
    L.wcfe_force_field.argtypes = [vp, C.c_int, C.c_double, C.c_double, _dp, _dp, _dp]
    L.wcfe_dormant_field.argtypes = [vp, C.c_int, C.c_double, C.c_double, _dp]
    L.wcfe_device_ptr.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)]
    L.wcfe_sync.argtypes = [vp]
    L.wcfe_grid_refreshes.argtypes = [vp]; L.wcfe_grid_refreshes.restype = C.c_longlong
    L.wcfe_selftest_quotient.argtypes = [C.c_int, C.c_longlong, C.c_uint, C.POINTER(C.c_longlong)]
    L.wcfe_halo_info_get.argtypes = [vp, C.POINTER(HaloInfo)]
    L.wcfe_halo_connect.argtypes = [vp, C.c_int, C.POINTER(HaloInfo), _ip, _ip, C.c_int, _ip, _ip, _ip, C.c_int, _ip]
    L.wcfe_halo_disconnect.argtypes = [vp]
    L.wcfe_cell_halo_info_get.argtypes = [vp, C.POINTER(CellHaloInfo)]
    L.wcfe_cell_halo_connect.argtypes = [vp, C.c_int, C.POINTER(CellHaloInfo), _ip, C.c_int, _ip, _ip, C.c_int, _ip, _ip, C.c_int, _ip, C.c_double]
    L.wcfe_cell_halo_disconnect.argtypes = [vp]
    L.wcfe_strip_partition.argtypes = [C.c_int, _dp, C.c_int, _ip, _dp, C.c_int, C.c_int, C.POINTER(vp)]
    L.wcfe_strip_sizes.argtypes = [vp] + [_ip] * 6
    L.wcfe_strip_mesh.argtypes = [vp] + [_ip] * 5
    L.wcfe_strip_halo.argtypes = [vp] + [_ip] * 6
    L.wcfe_strip_free.argtypes = [vp]
    L.wcfe_strip_owner.argtypes = [vp, _ip]
    L.wcfe_morton_order.argtypes = [C.c_int, _dp, C.c_int, _ip, _ip, _ip]
    L.wcfe_sweep_cell.argtypes = [C.c_int, C.c_int, _ip, C.c_double, C.c_double, C.c_int, C.POINTER(Params), C.POINTER(C.c_uint)]
    L.wcfe_ensemble_shard.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _ip]
    L.wcfe_partition_weights.argtypes = [C.c_int, _dp, C.c_int, _dp, C.c_double, C.c_double, _dp]
    L.wcfe_key_partition.argtypes = [C.c_int, _dp, C.c_int, _ip, _dp, C.c_int, C.c_int, C.POINTER(vp)]
    L.wcfe_cell_partition.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _ip, C.c_int, _dp, C.c_int, _ip, C.c_double, C.POINTER(vp)]
    L.wcfe_cell_plan_sizes.argtypes = [vp] + [_ip] * 4
    L.wcfe_cell_plan_lists.argtypes = [vp] + [_ip] * 7
    L.wcfe_cell_plan_free.argtypes = [vp]
    _lib = L
    return L
