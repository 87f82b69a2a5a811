"""ctypes declarations of include/simpimc_b200.h and the loader of the CUDA library.

There is no CPU fallback: if simpimc_b200/csrc/libsimpimc_b200.so is missing or cannot be
loaded, `lib()` raises.  (The struct declarations and table packers are importable without
the library so the oracle wrapper can reuse them.)
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# SIMPIMC_B200_LIB selects another build of the same library (A/B timing of kernel variants)
LIB_PATH = os.environ.get("SIMPIMC_B200_LIB") or os.path.join(HERE, "csrc", "libsimpimc_b200.so")

PIMC_OLD, PIMC_NEW = 0, 1
GRID_GENERAL, GRID_LOG, GRID_LINEAR = 0, 1, 2

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class Config(C.Structure):
    _fields_ = [("n_d", C.c_int32), ("pbc", C.c_int32), ("L", C.c_double), ("beta", C.c_double), ("n_bead", C.c_int32),
                ("n_species", C.c_int32), ("n_part", c_int32_p), ("lam", c_double_p), ("n_clones", C.c_int32),
                ("device", C.c_int32), ("slice_lo", C.c_int32), ("slice_hi", C.c_int32)]


class Table1D(C.Structure):
    _fields_ = [("n", C.c_int32), ("r", c_double_p), ("f", c_double_p)]


class Table2D(C.Structure):
    _fields_ = [("n_x", C.c_int32), ("n_y", C.c_int32), ("x", c_double_p), ("y", c_double_p), ("f", c_double_p)]


class LongRange(C.Structure):
    _fields_ = [("f_r", Table1D), ("f_r_0", C.c_double), ("n_k", C.c_int32), ("k", c_double_p), ("f_k", c_double_p),
                ("f_k_0", C.c_double)]


class IlkkaTables(C.Structure):
    _fields_ = [("u_xy", Table2D), ("du_xy", Table2D), ("v_r", Table1D), ("u_long", LongRange), ("du_long", LongRange),
                ("v_long", LongRange)]


class BareTables(C.Structure):
    _fields_ = [("v_r", Table1D), ("v_long", LongRange), ("is_coulomb", C.c_int32)]


class DavidTables(C.Structure):
    _fields_ = [("grid_type", C.c_int32), ("r_start", C.c_double), ("r_end", C.c_double), ("n_grid", C.c_int32),
                ("grid_points", c_double_p), ("n_order", C.c_int32), ("n_tau", C.c_int32), ("taus", c_double_p),
                ("u_kj", c_double_p), ("du_kj_dbeta", c_double_p), ("potential", c_double_p), ("n_k", C.c_int32),
                ("k_points", c_double_p), ("u_k", c_double_p), ("v_image", C.c_double)]


class _Keep:
    """Holds the numpy arrays a packed struct points into."""

    def __init__(self):
        self.arrays = []

    def ptr(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        self.arrays.append(a)
        return a.ctypes.data_as(c_double_p)



def _f(x):
    """Scalar dataset (a 0-d value, or the 1-element array the flat table container stores) -> float."""
    return float(np.asarray(x).reshape(-1)[0])


def _i(x):
    return int(np.asarray(x).reshape(-1)[0])


def _t1(keep, r, f):
    return Table1D(len(r), keep.ptr(r), keep.ptr(f))


def _t2(keep, x, y, f):
    return Table2D(len(x), len(y), keep.ptr(x), keep.ptr(y), keep.ptr(f))


def _lr(keep, t, obj, present):
    if not present:
        return LongRange()
    p = obj + "/diag/"
    return LongRange(_t1(keep, t[p + "r_long"], t[p + obj + "_long_r"]), _f(t[p + obj + "_long_r_0"]), _i(t[p + "n_k"]),
                     keep.ptr(t[p + "k"]), keep.ptr(t[p + obj + "_long_k"]), _f(t[p + obj + "_long_k_0"]))


def pack_ilkka(t, use_long_range):
    """Table dict (HDF5 dataset names of ilkka_pair_action_class.h:266-418) -> struct."""
    keep = _Keep()
    s = IlkkaTables()
    s.u_xy = _t2(keep, t["u/off_diag/x"], t["u/off_diag/y"], t["u/off_diag/u_xy"])
    s.du_xy = _t2(keep, t["du/off_diag/x"], t["du/off_diag/y"], t["du/off_diag/du_xy"])
    s.v_r = _t1(keep, t["v/diag/r"], t["v/diag/v_r"])
    s.u_long = _lr(keep, t, "u", use_long_range)
    s.du_long = _lr(keep, t, "du", use_long_range)
    s.v_long = _lr(keep, t, "v", use_long_range)
    return s, keep


def pack_bare(t, use_long_range, is_coulomb=False):
    keep = _Keep()
    s = BareTables()
    s.v_r = _t1(keep, t["v/diag/r"], t["v/diag/v_r"])
    s.v_long = _lr(keep, t, "v", use_long_range)
    s.is_coulomb = 1 if is_coulomb else 0
    return s, keep


def pack_david(t, n_order, use_long_range):
    keep = _Keep()
    g = "u_kj_%d" % n_order
    s = DavidTables()
    gt = t[g + "/grid/type"]
    s.grid_type = GRID_LOG if ("LOG" in gt and "LOGLIN" not in gt) else (GRID_LINEAR if "LINEAR" in gt else GRID_GENERAL)
    if "LOGLIN" in gt:
        raise ValueError("LOGLIN grids are fork-only in the reference's einspline and not supported")
    s.r_start = _f(t[g + "/grid/start"])
    s.r_end = _f(t[g + "/grid/end"])
    s.n_grid = _i(t[g + "/grid/n_grid_points"])
    s.grid_points = keep.ptr(t[g + "/grid/grid_points"])
    s.n_order = n_order
    s.n_tau = len(t[g + "/taus"])
    s.taus = keep.ptr(t[g + "/taus"])
    s.u_kj = keep.ptr(t[g + "/data"])
    s.du_kj_dbeta = keep.ptr(t["du_kj_dbeta_%d/data" % n_order])
    s.potential = keep.ptr(t["potential/data"])
    if use_long_range:
        s.n_k = _i(t["long_range/n_k"])
        s.k_points = keep.ptr(t["long_range/k_points"])
        s.u_k = keep.ptr(t["long_range/u_k"])
        s.v_image = _f(t["squarer/v_image"])
    return s, keep


def make_config(cfg, n_clones=1, device=0, slice_lo=0, slice_hi=None):
    """SystemConfig -> (Config struct, keep-alive list)."""
    n_part = (C.c_int32 * len(cfg.species))(*[s.n_part for s in cfg.species])
    lam = (C.c_double * len(cfg.species))(*[s.lam for s in cfg.species])
    c = Config(cfg.n_d, 1 if cfg.pbc else 0, cfg.L, cfg.beta, cfg.n_bead, len(cfg.species),
               C.cast(n_part, c_int32_p), C.cast(lam, c_double_p), n_clones, device, slice_lo,
               cfg.n_bead if slice_hi is None else slice_hi)
    return c, (n_part, lam)


EXPORTS = [
    "pimc_last_error", "pimc_version", "pimc_ctx_create", "pimc_ctx_destroy", "pimc_ctx_sync", "pimc_ctx_stream",
    "pimc_kspace_setup", "pimc_kspace_get", "pimc_positions_upload", "pimc_positions_download",
    "pimc_positions_set_device", "pimc_positions_device_ptr", "pimc_rhok_rebuild", "pimc_rhok_download",
    "pimc_action_create_ilkka", "pimc_action_create_bare", "pimc_action_create_david", "pimc_action_destroy",
    "pimc_action_dbeta", "pimc_action_potential", "pimc_action_dbeta_device", "pimc_action_potential_device",
    "pimc_action_get", "pimc_action_gradient", "pimc_action_laplacian", "pimc_action_total", "pimc_action_total_device", "pimc_action_accept", "pimc_action_reject",
    "pimc_action_calc_pair", "pimc_propose", "pimc_beads_download", "pimc_commit", "pimc_est_gofr", "pimc_est_gofr_counts", "pimc_est_sofk",
    "pimc_ctx_launch_count", "pimc_fp64_peak", "pimc_ctx_set_timing", "pimc_ctx_kernel_time",
    "pimc_action_calc_pair_fast", "pimc_debug_fast_sqrt", "pimc_debug_interval_table", "pimc_ctx_force_general", "pimc_bisect_sweep", "pimc_displace_sweep", "pimc_perm_table", "pimc_halo_pack", "pimc_halo_unpack", "pimc_rotate_pack", "pimc_rotate_apply",
    "pimc_comm_unique_id", "pimc_comm_init", "pimc_comm_destroy", "pimc_comm_bytes_sent", "pimc_halo_exchange", "pimc_allreduce_sum", "pimc_rotate",
    "pimc_action_create_kinetic", "pimc_move_set_images", "pimc_bisect_sweep_windows", "pimc_sharded_evaluate", "pimc_capture_begin", "pimc_capture_end", "pimc_graph_launch", "pimc_graph_nodes", "pimc_graph_destroy",
    "pimc_perm_bisect_sweep", "pimc_permutation_get", "pimc_permutation_set", "pimc_perm_last_cycle", "pimc_debug_bucket_spline",
]

_lib = None


def lib():
    """Load the CUDA library; raises if it has not been built (there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("simpimc_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    i32 = C.c_int32
    dbl = C.c_double
    L.pimc_last_error.restype = C.c_char_p
    L.pimc_ctx_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.pimc_ctx_destroy.argtypes = [vp]
    L.pimc_ctx_sync.argtypes = [vp]
    L.pimc_ctx_stream.restype = vp
    L.pimc_ctx_stream.argtypes = [vp]
    L.pimc_kspace_setup.argtypes = [vp, dbl, c_int32_p]
    L.pimc_kspace_get.argtypes = [vp, vp, vp]
    L.pimc_positions_upload.argtypes = [vp, i32, i32, i32, vp]
    L.pimc_positions_download.argtypes = [vp, i32, i32, i32, i32, vp]
    L.pimc_positions_set_device.argtypes = [vp, i32, vp]
    L.pimc_positions_device_ptr.restype = vp
    L.pimc_positions_device_ptr.argtypes = [vp, i32]
    L.pimc_rhok_rebuild.argtypes = [vp, i32]
    L.pimc_rhok_download.argtypes = [vp, i32, i32, i32, vp]
    L.pimc_action_create_ilkka.argtypes = [vp, i32, i32, C.POINTER(IlkkaTables), i32, i32, dbl, C.POINTER(vp)]
    L.pimc_action_create_bare.argtypes = [vp, i32, i32, C.POINTER(BareTables), i32, i32, dbl, C.POINTER(vp)]
    L.pimc_action_create_david.argtypes = [vp, i32, i32, C.POINTER(DavidTables), i32, i32, C.POINTER(vp)]
    L.pimc_action_destroy.argtypes = [vp]
    for n in ("pimc_action_dbeta", "pimc_action_potential", "pimc_action_dbeta_device", "pimc_action_potential_device",
              "pimc_action_total", "pimc_action_total_device"):
        getattr(L, n).argtypes = [vp, vp]
    L.pimc_action_get.argtypes = [vp, i32, vp, i32, i32, vp, vp, i32, vp]
    L.pimc_action_gradient.argtypes = [vp, vp, i32, i32, vp, vp, i32, vp]
    L.pimc_action_laplacian.argtypes = [vp, vp, i32, i32, vp, vp, i32, vp]
    L.pimc_action_accept.argtypes = [vp]
    L.pimc_action_reject.argtypes = [vp]
    L.pimc_action_calc_pair.argtypes = [vp, i32, i32, vp, vp, vp, i32, vp]
    L.pimc_propose.argtypes = [vp, i32, vp, vp, i32, vp]
    L.pimc_commit.argtypes = [vp, vp]
    L.pimc_beads_download.argtypes = [vp, i32, vp, vp, i32, vp]
    L.pimc_est_gofr.argtypes = [vp, i32, i32, dbl, dbl, i32, vp, vp]
    L.pimc_est_gofr_counts.argtypes = [vp, i32, i32, dbl, dbl, i32, vp]
    L.pimc_est_sofk.argtypes = [vp, i32, i32, dbl, vp, vp]
    L.pimc_ctx_launch_count.restype = C.c_int64
    L.pimc_ctx_launch_count.argtypes = [vp]
    L.pimc_fp64_peak.argtypes = [vp, c_double_p]
    L.pimc_ctx_set_timing.argtypes = [vp, i32]
    L.pimc_ctx_kernel_time.argtypes = [vp, i32, c_double_p, C.POINTER(C.c_int64)]
    L.pimc_action_calc_pair_fast.argtypes = [vp, i32, i32, vp, vp, vp, vp]
    L.pimc_debug_fast_sqrt.argtypes = [vp, i32, vp, vp]
    L.pimc_debug_interval_table.argtypes = [i32, i32, vp, i32, vp, vp, vp]
    L.pimc_debug_bucket_spline.argtypes = [i32, vp, vp, i32, vp, vp, vp, vp]
    L.pimc_ctx_force_general.argtypes = [vp, i32]
    L.pimc_halo_pack.argtypes = [vp, i32, vp]
    L.pimc_halo_unpack.argtypes = [vp, i32, vp]
    L.pimc_rotate_pack.argtypes = [vp, i32, i32, vp]
    L.pimc_rotate_apply.argtypes = [vp, i32, i32, vp]
    L.pimc_bisect_sweep.argtypes = [vp, i32, i32, i32, C.c_uint64, C.c_uint64, i32, vp]
    L.pimc_bisect_sweep_windows.argtypes = [vp, i32, i32, i32, C.c_uint64, C.c_uint64, i32, vp, c_int32_p]
    L.pimc_displace_sweep.argtypes = [vp, i32, C.c_double, i32, C.c_uint64, C.c_uint64, vp]
    L.pimc_perm_table.argtypes = [vp, i32, vp, i32, C.c_double, i32, vp]
    L.pimc_perm_bisect_sweep.argtypes = [vp, i32, i32, i32, C.c_uint64, C.c_uint64, i32, C.c_double, vp, vp, vp]
    L.pimc_permutation_get.argtypes = [vp, i32, vp]
    L.pimc_permutation_set.argtypes = [vp, i32, vp]
    L.pimc_perm_last_cycle.argtypes = [vp, vp, vp, vp, vp, vp]
    L.pimc_action_create_kinetic.argtypes = [vp, i32, i32, C.POINTER(vp)]
    L.pimc_move_set_images.argtypes = [vp, i32, i32]
    L.pimc_comm_unique_id.argtypes = [vp]
    L.pimc_comm_init.argtypes = [vp, vp, i32, i32, C.POINTER(vp)]
    L.pimc_comm_destroy.argtypes = [vp]
    L.pimc_comm_bytes_sent.restype = C.c_int64
    L.pimc_comm_bytes_sent.argtypes = [vp]
    L.pimc_halo_exchange.argtypes = [vp, vp, i32]
    L.pimc_allreduce_sum.argtypes = [vp, vp, vp, C.c_int64]
    L.pimc_rotate.argtypes = [vp, vp, i32]
    L.pimc_sharded_evaluate.argtypes = [vp, vp, i32, C.POINTER(vp), i32, vp]
    L.pimc_capture_begin.argtypes = [vp]
    L.pimc_capture_end.argtypes = [vp, C.POINTER(vp)]
    L.pimc_graph_launch.argtypes = [vp]
    L.pimc_graph_nodes.restype = C.c_int64
    L.pimc_graph_nodes.argtypes = [vp]
    L.pimc_graph_destroy.argtypes = [vp]
    _lib = L
    return L


def check(status):
    if status != 0:
        raise RuntimeError("simpimc_b200 C-ABI call failed (%d): %s" % (status, lib().pimc_last_error().decode()))
