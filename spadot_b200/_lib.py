"""ctypes binding of libspadot_b200.so (include/spadot_b200.h).

The library is the product: there is NO CPU fallback.  Importing this module without the
built shared object raises ImportError; calling into it without a B200-class device
raises RuntimeError (see `require_device`).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspadot_b200.so")

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_l = ctypes.c_int64
c_u = ctypes.c_uint64
c_d = ctypes.c_double
c_f = ctypes.c_float

# name -> argtypes; every function returns int status unless noted.  Kept in the same order as
# include/spadot_b200.h; tests/test_abi.py checks that header and table agree symbol by symbol.
SIGNATURES = {
    "sdb_device_check": [],
    "sdb_column_sums_f64": [c_p, c_l, c_i, c_p, c_p],
    "sdb_prep_points_f64": [c_p, c_l, c_i, c_p, c_p, c_l, c_i, c_p, c_p],
    "sdb_lse_pass_simt": [c_p, c_l, c_l, c_p, c_l, c_l, c_i, c_p, c_d, c_p, c_i, c_p, c_p],
    "sdb_sinkhorn_sweeps": [c_p, c_i, c_i, c_i, c_p],
    "sdb_sinkhorn_sweeps_persistent": [c_p, c_i, c_i, c_i, c_p, c_p],
    "sdb_sinkhorn_solve_persistent": [c_p, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p],
    "sdb_lse_finalize": [c_p, c_i, c_l, c_p, c_d, c_p, c_p],
    "sdb_lse_finalize_pred": [c_p, c_i, c_l, c_p, c_d, c_p, c_p, c_p, c_p],
    "sdb_finalize_update_pred": [c_p, c_i, c_l, c_p, c_d, c_p, c_p, c_d, c_d, c_d, c_p, c_p, c_p, c_p, c_p, c_i, c_d, c_d, c_p, c_p, c_p],
    "sdb_partial_sums_f64": [c_p, c_i, c_l, c_p, c_p, c_p, c_i, c_p, c_p],
    "sdb_update_from_sums_f64": [c_p, c_p, c_l, c_p, c_d, c_p, c_p, c_d, c_d, c_d, c_p, c_p, c_p, c_p, c_p, c_i, c_d, c_d, c_p, c_p],
    "sdb_nccl_available": [],
    "sdb_nccl_unique_id": [c_p],
    "sdb_nccl_comm_create": [c_p, c_i, c_i, c_p],
    "sdb_nccl_comm_destroy": [c_p],
    "sdb_nccl_allreduce_sum_f64": [c_p, c_p, c_l, c_p],
    "sdb_sinkhorn_sweeps_dist": [c_p, c_p, c_p, c_i, c_i, c_p],
    "sdb_lse_pass_tc_fused": [c_p, c_l, c_l, c_p, c_l, c_l, c_i, c_p, c_f, c_i, c_i, c_p, c_p, c_p, c_p],
    "sdb_potential_update_deferred": [c_l, c_p, c_p, c_p, c_d, c_d, c_d, c_d, c_p, c_p, c_p, c_p, c_p, c_i, c_d, c_d, c_p],
    "sdb_absorb_pending": [c_l, c_l, c_p, c_i, c_p, c_p, c_p, c_p, c_p, c_p],
    "sdb_lse_pass_tc_groups": [c_p, c_l, c_l, c_p, c_l, c_i, c_p, c_f, c_p, c_i, c_i, c_p, c_p],
    "sdb_lse_pass_tc_pred": [c_p, c_l, c_l, c_p, c_l, c_l, c_i, c_p, c_f, c_i, c_i, c_p, c_p, c_p],
    "sdb_potential_update": [c_l, c_p, c_p, c_p, c_d, c_d, c_d, c_d, c_p, c_p, c_p, c_p, c_p, c_i, c_d, c_d, c_p],
    "sdb_finalize_update": [c_p, c_i, c_l, c_p, c_d, c_p, c_p, c_d, c_d, c_d, c_p, c_p, c_p, c_p, c_p, c_i, c_d, c_d, c_p],
    "sdb_make_bias": [c_l, c_l, c_p, c_p, c_d, c_d, c_p, c_p],
    "sdb_absmax_centered_f64": [c_p, c_l, c_i, c_p, c_p, c_p],
    "sdb_prep_points_split_f16": [c_p, c_l, c_i, c_p, c_i, c_p, c_l, c_i, c_p, c_p],
    "sdb_prep_points_split_f16_scaled": [c_p, c_l, c_i, c_p, c_d, c_p, c_l, c_i, c_p, c_p],
    "sdb_lse_pass_tc": [c_p, c_l, c_l, c_p, c_l, c_l, c_i, c_p, c_f, c_i, c_i, c_p, c_p],
    "sdb_absorb": [c_l, c_l, c_p, c_i, c_p, c_p, c_p, c_p, c_p],
    "sdb_stage_criterion": [c_l, c_l, c_p, c_p, c_p, c_p, c_p, c_p, c_d, c_p, c_p, c_p],
    "sdb_gap_terms": [c_l, c_l, c_p, c_p, c_p, c_p, c_p, c_p, c_d, c_d, c_d, c_d, c_d, c_p, c_p, c_p],
    "sdb_sum_exp": [c_l, c_p, c_p, c_d, c_p, c_p, c_p],
    "sdb_plan_dense_f64": [c_p, c_p, c_l, c_l, c_i, c_p, c_p, c_d, c_d, c_d, c_p, c_p],
    "sdb_row_mass": [c_l, c_p, c_p, c_d, c_d, c_p, c_p],
    "sdb_pair_distances_f64": [c_p, c_p, c_i, c_p, c_p, c_l, c_p, c_p],
    "sdb_cost_histogram": [c_p, c_l, c_l, c_p, c_l, c_l, c_i, c_p, c_i, c_f, c_f, c_i, c_p, c_p, c_p],
    "sdb_cost_collect": [c_p, c_l, c_l, c_p, c_l, c_l, c_i, c_p, c_i, c_f, c_f, c_p, c_p, c_i, c_p, c_u, c_p, c_p],
    "sdb_cost_histogram_tc": [c_p, c_l, c_l, c_p, c_l, c_l, c_i, c_p, c_p, c_f, c_i, c_i, c_f, c_f, c_i, c_p, c_p, c_p],
    "sdb_cost_collect_tc": [c_p, c_l, c_l, c_p, c_l, c_l, c_i, c_p, c_p, c_f, c_i, c_i, c_f, c_f, c_p, c_p, c_i, c_p, c_u, c_p, c_p],
    "sdb_radix_digit_hist": [c_p, c_u, c_i, c_u, c_p, c_p],
    "sdb_select_ranks_f64": [c_p, c_u, c_p, c_i, c_p, c_p, c_p],
    "sdb_transition_accumulate": [c_p, c_i, c_l, c_p, c_d, c_p, c_d, c_d, c_p, c_i, c_p, c_p],
    "sdb_dense_row_lse_f64": [c_p, c_l, c_l, c_l, c_p, c_d, c_p, c_p],
    "sdb_dense_col_lse_f64": [c_p, c_l, c_l, c_l, c_p, c_d, c_p, c_p, c_i, c_p],
    "sdb_dense_plan_f64": [c_p, c_l, c_l, c_l, c_p, c_p, c_d, c_d, c_p, c_p],
    "sdb_kernel_block_f64": [c_p, c_l, c_p, c_l, c_i, c_i, c_d, c_d, c_p, c_l, c_p],
    "sdb_kernel_diag_f64": [c_p, c_p, c_l, c_i, c_i, c_d, c_p, c_p],
    "sdb_quad_form_rows_f64": [c_p, c_p, c_l, c_i, c_p, c_p],
    "sdb_gat_forward": [c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_d, c_i, c_p, c_p, c_p],
    "sdb_kmeans_assign": [c_p, c_p, c_l, c_i, c_i, c_p, c_p, c_p, c_p, c_p],
    "sdb_kmeans_update": [c_p, c_p, c_p, c_p, c_i, c_i, c_p, c_p],
    "sdb_kmeans_inertia": [c_p, c_p, c_p, c_l, c_i, c_p, c_p, c_p],
    "sdb_kmeans_lloyd_runs": [c_p, c_l, c_i, c_i, c_p, c_i, c_i, c_d, c_p, c_p, c_p, c_p, c_p, c_p],
    "sdb_knn_f64": [c_p, c_l, c_i, c_i, c_p, c_p, c_p],
    "sdb_knn_grid_f64": [c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_d, c_d, c_d, c_i, c_p, c_p, c_p],
    "sdb_pipe_peak": [c_i, c_i, c_i, c_p, c_p, c_p],
    "sdb_gat_backward": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_d, c_i, c_p, c_p, c_p, c_p, c_p, c_p, c_p],
    "sdb_gat_backward_prefix": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_l, c_i, c_i, c_d, c_i, c_p, c_p, c_p, c_p, c_p,
                                c_p, c_p],
    "sdb_gat_forward_shared": [c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_i, c_i, c_d, c_i, c_p, c_p, c_p],
    "sdb_gat_backward_shared": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_l, c_l, c_i, c_i, c_d, c_i, c_p, c_p, c_p, c_p, c_p,
                                c_p, c_p],
}

class SweepDesc(ctypes.Structure):
    """struct sdb_sweep_desc of include/spadot_b200.h (field order and types must match)."""
    _fields_ = [("n", c_l), ("m", c_l), ("n_total", c_l),
                ("use_tc", ctypes.c_int32), ("dpad", ctypes.c_int32), ("dp", ctypes.c_int32), ("n_ctas", ctypes.c_int32),
                ("xt", c_p), ("ldx", c_l), ("yt", c_p), ("ldy", c_l),
                ("x16", c_p), ("n_pad", c_l), ("y16", c_p), ("m_pad", c_l),
                ("norms_x", c_p), ("norms_y", c_p), ("bounds_row", c_p), ("bounds_col", c_p),
                ("ns_row", ctypes.c_int32), ("ns_col", ctypes.c_int32), ("tps_row", ctypes.c_int32), ("tps_col", ctypes.c_int32),
                ("partial_row", c_p), ("partial_col", c_p), ("bias_x", c_p), ("bias_y", c_p), ("m_bias", c_l),
                ("f", c_p), ("g", c_p), ("u", c_p), ("v", c_p), ("la_old", c_p), ("lb_old", c_p), ("Lr", c_p), ("Lc", c_p),
                ("logp", c_p), ("logq", c_p), ("flag", c_p),
                ("eps", c_d), ("inv_med", c_d), ("alpha1", c_d), ("alpha2", c_d), ("log_tau", c_d), ("log_floor", c_d),
                ("pow2_scale", c_d),
                ("m_x", c_p), ("m_y", c_p), ("bad_flag", c_p), ("pred_from_row", ctypes.c_int32), ("pred_from_col", ctypes.c_int32),
                ("simt_direct", ctypes.c_int32), ("reserved0", ctypes.c_int32), ("flag2", c_p), ("tile_counters", c_p)]


class SolveParams(ctypes.Structure):
    """struct sdb_solve_params"""
    _fields_ = [("lambda1", c_d), ("lambda2", c_d), ("epsilon", c_d), ("epsilon0", c_d), ("tolerance", c_d), ("tau", c_d),
                ("max_iter", c_d), ("eps_stage", c_d * 6), ("xy_max", c_d), ("dot_limit", c_d),
                ("batch_size", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class SolveResult(ctypes.Structure):
    """struct sdb_solve_result"""
    _fields_ = [("iters", ctypes.c_int32 * 6), ("total_iters", ctypes.c_int32), ("status", ctypes.c_int32),
                ("max_iter_reached", ctypes.c_int32), ("last_tick", ctypes.c_int32), ("gap", c_d), ("eps_final", c_d),
                ("stage_gap", c_d * 6)]


SOLVE_MAX_CTAS = 1024

_lib = None


class SpadotB200Error(RuntimeError):
    pass


def load():
    """Load the shared object (no GPU needed to load it or to resolve its symbols)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(spadot_b200 has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.sdb_version.restype = c_i
    lib.sdb_version.argtypes = []
    lib.sdb_error_string.restype = ctypes.c_char_p
    lib.sdb_error_string.argtypes = [c_i]
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_i
    _lib = lib
    return lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = load().sdb_error_string(status).decode()
        raise SpadotB200Error(f"{what or 'libspadot_b200'} failed: {msg} (status {status})")


_fn_cache = {}


def call(name: str, *args):
    fn = _fn_cache.get(name)
    if fn is None:
        fn = _fn_cache[name] = getattr(load(), name)
    status = fn(*args)
    if status != 0:
        check(status, name)


_device_ok = set()


def require_device():
    """Fail loudly when the CUDA path cannot run (no silent fallback).  The (slow) device-property query
    runs once per device per process."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("spadot_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.cuda.current_device()
    if dev not in _device_ok:
        check(load().sdb_device_check(), "sdb_device_check")
        _device_ok.add(dev)
