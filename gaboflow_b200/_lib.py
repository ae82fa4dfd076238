"""ctypes binding of include/gabo_b200.h.  There is no fallback: if the CUDA library is missing or a call
fails, this raises."""
from __future__ import annotations

import ctypes as C
import os

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('GABO_LIB_PATH') or os.path.join(_PKG_DIR, 'libgabo_b200.so')   # the override is for A/B builds (tools/)

i32, i64, f64, f32, vp, sz = C.c_int, C.c_int64, C.c_double, C.c_float, C.c_void_p, C.c_size_t

# name -> (restype, argtypes); mirrors include/gabo_b200.h one to one
ABI_VERSION = 2      # == GABO_ABI_VERSION in include/gabo_b200.h (tests/test_abi_and_host.py pins the two)

SIGNATURES = {
    'gabo_abi_version': (i32, []),
    'gabo_launch_count': (C.c_longlong, []),
    'gabo_device_info': (i32, [C.POINTER(i32)] * 3),
    'gabo_status_string': (C.c_char_p, [i32]),
    'gabo_points_gram_workspace_bytes': (sz, [i64, i64, i32]),
    'gabo_points_gram_f64': (i32, [i32, vp, i64, i64, vp, i64, i64, i32, f64, f64, vp, i64, i64, i64, i32, vp, sz, vp]),
    'gabo_points_gram_f32': (i32, [i32, vp, i64, i64, vp, i64, i64, i32, f32, f32, vp, i64, i64, i64, i32, vp, sz, vp]),
    'gabo_points_gram_i8_workspace_bytes': (sz, [i64, i64]),
    'gabo_points_gram_i8_f64': (i32, [i32, vp, i64, i64, vp, i64, i64, i32, f64, f64, vp, i64, i64, i64, i32, i32, i32, vp, sz, vp]),
    'gabo_kdiag_const_f64': (i32, [i64, f64, vp, vp]),
    'gabo_spd_prep_f64': (i32, [vp, i64, i64, i32, vp, vp, vp, vp, vp, i64, vp, vp]),
    'gabo_spd_gram_f64': (i32, [i32, i32, vp, vp, vp, i64, vp, vp, i64, vp, i64, i32, f64, f64, vp, i64, i64, i64, i32, vp]),
    'gabo_spd_gram_cyclic_f64': (i32, [i32, i32, vp, vp, vp, vp, i64, i64, f64, f64, vp, i64, i32, i32, vp]),
    'gabo_spd_stein_kdiag_f64': (i32, [vp, i64, f64, f64, vp, vp]),
    'gabo_potrf_workspace_bytes': (sz, [i64]),
    'gabo_potrf_f64': (i32, [i64, vp, i64, f64, vp, vp, sz, vp]),
    'gabo_trsm_workspace_bytes': (sz, [i64, i64]),
    'gabo_trsm_f64': (i32, [i32, i64, i64, vp, i64, vp, i64, vp, sz, vp]),
    'gabo_trsm_winv_workspace_bytes': (sz, []),
    'gabo_trsm_winv_f64': (i32, [i32, i64, i64, vp, i64, vp, vp, i64, vp, sz, vp]),
    'gabo_logdet_f64': (i32, [i64, vp, i64, vp, vp]),
    'gabo_gp_loglik_f64': (i32, [i64, i64, vp, i64, vp, i64, vp, vp]),
    'gabo_gp_grad_beta_workspace_bytes': (sz, []),
    'gabo_gp_grad_beta_f64': (i32, [i64, i64, vp, i64, vp, i64, vp, i64, f64, f64, vp, vp, sz, vp]),
    'gabo_gp_predict_f64': (i32, [i64, i64, i64, vp, i64, vp, i64, vp, f64, vp, vp, vp]),
    'gabo_gp_append_row_f64': (i32, [i64, i64, vp, i64, f64, vp, i64, vp, f64, vp, vp, vp, vp]),
    'gabo_sphere_posterior_grad_f64': (i32, [i32, i64, i64, i32, vp, i64, vp, i64, vp, vp, i64, f64, f64, vp, vp, f64, vp, vp, vp, vp]),
    'gabo_spd_ai_posterior_grad_f64': (i32, [i32, i32, i64, i64, vp, vp, vp, vp, i64, f64, f64, vp, vp, f64, vp, vp, vp, vp]),
    'gabo_sphere_acq_fused_f64': (i32, [i32, i64, i64, i32, vp, i64, vp, i64, vp, i64, vp, vp, f64, f64, f64, f64, vp, vp, vp, vp, vp, vp, vp]),
    'gabo_expected_improvement_f64': (i32, [i64, vp, vp, f64, vp, vp]),
    'gabo_gp_small_batch_f64': (i32, [i64, i64, vp, i64, vp, i64, i64, vp, i32, vp, vp, vp, i64, i64, vp, vp]),
    'gabo_spd_expmap_f64': (i32, [i32, vp, i64, vp, i64, i64, i32, vp, i64, vp, vp]),
    'gabo_spd_logmap_f64': (i32, [i32, vp, i64, vp, i64, i64, i32, vp, i64, vp, vp]),
    'gabo_sym_to_mandel_f64': (i32, [i32, vp, i64, i64, vp, i64, vp]),
    'gabo_log_gram_f64': (i32, [vp, i64, vp, i64, i64, i64, f64, f64, vp]),
    'gabo_lanczos_sweep_workspace_bytes': (sz, [i64, i32, i32, i32]),
    'gabo_lanczos_sweep_f64': (i32, [vp, i64, i64, i64, i32, vp, i32, f64, i32, f64, C.c_ulonglong, vp, vp, vp, vp, sz, vp]),
    'gabo_math_probe_f64': (i32, [i32, vp, i64, f64, vp, vp]),
    'gabo_profile_trailing_updates': (i32, [i32]),
    'gabo_profile_trailing_read': (i32, [C.POINTER(i64), C.POINTER(f64), C.POINTER(f64), C.POINTER(f64), C.POINTER(f64)]),
    'gabo_gemm_nt_sub_f64': (i32, [i64, i64, i64, vp, i64, vp, i64, vp, i64, i32, vp]),
    'gabo_points_gram_cyclic_f64': (i32, [i32, vp, i64, i64, i32, f64, f64, vp, i64, i32, i32, vp, sz, vp]),
    'gabo_points_gram_cyclic_workspace_bytes': (sz, [i64, i32]),
    'gabo_points_gram_cyclic_hybrid_f64': (i32, [i32, vp, i64, i64, i32, f64, f64, vp, i64, i32, i32, i32, vp, sz, vp]),
    'gabo_points_gram_cyclic_stage_bytes': (sz, [i64, i32, i32, i32, i32, i32]),
    'gabo_points_gram_cyclic_staged_f64': (i32, [i32, vp, i64, i64, i32, f64, f64, vp, i64, i32, i32, i32, i32, i32, vp, sz, vp, i64, C.c_uint32, vp, sz, vp]),
    'gabo_enable_peer_access': (i32, [i32]),
    'gabo_shared_alloc': (i32, [sz, C.POINTER(vp)]),
    'gabo_shared_free': (i32, [vp]),
    'gabo_ipc_export': (i32, [vp, C.c_char_p]),
    'gabo_ipc_open': (i32, [C.c_char_p, C.POINTER(vp)]),
    'gabo_ipc_close': (i32, [vp]),
    'gabo_trsm_right_workspace_bytes': (sz, [i64, i64]),
    'gabo_trsm_right_f64': (i32, [i64, i64, vp, i64, vp, i64, vp, sz, vp]),
    'gabo_potrf_diag_f64': (i32, [i64, vp, i64, f64, vp, vp, vp, sz, vp]),
    'gabo_trsm_right_winv_f64': (i32, [i64, i64, vp, i64, vp, vp, i64, vp, sz, vp]),
    'gabo_trsm_right_scatter_f64': (i32, [i64, i64, vp, i64, vp, vp, i64, vp, i32, i64, i64, i32, i32, vp, sz, vp]),
    'gabo_flag_signal': (i32, [vp, i32, C.c_uint32, vp]),
    'gabo_flag_wait_geq': (i32, [vp, C.c_uint32, vp]),
    'gabo_push_signal_f64': (i32, [vp, i64, vp, vp, i32, C.c_uint32, vp]),
    'gabo_copy_async': (i32, [vp, vp, sz, vp]),
    'gabo_event_create': (i32, [C.POINTER(vp)]),
    'gabo_event_destroy': (i32, [vp]),
    'gabo_potrf_cyclic_p2p_f64': (i32, [i64, i32, i32, i32, vp, i64, f64, vp, vp, vp, C.c_uint32, vp, vp, vp, sz, vp, sz, vp, vp, vp, vp]),
    'gabo_potrf_cyclic_chain_f64': (i32, [i64, i32, i32, i32, vp, i64, f64, i32, vp, vp, vp, C.c_uint32, vp, vp, vp, sz, vp, sz, vp, vp]),
    'gabo_solve_cyclic_p2p_f64': (i32, [i64, i32, i32, i32, vp, i64, vp, i64, i64, vp, vp, vp, C.c_uint32, vp, sz, vp]),
    'gabo_syrk_cyclic_f64': (i32, [i64, i64, i64, vp, i64, vp, i64, vp, i64, i64, i64, i32, i32, i32, vp]),
    'gabo_solve_update_f64': (i32, [i64, i64, i64, vp, i64, vp, i64, vp, i64, vp]),
}

_lib = None


class GaboError(RuntimeError):
    def __init__(self, fn, status, text):
        super().__init__(f'{fn} failed with status {status}: {text}')
        self.status = status


def lib():
    """Load libgabo_b200.so once.  Raises ImportError (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} is missing: build it with `python -m gaboflow_b200.build` '
                '(there is no CPU or PyTorch fallback for the gaboflow_b200 kernels)')
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError if the ABI and the header ever diverge
            fn.restype = res
            fn.argtypes = args
        if handle.gabo_abi_version() != ABI_VERSION:
            raise ImportError('libgabo_b200.so ABI version mismatch; rebuild')
        _lib = handle
    return _lib


def check(fn_name, status):
    if status != 0:
        raise GaboError(fn_name, status, lib().gabo_status_string(status).decode())
