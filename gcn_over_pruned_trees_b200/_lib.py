"""ctypes binding of libgptb200.so (C ABI declared in include/gpt_b200.h).

The library is built in-tree by ``csrc/build.py`` (nvcc, sm_100a).  There is no fallback: if the shared object is
missing, ``lib()`` raises, and every op in ``ops.py`` goes through ``lib()``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libgptb200.so')

_c_int = ctypes.c_int
_c_f = ctypes.c_float
_c_u32 = ctypes.c_uint32
_p = ctypes.c_void_p
_c_ll = ctypes.c_longlong

# name -> argtypes; every function returns int except gpt_error_string
SIGNATURES = {
    'gpt_version': [],
    'gpt_launch_count': [],
    'gpt_l2_prefetch': [_p, _c_ll, _p],
    'gpt_prune_csr': [_p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _p, _p, _p, _p, _p, _p, _p, _p],
    'gpt_gcn_aggregate_fwd': [_p, _p, _p, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _c_f, _p, _c_u32, _p,
                              _c_int, _p],
    'gpt_gcn_aggregate_fwd_pool': [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _p],
    'gpt_gcn_aggregate_fwd_pool_supported': [_c_int, _c_int, _c_int],
    'gpt_gcn_aggregate_bwd_pool': [_p, _p, _p, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _p],
    'gpt_gcn_aggregate_bwd': [_p, _p, _p, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _c_f, _p, _c_int, _p],
    'gpt_gcn_aggregate_bwd_pre': [_p, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _c_int, _p],
    'gpt_gcn_aggregate_bwd_pre_c': [_p, _p, _p, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _c_int, _p],
    'gpt_gcn_aggregate_bwd_pool_c': [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _p],
    'gpt_pool3_bwd_masked': [_p, _p, _p, _p, _p, _c_f, _c_int, _c_int, _c_int, _c_int, _p, _p],
    'gpt_linear_dgrad_tf32x3_masked': [_p, _p, _p, _p, _p, _c_f, _c_int, _c_int, _c_int, _c_int, _p],
    'gpt_pool3_fwd': [_p, _p, _c_int, _c_int, _c_int, _c_int, _p, _p, _p],
    'gpt_pool3_bwd': [_p, _p, _p, _c_int, _c_int, _c_int, _c_int, _p, _p],
    'gpt_linear_fwd_f32': [_p, _p, _p, _c_int, _c_int, _c_int, _p],
    'gpt_linear_dgrad_f32': [_p, _p, _p, _c_int, _c_int, _c_int, _p],
    'gpt_linear_wgrad_f32': [_p, _p, _p, _c_int, _c_int, _c_int, _p],
    'gpt_linear_wgrad_f32_acc': [_p, _p, _p, _c_int, _c_int, _c_int, _p],
    'gpt_linear_wgrad_rows_f32': [_p, _p, _p, _p, _c_int, _c_int, _c_int, _p],
    'gpt_linear_wgrad_tf32x3': [_p, _p, _p, _p, _c_ll, _c_int, _c_int, _p],
    'gpt_gemm_persist_config': [_c_int, _c_ll],
    'gpt_weight_prep_bf16': [_p, _p, _c_int, _c_int, _p],
    'gpt_linear_fwd_bf16': [_p, _p, _p, _c_int, _c_int, _c_int, _p],
    'gpt_linear_dgrad_bf16': [_p, _p, _p, _c_int, _c_int, _c_int, _p],
    'gpt_linear_fwd_tf32': [_p, _p, _p, _c_int, _c_int, _c_int, _p],
    'gpt_linear_dgrad_tf32': [_p, _p, _p, _p, _c_int, _c_int, _c_int, _p],
    'gpt_weight_prep_tf32x3': [_p, _p, _c_int, _c_int, _p],
    'gpt_weight_prep_tf32x3_batch': [_p, _p, _p, _p, _c_int, _p],
    'gpt_linear_fwd_tf32x3': [_p, _p, _p, _c_int, _c_int, _c_int, _p],
    'gpt_linear_dgrad_tf32x3': [_p, _p, _p, _c_int, _c_int, _c_int, _p],
    'gpt_embed_fwd': [_p, _p, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f, _p, _c_u32, _p],
    'gpt_embed_fwd_prep': [_p, _p, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f, _p, _c_u32, _p, _p, _p, _p,
                           _c_int, _p],
    'gpt_embed_bwd': [_p, _p, _p, _p, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f, _p,
                      _c_u32, _p],
    'gpt_embed_bwd_grouped_workspace': [_c_int, _c_int],
    'gpt_embed_bwd_grouped': [_p, _p, _p, _p, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_f, _p,
                              _c_u32, _p, _p],
    'gpt_embed_rows_sqnorm': [_p, _p, _p, _c_int, _c_int, _c_int, _p, _p],
    'gpt_embed_rows_sgd': [_p, _p, _p, _p, _c_int, _c_int, _c_int, _p, _c_f, _c_f, _p],
    'gpt_head_fwd_bwd': [_p, _p, _p, _p, _c_int, _p, _p, _c_int, _c_int, _c_int, _c_f, _c_int, _p, _p, _p, _p, _p, _p,
                         _p],
    'gpt_head_wgrad': [_p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _p, _p, _p, _p, _p, _p],
    'gpt_predict_result_bytes': [_c_int, _c_int],
    'gpt_predict_tail': [_p, _p, _p, _c_int, _c_int, _p, _p],
    'gpt_dp_region_bytes': [_c_int, _c_int, _c_int, _c_int, _c_ll],
    'gpt_dp_partials': [_c_int, _c_int, _c_int, _c_int, _c_ll],
    'gpt_dp_alloc': [_c_ll, _p, _p],
    'gpt_dp_open': [_p, _p],
    'gpt_dp_close': [_p],
    'gpt_dp_free': [_p],
    'gpt_dp_region_init': [_p, _c_int, _c_int, _c_int, _c_int, _c_ll, _p],
    'gpt_dp_push': [_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_ll, _p, _p, _p, _p, _c_int, _c_int, _p],
    'gpt_dp_push_multicast': [_p, _p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_ll, _p, _p, _p, _p, _c_int, _c_int, _p],
    'gpt_dp_signal': [_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_ll, _p],
    'gpt_dp_reduce': [_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_ll, _p, _p, _p],
    'gpt_dp_apply': [_p, _c_int, _c_int, _c_int, _c_int, _c_ll, _p, _p, _p, _p, _c_f, _c_f, _p, _p, _p],
    'gpt_build_batch': [_p, _p, _p, _p, _c_int, _c_int, _c_f, ctypes.c_uint64, ctypes.c_uint64, _p, _p, _p, _p],
    'gpt_relmix_fwd': [_p, _p, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _p, _p, _p, _p],
    'gpt_relmix_bwd': [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _p, _p, _p],
    'gpt_diagmix_fwd': [_p, _p, _p, _p, _c_int, _c_int, _p, _p, _p, _p],
    'gpt_diagmix_bwd': [_p, _p, _p, _p, _p, _p, _p, _c_int, _c_int, _p, _p, _p],
    'gpt_agg3_fwd': [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _c_f, _p, _c_u32, _c_int, _c_int, _c_f, _p, _c_int, _c_int,
                     _c_int, _p, _p],
    'gpt_agg3_bwd': [_p, _p, _p, _p, _p, _p, _p, _p, _p, _c_f, _p, _c_u32, _c_int, _c_int, _c_f, _p, _c_int, _c_int,
                     _c_int, _p, _p, _p, _p],
    'gpt_edge_keep_dense': [_p, _c_int, _c_int, _c_u32, _c_int, _c_f, _p, _p],
    'gpt_relation_keep_tokens': [_p, _c_int, _c_u32, _c_f, _p, _p, _p],
    'gpt_colsum_acc': [_p, _c_ll, _c_int, _p, _p],
    'gpt_live_rows': [_p, _c_int, _p, _p, _p, _p, _p],
    'gpt_gather_rows': [_p, _p, _p, _c_int, _c_int, _p, _p],
    'gpt_scatter_rows': [_p, _p, _c_int, _c_int, _p, _p],
    'gpt_linear_fwd_tf32x3_rows': [_p, _p, _p, _p, _c_int, _c_int, _c_int, _p, _p],
    'gpt_linear_dgrad_tf32x3_rows': [_p, _p, _p, _c_int, _c_int, _c_int, _p, _p],
    'gpt_linear_wgrad_tf32x3_rows': [_p, _p, _p, _p, _c_ll, _c_int, _c_int, _p, _p],
    'gpt_relmix_fwd_rows': [_p, _p, _p, _p, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _p, _p, _p, _p],
    'gpt_relmix_bwd_rows': [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _c_int, _c_int, _c_int, _c_int, _p, _p, _p],
    'gpt_colsum_acc_rows': [_p, _c_ll, _c_int, _p, _p, _p],
    'gpt_update_partials': [_c_ll, _c_int],
    'gpt_update_sqnorm': [_p, _c_ll, _p, _p, _p, _c_int, _c_int, _c_int, _p, _p],
    'gpt_update_apply': [_p, _p, _c_ll, _p, _p, _p, _p, _c_int, _c_int, _c_int, _p, _c_f, _c_f, _c_f, _p, _p, _p],
}

_lib = None


class GptError(RuntimeError):
    pass


def lib():
    """Load (once) and return the ctypes handle; raise loudly when the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GptError('%s not found: build it with `python %s` (nvcc, sm_100a). There is no CPU fallback.'
                           % (LIB_PATH, os.path.join(_HERE, 'csrc', 'build.py')))
        handle = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = _c_int
        handle.gpt_launch_count.restype = ctypes.c_ulonglong
        handle.gpt_dp_region_bytes.restype = ctypes.c_longlong
        handle.gpt_predict_result_bytes.restype = ctypes.c_longlong
        handle.gpt_embed_bwd_grouped_workspace.restype = ctypes.c_longlong
        handle.gpt_error_string.argtypes = [_c_int]
        handle.gpt_error_string.restype = ctypes.c_char_p
        _lib = handle
    return _lib


def check(code, what):
    if code != 0:
        raise GptError('%s failed with code %d: %s' % (what, code, lib().gpt_error_string(code).decode()))
