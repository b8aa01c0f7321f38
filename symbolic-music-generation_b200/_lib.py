"""ctypes binding of `libtxl_b200.so` (the C ABI declared in include/txl_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `build.py`; there is NO fallback: if it is
missing, or an entry point fails, a `TxlError` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libtxl_b200.so')

F32, BF16 = 0, 1
EPI_RELU, EPI_ACCUM, EPI_MASK_POS, EPI_DROPOUT, EPI_BIAS_ROW, EPI_TRANSPOSE, EPI_MASK_SCALE, EPI_EMIT_LIVE, EPI_MASK_LIVE = 1, 2, 4, 8, 16, 32, 64, 128, 256


class TxlError(RuntimeError):
    pass


class TxlBand(C.Structure):
    _fields_ = [('T', C.c_int), ('mlen', C.c_int), ('mem_len', C.c_int), ('clamp_len', C.c_int), ('same_length', C.c_int)]


class TxlEpilogue(C.Structure):
    _fields_ = [('bias', C.c_void_p), ('aux', C.c_void_p), ('colsum', C.c_void_p), ('drop_p', C.c_float),
                ('seed', C.c_uint64), ('site', C.c_uint32), ('flags', C.c_int), ('live_bits', C.c_void_p)]


class TxlAttnDims(C.Structure):
    _fields_ = [('B', C.c_int), ('H', C.c_int), ('dh', C.c_int), ('band', TxlBand),
                ('ldq', C.c_int64), ('ldkv_mem', C.c_int64), ('ldkv_cur', C.c_int64), ('dtype', C.c_int)]


_vp, _i, _i64, _f, _u64, _u32 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64, C.c_uint32

# name -> (restype, argtypes); this table is also what tests/test_cabi.py checks against include/txl_b200.h
SIGNATURES = {
    'txl_version': (_i, []),
    'txl_last_error': (C.c_char_p, []),
    'txl_device_ok': (_i, []),
    'txl_launch_count': (C.c_ulonglong, []),
    'txl_relattn_index_map': (_i, [C.POINTER(TxlBand), _vp, _vp, _vp, _vp, _vp]),
    'txl_embed_fwd': (_i, [_vp, _vp, _vp, _i64, _i, _i, _f, _i, _f, _u64, _u32, _vp]),
    'txl_embed_bwd': (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _f, _i, _f, _u64, _u32, _vp]),
    'txl_posemb_table': (_i, [_vp, _i, _i, _i, _i, _f, _u64, _u32, _vp]),
    'txl_gemm': (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _i64, _i64, _i64, _i, _i, _i, _i, C.POINTER(TxlEpilogue), _vp]),
    'txl_add_ln_fwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, _f, _i, _f, _u64, _u32, _vp]),
    'txl_gemm_add_ln_fwd': (_i, [_vp] * 10 + [_i64] * 5 + [_f, _f, _u64, _u32, _vp, _vp]),
    'txl_add_ln_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i64, _i, _i, _f, _u64, _u32, _vp]),
    'txl_colsum': (_i, [_vp, _i64, _i64, _i64, _i, _vp, _vp]),
    'txl_dropout': (_i, [_vp, _vp, _i64, _i, _f, _u64, _u32, _vp]),
    'txl_relattn_saved_bytes': (_i64, [C.POINTER(TxlAttnDims)]),
    'txl_relattn_fwd': (_i, [_vp] * 11 + [C.POINTER(TxlAttnDims), _vp]),
    'txl_relattn_bwd_probe': (_i, [_i, _i]),
    'txl_relattn_bwd_workspace': (_i64, [C.POINTER(TxlAttnDims)]),
    'txl_relattn_bwd': (_i, [_vp] * 21 + [C.POINTER(TxlAttnDims), _vp]),
    'txl_logsoftmax_nll_fwd': (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp]),
    'txl_logsoftmax_nll_bwd': (_i, [_vp, _i64, _i, _vp, _i64, _i, _vp, _vp, _vp, _i64, _i, _vp]),
    'txl_adaptive_lsm_nll_fwd': (_i, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _i, _i, _vp, _i, _vp]),
    'txl_adaptive_lsm_nll_bwd': (_i, [_vp, _i64, _i, _vp, _i64, _i, _vp, _vp, _vp, _i64, _i, _i, _vp, _vp]),
    'txl_pack_losses': (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'txl_masked_mean': (_i, [_vp, _i64, _vp, _vp, _vp]),
    'txl_ntp_acc': (_i, [_vp, _i64, _vp, _i64, _i, _i, _vp, _vp]),
    'txl_clm_labels': (_i, [_vp, _vp, _i64, _i64, _vp]),
    'txl_shift_labels': (_i, [_vp, _i64, _i, _i, _i64, _i, _vp, _vp, _vp]),
    'txl_last_index_of': (_i, [_vp, _i64, _i, _i, _i64, _vp, _vp]),
    'txl_cast_f32_to_bf16': (_i, [_vp, _vp, _i64, _vp]),
    'txl_cast_bf16_to_f32': (_i, [_vp, _vp, _i64, _vp]),
    'txl_transpose': (_i, [_vp, _vp, _i64, _i64, _i, _vp]),
    'txl_adamw_step': (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _f, _f, _f, _f, _f, _i, _vp, _vp, _vp]),
    'txl_sumsq': (_i, [_vp, _i64, _vp, _vp]),
    'txl_sample': (_i, [_vp, _i, _i, _i, _f, _i, _f, _vp, _vp, _vp, _vp, _vp]),
    'txl_decode_cache_init': (_i, [_vp, _i64, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'txl_decode_attn': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'txl_skinny_gemm': (_i, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    'txl_decode_uniform': (_i, [_vp, _i, _u64, _i64, _vp, _vp]),
    'txl_decode_commit': (_i, [_vp, _vp, _vp, _vp, _i64, _i, _vp, _i, _i64, _i64, _i, _vp]),
    'txl_dec_linear': (_i, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _vp, _i64, _vp]),
    'txl_dec_add_ln': (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _i, _f, _vp, _i64, _vp]),
    'txl_decode_rtab_head_major': (_i, [_vp, _vp, _i, _i, _i, _vp]),
    'txl_decode_attn_pipe_ws_bytes': (_i64, [_i, _i, _i, _i]),
    'txl_decode_cache_init_kv': (_i, [_vp, _i64, _vp, _i, _i, _i, _i, _vp]),
    'txl_decode_attn_pipe': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    'txl_decode_persist_set_timestamps': (_i, [_vp]),
    'txl_decode_persist_supported': (_i, [_i] * 8),
    'txl_decode_persist_ws_bytes': (_i64, [_i] * 8),
    'txl_decode_persist_step': (_i, [_vp] * 15 + [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i] + [_i] * 8 + [_f, _vp]),
    'txl_decode_cluster_set_timestamps': (_i, [_vp]),
    'txl_decode_cluster_max_clusters': (_i, [_i]),
    'txl_decode_cluster_supported': (_i, [_i] * 8),
    'txl_decode_cluster_ws_bytes': (_i64, [_i] * 8),
    'txl_decode_cluster_step': (_i, [_vp] * 15 + [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i] + [_i] * 8 + [_f, _vp]),
    'txl_decode_tail': (_i, [_vp, _i64, _vp, _i, _i, _i, _f, _i, _f, _u64, _i64, _vp, _vp, _vp, _i64, _i, _vp, _vp, _i64, _i64, _i, _vp, _vp, _i, _f, _vp]),
    'txl_set_pdl': (_i, [_i]),
    'txl_decode_attn_pipe_config': (_i, [_i]),
    'txl_tm_to_bm': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'txl_bm_to_tm': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
}

_lib = None


def load():
    """Load the shared library (once) and attach signatures.  Raises TxlError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TxlError(f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                       f'(there is no CPU or PyTorch fallback for this path)')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here means header and library drifted apart
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int, what: str = ''):
    if rc != 0:
        msg = load().txl_last_error().decode('utf-8', 'replace')
        raise TxlError(f'{what} failed (rc={rc}): {msg}')


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise TxlError(f'unsupported dtype {dt}')


def ptr(t):
    """Device pointer of a tensor (None -> NULL).  Tensors must be CUDA and dense in their last dim."""
    if t is None:
        return None
    if not t.is_cuda:
        raise TxlError('the B200 path takes CUDA tensors only (no CPU fallback)')
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream
