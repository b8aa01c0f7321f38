"""Tensor-level wrappers over the C ABI (include/txl_b200.h).  PyTorch is used only to own device memory and
streams; every arithmetic op here is one call into `libtxl_b200.so`.  Nothing in this module falls back to
torch math: a missing library or a non-CUDA tensor raises `TxlError`.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L
from ._lib import TxlAttnDims, TxlBand, TxlEpilogue, TxlError, check, dtype_code, ptr, stream_ptr


def _lib():
    return L.load()


def device_ok():
    check(_lib().txl_device_ok(), 'txl_device_ok')


def make_band(T, mlen, mem_len, clamp_len, same_length) -> TxlBand:
    return TxlBand(int(T), int(mlen), int(mem_len), int(clamp_len), int(bool(same_length)))


def num_r(T, mlen, clamp_len=0) -> int:
    """rows of the r table = klen (HF's r_head_k; the clamp is baked into the position table)."""
    return mlen + T


# ----------------------------------------------------------------------------- index maps
def relattn_index_map(T, mlen, mem_len, clamp_len, same_length, device='cuda'):
    klen = T + mlen
    masked = torch.empty(T, klen, dtype=torch.uint8, device=device)
    ridx = torch.empty(T, klen, dtype=torch.int32, device=device)
    lo = torch.empty(T, dtype=torch.int32, device=device)
    hi = torch.empty(T, dtype=torch.int32, device=device)
    band = make_band(T, mlen, mem_len, clamp_len, same_length)
    check(_lib().txl_relattn_index_map(C.byref(band), ptr(masked), ptr(ridx), ptr(lo), ptr(hi), stream_ptr()), 'relattn_index_map')
    return masked, ridx, lo, hi


# ----------------------------------------------------------------------------- embedding / positions
def embed_fwd(ids, E, scale, drop_p=0.0, seed=0, site=0):
    n_tok, (V, d) = ids.numel(), E.shape
    out = torch.empty(n_tok, d, dtype=E.dtype, device=E.device)
    check(_lib().txl_embed_fwd(ptr(ids), ptr(E), ptr(out), n_tok, d, V, float(scale), dtype_code(E.dtype),
                               float(drop_p), int(seed), int(site), stream_ptr()), 'embed_fwd')
    return out


def embed_bwd(ids, dOut, dE, scale, drop_p=0.0, seed=0, site=0, dOut2=None):
    V, d = dE.shape
    assert dE.dtype == torch.float32
    check(_lib().txl_embed_bwd(ptr(ids), ptr(dOut), ptr(dOut2), ptr(dE), ids.numel(), d, V, float(scale), dtype_code(dOut.dtype),
                               float(drop_p), int(seed), int(site), stream_ptr()), 'embed_bwd')


def posemb_table(klen, clamp_len, d, dtype, device, drop_p=0.0, seed=0, site=0):
    """HF pos_emb: row x <-> position min(klen-1-x, clamp_len)."""
    out = torch.empty(klen, d, dtype=dtype, device=device)
    check(_lib().txl_posemb_table(ptr(out), klen, int(clamp_len), d, dtype_code(dtype), float(drop_p), int(seed), int(site), stream_ptr()), 'posemb_table')
    return out


# ----------------------------------------------------------------------------- GEMM
def gemm(A, B, *, transA=False, transB=False, out=None, out_dtype=None, bias=None, relu=False, accumulate=False,
         mask_pos_aux=None, colsum=None, drop_p=0.0, seed=0, site=0, M=None, N=None, K=None, bias_row=False, transpose_out=False,
         aux_is_dropped=False, emit_live_bits=None, mask_live_bits=None):
    """C = epi(op(A) op(B)).  A, B are 2-D row-major tensors (possibly column-sliced views: stride(0) is the ld)."""
    assert A.dim() == 2 and B.dim() == 2 and A.stride(1) == 1 and B.stride(1) == 1
    if M is None:
        M = A.shape[1] if transA else A.shape[0]
    if K is None:
        K = A.shape[0] if transA else A.shape[1]
    if N is None:
        N = B.shape[0] if transB else B.shape[1]
    if out is None:
        out = torch.empty((N, M) if transpose_out else (M, N), dtype=out_dtype or A.dtype, device=A.device)
    assert out.stride(1) == 1 and A.dtype == B.dtype
    flags = (L.EPI_RELU if relu else 0) | (L.EPI_ACCUM if accumulate else 0) | (L.EPI_BIAS_ROW if bias_row else 0) | (L.EPI_TRANSPOSE if transpose_out else 0)
    if mask_pos_aux is not None:
        flags |= L.EPI_MASK_POS
        assert mask_pos_aux.dtype == out.dtype and mask_pos_aux.stride(0) == out.stride(0)
    bits = None
    if emit_live_bits is not None:       # forward of relu(+dropout): also write the 1-bit-per-element backward mask
        bits, flags = emit_live_bits, flags | L.EPI_EMIT_LIVE
    if mask_live_bits is not None:       # backward: mask from those bits (they already carry the dropout mask of the site)
        assert mask_pos_aux is None and emit_live_bits is None
        bits, flags = mask_live_bits, flags | L.EPI_MASK_LIVE
    if bits is not None:
        assert bits.dtype == torch.int32 and bits.is_contiguous() and bits.numel() >= ((N + 31) // 32) * M
    if drop_p > 0.0:
        # aux_is_dropped: aux is the forward's post-dropout activation, so (aux > 0) already carries this site's dropout mask
        masked = mask_pos_aux is not None or mask_live_bits is not None
        flags |= L.EPI_MASK_SCALE if (masked and (aux_is_dropped or mask_live_bits is not None)) else L.EPI_DROPOUT
    epi = TxlEpilogue(ptr(bias), ptr(mask_pos_aux), ptr(colsum), float(drop_p), int(seed), int(site), flags, ptr(bits))
    check(_lib().txl_gemm(ptr(A), ptr(B), ptr(out), M, N, K, A.stride(0), B.stride(0), out.stride(0), int(transA), int(transB),
                          dtype_code(A.dtype), dtype_code(out.dtype), C.byref(epi), stream_ptr()), 'gemm')
    return out


# ----------------------------------------------------------------------------- residual + LayerNorm
def add_ln_fwd(x, r, gamma, beta, eps, drop_p=0.0, seed=0, site=0, save=True):
    rows, d = x.shape
    y = torch.empty_like(x)
    z = torch.empty_like(x) if save else None
    mean = torch.empty(rows, dtype=torch.float32, device=x.device) if save else None
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if save else None
    check(_lib().txl_add_ln_fwd(ptr(x), ptr(r), ptr(gamma), ptr(beta), ptr(y), ptr(z), ptr(mean), ptr(rstd), rows, d, float(eps),
                                dtype_code(x.dtype), float(drop_p), int(seed), int(site), stream_ptr()), 'add_ln_fwd')
    return y, z, mean, rstd


def gemm_add_ln_fwd(A, W, bias, resid, gamma, beta, eps, drop_p=0.0, seed=0, site=0, save=True):
    """y = LN(resid + dropout(A W^T + bias)): one tensor-core kernel when the library covers the shape (bf16, N in {256, 512}), else the
    GEMM followed by the residual + LayerNorm kernel.  Returns (y, z, mean, rstd) exactly like add_ln_fwd."""
    M, K = A.shape
    N = W.shape[0]
    if A.dtype == torch.bfloat16 and A.stride(1) == 1 and W.stride(1) == 1 and resid.is_contiguous():
        y = torch.empty(M, N, dtype=A.dtype, device=A.device)
        z = torch.empty_like(y) if save else None
        mean = torch.empty(M, dtype=torch.float32, device=A.device) if save else None
        rstd = torch.empty(M, dtype=torch.float32, device=A.device) if save else None
        handled = C.c_int(0)
        check(_lib().txl_gemm_add_ln_fwd(ptr(A), ptr(W), ptr(bias), ptr(resid), ptr(gamma), ptr(beta), ptr(y), ptr(z), ptr(mean), ptr(rstd), M, N, K,
                                         A.stride(0), W.stride(0), float(eps), float(drop_p), int(seed), int(site), stream_ptr(), C.byref(handled)),
              'gemm_add_ln_fwd')
        if handled.value:
            return y, z, mean, rstd
    r = gemm(A, W, transB=True, bias=bias)
    return add_ln_fwd(resid, r, gamma, beta, eps, drop_p, seed, site, save)


def add_ln_bwd(dy, z, gamma, mean, rstd, dgamma, dbeta, dx_out=None, accumulate_dx=False, want_dr=True, drop_p=0.0, seed=0, site=0, dy2=None):
    rows, d = dy.shape
    if dx_out is None:
        dx_out = torch.empty_like(dy)
    dr = torch.empty_like(dy) if want_dr else None
    check(_lib().txl_add_ln_bwd(ptr(dy), ptr(dy2), ptr(z), ptr(gamma), ptr(mean), ptr(rstd), ptr(dx_out), int(accumulate_dx), ptr(dr),
                                ptr(dgamma), ptr(dbeta), rows, d, dtype_code(dy.dtype), float(drop_p), int(seed), int(site),
                                stream_ptr()), 'add_ln_bwd')
    return dx_out, dr


def dropout(x, drop_p, seed, site, out=None):
    if out is None:
        out = torch.empty_like(x)
    check(_lib().txl_dropout(ptr(x), ptr(out), x.numel(), dtype_code(x.dtype), float(drop_p), int(seed), int(site), stream_ptr()), 'dropout')
    return out


def colsum(X, out):
    """out[n] += sum_m X[m, n]  (fp32 accumulate)."""
    assert X.dim() == 2 and X.stride(1) == 1 and out.dtype == torch.float32
    check(_lib().txl_colsum(ptr(X), X.shape[0], X.shape[1], X.stride(0), dtype_code(X.dtype), ptr(out), stream_ptr()), 'colsum')


# ----------------------------------------------------------------------------- attention
def _attn_dims(q, k_mem, k_cur, B, T, H, dh, band):
    return TxlAttnDims(B, H, dh, band, q.stride(0), k_mem.stride(0) if k_mem is not None else 0, k_cur.stride(0), dtype_code(q.dtype))


def relattn_fwd(q, k_mem, v_mem, k_cur, v_cur, r, rwb, rrb, B, T, H, dh, band: TxlBand, save=False):
    """q/k_cur/v_cur: [B*T, >=H*dh] views; k_mem/v_mem: [B*mlen, >=H*dh] views or None; r: [P, H*dh].
    save=True also returns the forward state `relattn_bwd(saved=...)` reuses (None when no kernel would use it)."""
    out = torch.empty(B * T, H * dh, dtype=q.dtype, device=q.device)
    lse = torch.empty(B, H, T, dtype=torch.float32, device=q.device)
    dims = _attn_dims(q, k_mem, k_cur, B, T, H, dh, band)
    saved = None
    if save:
        nbytes = _lib().txl_relattn_saved_bytes(C.byref(dims))
        saved = torch.empty(nbytes, dtype=torch.uint8, device=q.device) if nbytes > 0 else None
    check(_lib().txl_relattn_fwd(ptr(q), ptr(k_mem), ptr(v_mem), ptr(k_cur), ptr(v_cur), ptr(r), ptr(rwb), ptr(rrb), ptr(out), ptr(lse),
                                 ptr(saved), C.byref(dims), stream_ptr()), 'relattn_fwd')
    return (out, lse, saved) if save else (out, lse)


def relattn_bwd(q, k_mem, v_mem, k_cur, v_cur, r, rwb, rrb, out, lse, dout, dq, dk_mem, dv_mem, dk_cur, dv_cur, dr, drwb, drrb,
                B, T, H, dh, band: TxlBand, saved=None):
    dims = _attn_dims(q, k_mem, k_cur, B, T, H, dh, band)
    assert dq.stride(0) == q.stride(0) and dk_cur.stride(0) == k_cur.stride(0)
    assert dk_mem is None or dk_mem.stride(0) == k_mem.stride(0)
    nbytes = _lib().txl_relattn_bwd_workspace(C.byref(dims))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=q.device)
    check(_lib().txl_relattn_bwd(ptr(q), ptr(k_mem), ptr(v_mem), ptr(k_cur), ptr(v_cur), ptr(r), ptr(rwb), ptr(rrb), ptr(out), ptr(lse),
                                 ptr(dout), ptr(dq), ptr(dk_mem), ptr(dv_mem), ptr(dk_cur), ptr(dv_cur), ptr(dr), ptr(drwb), ptr(drrb),
                                 ptr(ws), ptr(saved), C.byref(dims), stream_ptr()), 'relattn_bwd')


# ----------------------------------------------------------------------------- LM head
def logsoftmax_nll_fwd(logits, V, labels=None, want_logprobs=False, want_argmax=False):
    N, ldl = logits.shape[0], logits.stride(0)
    dev = logits.device
    losses = torch.empty(N, dtype=torch.float32, device=dev) if labels is not None else None
    lse = torch.empty(N, dtype=torch.float32, device=dev)
    logprobs = torch.empty(N, V, dtype=torch.float32, device=dev) if want_logprobs else None
    argmax = torch.empty(N, dtype=torch.int64, device=dev) if want_argmax else None
    check(_lib().txl_logsoftmax_nll_fwd(ptr(logits), ldl, ptr(labels), ptr(losses), ptr(lse), ptr(logprobs), ptr(argmax), N, V,
                                        dtype_code(logits.dtype), stream_ptr()), 'logsoftmax_nll_fwd')
    return losses, lse, logprobs, argmax


def logsoftmax_nll_bwd(logits, V, labels, lse, grow, out_dtype=None):
    """dlogits; in place when out_dtype is None or equals logits.dtype, else a new buffer of out_dtype with the same pitch."""
    out = logits if out_dtype in (None, logits.dtype) else torch.empty(logits.shape[0], logits.stride(0), dtype=out_dtype, device=logits.device)
    check(_lib().txl_logsoftmax_nll_bwd(ptr(logits), logits.stride(0), dtype_code(logits.dtype), ptr(out), out.stride(0), dtype_code(out.dtype),
                                        ptr(labels), ptr(lse), ptr(grow), logits.shape[0], V, stream_ptr()), 'logsoftmax_nll_bwd')
    return out


def _cuts(cutoffs):
    arr = (C.c_int * len(cutoffs))(*[int(c) for c in cutoffs])
    return arr


def adaptive_lsm_nll_fwd(logits, V, cutoffs, labels=None, want_logprobs=False, want_argmax=False):
    """Cluster path of HF's ProjectedAdaptiveLogSoftmax: logits (N, >= V + n_clusters) = token logits then cluster logits."""
    N, nc, dev = logits.shape[0], len(cutoffs), logits.device
    losses = torch.empty(N, dtype=torch.float32, device=dev) if labels is not None else None
    lse = torch.empty(N, nc + 1, dtype=torch.float32, device=dev)
    logprobs = torch.empty(N, V, dtype=torch.float32, device=dev) if want_logprobs else None
    argmax = torch.empty(N, dtype=torch.int64, device=dev) if want_argmax else None
    check(_lib().txl_adaptive_lsm_nll_fwd(ptr(logits), logits.stride(0), ptr(labels), ptr(losses), ptr(lse), ptr(logprobs), ptr(argmax), N, V, nc,
                                          _cuts(cutoffs), dtype_code(logits.dtype), stream_ptr()), 'adaptive_lsm_nll_fwd')
    return losses, lse, logprobs, argmax


def adaptive_lsm_nll_bwd(logits, V, cutoffs, labels, lse, grow, out_dtype=None):
    out = logits if out_dtype in (None, logits.dtype) else torch.empty(logits.shape[0], logits.stride(0), dtype=out_dtype, device=logits.device)
    check(_lib().txl_adaptive_lsm_nll_bwd(ptr(logits), logits.stride(0), dtype_code(logits.dtype), ptr(out), out.stride(0), dtype_code(out.dtype),
                                          ptr(labels), ptr(lse), ptr(grow), logits.shape[0], V, len(cutoffs), _cuts(cutoffs), stream_ptr()),
          'adaptive_lsm_nll_bwd')
    return out


def pack_losses(pos_losses, labels_shift, B, T, V, cutoffs):
    """HF keep_order=False ordering of the returned loss vector: (packed (B, T-1) fp32, perm (B*(T-1),) int64 of source rows b*T+t / -1)."""
    packed = torch.empty(B, T - 1, dtype=torch.float32, device=pos_losses.device)
    perm = torch.empty(B * (T - 1), dtype=torch.int64, device=pos_losses.device)
    check(_lib().txl_pack_losses(ptr(pos_losses), ptr(labels_shift), B, T, V, len(cutoffs), _cuts(cutoffs), ptr(packed), ptr(perm), stream_ptr()),
          'pack_losses')
    return packed, perm


def ntp_acc(preds, labels, out=None):
    """(matches, non-pad count) of next-token prediction, accumulated into the int64[2] device tensor `out` (created zeroed if None)."""
    assert preds.dtype == torch.int64 and labels.dtype == torch.int64 and preds.shape == labels.shape and preds.dim() == 2
    assert preds.stride(1) == 1 and labels.stride(1) == 1
    if out is None:
        out = torch.zeros(2, dtype=torch.int64, device=preds.device)
    check(_lib().txl_ntp_acc(ptr(preds), preds.stride(0), ptr(labels), labels.stride(0), preds.shape[0], preds.shape[1], ptr(out), stream_ptr()), 'ntp_acc')
    return out


def shift_labels(labels, eos, V, bad=None):
    """labels (B, T) int64 CUDA, last dim dense -> (B*T,) shifted labels for the LM-head kernels; applies the reference's all-pad first-row
    fix-up to `labels` in place on the device (transformer_xl.py:176-182) — no host round trip."""
    assert labels.dtype == torch.int64 and labels.dim() == 2 and labels.stride(1) == 1
    B, T = labels.shape
    out = torch.empty(B * T, dtype=torch.int64, device=labels.device)
    check(_lib().txl_shift_labels(ptr(labels), labels.stride(0), B, T, int(eos), int(V), ptr(out), ptr(bad), stream_ptr()), 'shift_labels')
    return out


def masked_mean(losses):
    out = torch.empty(2, dtype=torch.float32, device=losses.device)
    check(_lib().txl_masked_mean(ptr(losses), losses.numel(), ptr(out[0:1]), ptr(out[1:2]), stream_ptr()), 'masked_mean')
    return out[0], out[1]


# ----------------------------------------------------------------------------- parameters / optimiser
def cast_f32_to_bf16(src, dst):
    check(_lib().txl_cast_f32_to_bf16(ptr(src), ptr(dst), src.numel(), stream_ptr()), 'cast_f32_to_bf16')
    return dst


def cast_bf16_to_f32(src, dst):
    check(_lib().txl_cast_bf16_to_f32(ptr(src), ptr(dst), src.numel(), stream_ptr()), 'cast_bf16_to_f32')
    return dst


def transpose(A, out=None):
    rows, cols = A.shape
    if out is None:
        out = torch.empty(cols, rows, dtype=A.dtype, device=A.device)
    check(_lib().txl_transpose(ptr(A), ptr(out), rows, cols, dtype_code(A.dtype), stream_ptr()), 'transpose')
    return out


def adamw_step(p, g, m, v, decay_mask, lr, beta1, beta2, eps, weight_decay, step, grad_scale=None, bf16_shadow=None):
    check(_lib().txl_adamw_step(ptr(p), ptr(g), ptr(m), ptr(v), ptr(decay_mask), p.numel(), float(lr), float(beta1), float(beta2),
                                float(eps), float(weight_decay), int(step), ptr(grad_scale), ptr(bf16_shadow), stream_ptr()), 'adamw_step')


def sumsq(g, out):
    check(_lib().txl_sumsq(ptr(g), g.numel(), ptr(out), stream_ptr()), 'sumsq')


# ----------------------------------------------------------------------------- sampling / layout
def sample(scores, do_sample, temperature=1.0, top_k=0, top_p=1.0, u=None, want_keep=False, want_warped=False):
    B, V = scores.shape
    assert scores.dtype == torch.float32 and scores.is_contiguous()
    nxt = torch.empty(B, dtype=torch.int64, device=scores.device)
    keep = torch.empty(B, V, dtype=torch.uint8, device=scores.device) if want_keep else None
    warped = torch.empty(B, V, dtype=torch.float32, device=scores.device) if want_warped else None
    check(_lib().txl_sample(ptr(scores), B, V, int(bool(do_sample)), float(temperature), int(top_k or 0), float(top_p), ptr(u), ptr(nxt),
                            ptr(keep), ptr(warped), stream_ptr()), 'sample')
    return nxt, keep, warped


def tm_to_bm(src, dst_dtype):
    rows, B, d = src.shape
    dst = torch.empty(B, rows, d, dtype=dst_dtype, device=src.device)
    check(_lib().txl_tm_to_bm(ptr(src.contiguous()), ptr(dst), rows, B, d, dtype_code(src.dtype), dtype_code(dst_dtype), stream_ptr()), 'tm_to_bm')
    return dst


def bm_to_tm(src, dst_dtype):
    B, rows, d = src.shape
    dst = torch.empty(rows, B, d, dtype=dst_dtype, device=src.device)
    check(_lib().txl_bm_to_tm(ptr(src.contiguous()), ptr(dst), rows, B, d, dtype_code(src.dtype), dtype_code(dst_dtype), stream_ptr()), 'bm_to_tm')
    return dst
