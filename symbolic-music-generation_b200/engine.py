"""Forward / backward orchestration of the Transformer-XL path on top of the C-ABI ops.

Restates HF `TransfoXLModel.forward` + `RelPartialLearnableDecoderLayer` + `ProjectedAdaptiveLogSoftmax`
 in batch-major layout, with a hand-scheduled backward (no autograd inside):
the whole step is one `torch.autograd.Function`, so `loss.backward()` works for the reference's HF-Trainer caller
(musicnlp/util/train/train_util_wrap.py:88-144) while the kernels, buffers and the gradient all-reduce order are ours.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import torch

from . import ops

# dropout sites (one counter-based stream each; layer-local sites are offset by 8*layer)
SITE_EMB, SITE_POS, SITE_FINAL = 0, 1, 2
SITE_LAYER0 = 8
S_ATTN_OUT, S_FF_INNER, S_FF_OUT = 0, 1, 2


@dataclass
class LayerW:
    """Per-layer parameter views (compute dtype for matrices, fp32 for vectors) + fp32 grad views."""
    qkv: torch.Tensor = None      # (3d, d)
    r: torch.Tensor = None        # (d, d)
    o: torch.Tensor = None        # (d, d)
    rrb: torch.Tensor = None      # (H*dh) fp32
    rwb: torch.Tensor = None
    ln1_w: torch.Tensor = None
    ln1_b: torch.Tensor = None
    w1: torch.Tensor = None       # (di, d)
    b1: torch.Tensor = None
    w2: torch.Tensor = None       # (d, di)
    b2: torch.Tensor = None
    ln2_w: torch.Tensor = None
    ln2_b: torch.Tensor = None


@dataclass
class Saved:
    layers: list = field(default_factory=list)
    ids: torch.Tensor = None
    labels: torch.Tensor = None
    pos: torch.Tensor = None
    core_in: torch.Tensor = None      # input of the final dropout (= last layer output)
    core: torch.Tensor = None
    logits: torch.Tensor = None
    lse: torch.Tensor = None
    losses: torch.Tensor = None
    count: torch.Tensor = None
    seed: int = 0
    drop_p: float = 0.0
    mems_real: bool = False
    mems: list = None
    B: int = 0
    T: int = 0
    mlen: int = 0


class Geometry:
    def __init__(self, cfg, B, T, mlen):
        self.B, self.T, self.mlen = B, T, mlen
        self.d, self.H, self.dh, self.di, self.V = cfg.d_model, cfg.n_head, cfg.d_head, cfg.d_inner, cfg.vocab_size
        self.L = cfg.n_layer
        self.N = B * T
        self.band = ops.make_band(T, mlen, cfg.mem_len, cfg.clamp_len, cfg.same_length)
        self.P = ops.num_r(T, mlen, cfg.clamp_len)
        self.cutoffs = list(getattr(cfg, 'cutoffs', []) or [])
        self.Vx = self.V + len(self.cutoffs)          # LM-head columns: token logits, then one logit per adaptive-softmax cluster
        self.Vp = (self.Vx + 7) // 8 * 8


def forward(cfg, W: List[LayerW], E, out_bias, ids, mems_bm, labels_shift, *, drop_p, seed, save: bool,
            want_logprobs: bool, want_argmax: bool = False, zero_kvm=None):
    """ids (B,T) int64; mems_bm: list of contiguous (B, mlen, d) tensors in compute dtype or None (=> zero mems of
    length cfg.mem_len, HF `init_mems`); labels_shift (B*T,) int64 (labels moved one step left, -100 in the last column) or None.
    Returns dict(losses, loss, count, logprobs, argmax, hid_in (list), saved)."""
    B, T = ids.shape
    mems_real = mems_bm is not None
    mlen = mems_bm[0].shape[1] if mems_real else (cfg.mem_len if cfg.mem_len > 0 else 0)
    g = Geometry(cfg, B, T, mlen)
    d, H, dh, N = g.d, g.H, g.dh, g.N
    dt, dev = E.dtype, E.device
    sv = Saved(ids=ids, labels=labels_shift, seed=seed, drop_p=drop_p, mems_real=mems_real, B=B, T=T, mlen=mlen) if save else None

    x = ops.embed_fwd(ids.reshape(-1), E[:g.V], math.sqrt(d), drop_p, seed, SITE_EMB)     # E = [embedding ; cluster_weight] (V + n_clusters, d)
    pos = ops.posemb_table(g.P, cfg.clamp_len, d, dt, dev, drop_p, seed, SITE_POS)
    if save:
        sv.pos = pos
        sv.mems = mems_bm
    kvm_zero = None
    if not mems_real and mlen > 0:
        kvm_zero = zero_kvm(B * mlen, 2 * d, dt, dev) if zero_kvm else torch.zeros(B * mlen, 2 * d, dtype=dt, device=dev)
    hid_in = []
    for li, w in enumerate(W):
        site = SITE_LAYER0 + 8 * li
        hid_in.append(x)
        qkv = ops.gemm(x, w.qkv, transB=True)                                  # (N, 3d)
        if mems_real and mlen > 0:
            kvm = ops.gemm(mems_bm[li].reshape(B * mlen, d), w.qkv[d:], transB=True)   # (B*mlen, 2d)
        else:
            kvm = kvm_zero
        r = ops.gemm(pos, w.r, transB=True)                                    # (P, d)
        k_mem = kvm[:, :d] if kvm is not None else None
        v_mem = kvm[:, d:] if kvm is not None else None
        att = ops.relattn_fwd(qkv[:, :d], k_mem, v_mem, qkv[:, d:2 * d], qkv[:, 2 * d:], r, w.rwb, w.rrb, B, T, H, dh, g.band, save=bool(save))
        vec, lse, att_saved = att if save else (att[0], att[1], None)
        # o_net + dropout + residual + LayerNorm: one kernel (GEMM epilogue holds whole rows); falls back to GEMM + add_ln for other shapes
        y1, z1, mean1, rstd1 = ops.gemm_add_ln_fwd(vec, w.o, None, x, w.ln1_w, w.ln1_b, cfg.layer_norm_epsilon, drop_p, seed, site + S_ATTN_OUT, save)
        # the backward mask of dropout(relu(.)) at one bit per element, written by the same epilogue ([di/32, N] words)
        hbits = torch.empty((cfg.d_inner + 31) // 32, N, dtype=torch.int32, device=dev) if save else None
        h = ops.gemm(y1, w.w1, transB=True, bias=w.b1, relu=True, drop_p=drop_p, seed=seed, site=site + S_FF_INNER, emit_live_bits=hbits)
        y2, z2, mean2, rstd2 = ops.gemm_add_ln_fwd(h, w.w2, w.b2, y1, w.ln2_w, w.ln2_b, cfg.layer_norm_epsilon, drop_p, seed, site + S_FF_OUT, save)
        if save:
            sv.layers.append(dict(x=x, qkv=qkv, kvm=kvm if mems_real else None, kvm_fwd=kvm, r=r, vec=vec, lse=lse, att_saved=att_saved, z1=z1, mean1=mean1,
                                  rstd1=rstd1, y1=y1, h=h, hbits=hbits, z2=z2, mean2=mean2, rstd2=rstd2))
        x = y2
    core = ops.dropout(x, drop_p, seed, SITE_FINAL) if drop_p > 0 else x
    # LM head: logits for every position; label shifting is done by the caller (labels_shift)
    # logits stay fp32 (the log-softmax / NLL reads them once; bf16 logits would cost ~1.5e-2 absolute on every log-prob)
    logits = torch.empty(N, g.Vp, dtype=torch.float32, device=dev)
    if g.Vp != g.Vx:
        logits[:, g.Vx:].zero_()
    ops.gemm(core, E, transB=True, bias=out_bias, out=logits, N=g.Vx)
    if g.cutoffs:      # adaptive softmax: head over the shortlist + cluster logits, one tail per cluster (A.6, cluster path)
        losses, lse_v, logprobs, argmax = ops.adaptive_lsm_nll_fwd(logits, g.V, g.cutoffs, labels_shift, want_logprobs, want_argmax)
    else:
        losses, lse_v, logprobs, argmax = ops.logsoftmax_nll_fwd(logits, g.V, labels_shift, want_logprobs, want_argmax)
    loss = count = None
    if labels_shift is not None:
        loss, count = ops.masked_mean(losses)
    if save:
        sv.core, sv.logits, sv.lse, sv.losses, sv.count = core, logits, lse_v, losses, count
    return dict(losses=losses, loss=loss, count=count, logprobs=logprobs, argmax=argmax, hid_in=hid_in, saved=sv, geom=g)


def backward(cfg, W: List[LayerW], G: List[LayerW], E, gE, g_out_bias, sv: Saved, grow, on_layer_done=None):
    """Hand-scheduled backward.  `G` mirrors `W` with fp32 gradient views that are ACCUMULATED into.
    grow (N,) fp32 = d loss / d losses[n].  on_layer_done(li) is called once layer li's parameter gradients are final
    (reverse order; used to launch the bucketed gradient all-reduce while earlier layers are still computing)."""
    B, T, mlen = sv.B, sv.T, sv.mlen
    g = Geometry(cfg, B, T, mlen)
    d, H, dh, N = g.d, g.H, g.dh, g.N
    dt = E.dtype
    p, seed = sv.drop_p, sv.seed
    # ---- LM head
    if g.cutoffs:
        dlogits = ops.adaptive_lsm_nll_bwd(sv.logits, g.V, g.cutoffs, sv.labels, sv.lse, grow, out_dtype=dt)
    else:
        dlogits = ops.logsoftmax_nll_bwd(sv.logits, g.V, sv.labels, sv.lse, grow, out_dtype=dt)       # (N, Vp) in the compute dtype
    ops.colsum(dlogits[:, :g.Vx], g_out_bias)
    dl = dlogits[:, :g.Vx]
    ops.gemm(dl, sv.core, transA=True, out=gE, accumulate=True)                                       # dE += dlogits^T core
    dcore = ops.gemm(dl, E)                                                                           # (N, d)
    sv.logits = None
    dx = ops.dropout(dcore, p, seed, SITE_FINAL, out=dcore) if p > 0 else dcore
    dx2 = None      # second branch meeting dx at the residual node (summed on read by the next LayerNorm backward)
    for li in range(len(W) - 1, -1, -1):
        w, gw, s = W[li], G[li], sv.layers[li]
        site = SITE_LAYER0 + 8 * li
        # LN2 / FF
        dy1, df = ops.add_ln_bwd(dx, s['z2'], w.ln2_w, s['mean2'], s['rstd2'], gw.ln2_w, gw.ln2_b, drop_p=p, seed=seed, site=site + S_FF_OUT, dy2=dx2)
        ops.colsum(df, gw.b2)
        ops.gemm(df, s['h'], transA=True, out=gw.w2, accumulate=True)                                 # dW2 += df^T h
        dh_ = ops.gemm(df, w.w2, mask_live_bits=s['hbits'], colsum=gw.b1, drop_p=p, seed=seed, site=site + S_FF_INNER)   # (N, di)
        ops.gemm(dh_, s['y1'], transA=True, out=gw.w1, accumulate=True)                               # dW1 += dh^T y1
        dff = ops.gemm(dh_, w.w1)                                                                     # dh W1  (N, d)
        del dh_, df
        # LN1 / o_net
        dxn, dao = ops.add_ln_bwd(dy1, s['z1'], w.ln1_w, s['mean1'], s['rstd1'], gw.ln1_w, gw.ln1_b, drop_p=p, seed=seed, site=site + S_ATTN_OUT, dy2=dff)
        ops.gemm(dao, s['vec'], transA=True, out=gw.o, accumulate=True)                               # dWo += dao^T vec
        dvec = ops.gemm(dao, w.o)                                                                     # (N, d)
        # attention
        qkv, kvm = s['qkv'], s['kvm_fwd']
        dqkv = torch.empty_like(qkv)
        dkvm = torch.empty_like(kvm) if s['kvm'] is not None else None
        dr = torch.zeros(g.P, d, dtype=torch.float32, device=qkv.device)
        k_mem = kvm[:, :d] if kvm is not None else None
        v_mem = kvm[:, d:] if kvm is not None else None
        ops.relattn_bwd(qkv[:, :d], k_mem, v_mem, qkv[:, d:2 * d], qkv[:, 2 * d:], s['r'], w.rwb, w.rrb, s['vec'], s['lse'], dvec,
                        dqkv[:, :d], dkvm[:, :d] if dkvm is not None else None, dkvm[:, d:] if dkvm is not None else None,
                        dqkv[:, d:2 * d], dqkv[:, 2 * d:], dr, gw.rwb, gw.rrb, B, T, H, dh, g.band, saved=s['att_saved'])
        # r_net:  r = pos Wr^T
        dr_c = dr if dt == torch.float32 else ops.cast_f32_to_bf16(dr, torch.empty_like(dr, dtype=dt))
        ops.gemm(dr_c, sv.pos, transA=True, out=gw.r, accumulate=True)
        # qkv_net
        ops.gemm(dqkv, s['x'], transA=True, out=gw.qkv, accumulate=True)                              # dWqkv += dqkv^T x
        dx2 = ops.gemm(dqkv, w.qkv)                                                                   # dqkv Wqkv  (N, d)
        if dkvm is not None:
            ops.gemm(dkvm, sv.mems[li].reshape(B * mlen, d), transA=True, out=gw.qkv[d:], accumulate=True)
        dx = dxn
        sv.layers[li] = None
        if on_layer_done is not None:
            on_layer_done(li)
    ops.embed_bwd(sv.ids.reshape(-1), dx, gE[:g.V], math.sqrt(d), p, seed, SITE_EMB, dOut2=dx2)
