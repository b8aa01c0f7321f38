"""Mems-cached decode step for `generate`: one new token per sequence per step, projected-K/V ring cache, every per-step op a C-ABI
kernel, the whole step replayed as a CUDA graph (no host round trip between tokens).

Follows HF 4.25 `sample` / `greedy_search` step for step: embed last token -> L x (qkv, relative-position attention
over [mems ; current] with the same_length band, o_net + LN, FF + LN) -> log-softmax -> warpers -> draw -> eos/pad bookkeeping -> append.

Three engines share this file:
  * bf16, persistent (the default where its geometry applies: d_head 64, d_model 128 / 512, at most 64 sequences per chain): ONE cooperative
    kernel per step for all layers + the LM-head GEMM (`txl_decode_persist_step`, csrc/decode_persist.cu) over a ring of cached HIDDEN
    states - HF's `mems` themselves, half the bytes of a projected k|v cache - with the key / value projections absorbed into the query /
    output side; grid barriers instead of 84 kernel boundaries.  Then `txl_decode_tail`.  `TXL_DECODE_PERSIST=0` selects the next one.
  * bf16 (the measured path): `txl_dec_linear` / `txl_dec_add_ln` / `txl_decode_attn_pipe` / `txl_decode_tail` (csrc/decode_stream.cu,
    csrc/sample.cu) over an interleaved k|v ring, launched with programmatic stream serialization; `GroupedDecoder` captures groups of 16
    sequences as parallel branches of one graph.  74 launches per step and group.
  * fp32 parity mode (token-identical greedy decode against the oracle): the exact-FMA kernels of csrc/decode.cu.
Environment switches (A/B measurements, see profiles/r01_decode_ab.txt; defaults are the measured best): TXL_DECODE_PDL,
TXL_DECODE_TAIL, TXL_DECODE_GROUPS, TXL_DECODE_ATTN_SPLITS, TXL_DECODE_ATTN_CFG, TXL_DECODE_PREFETCH_MB, TXL_DEC_SPLIT_COLS,
TXL_DEC_LINEAR_2STAGE, TXL_DECODE_ABL (timing ablation: results are garbage).
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch

from . import ops
from ._lib import check, dtype_code, load, ptr, stream_ptr

SK_MAXM = 64


def supported(model, B):
    """Any batch size: more than SK_MAXM sequences are decoded as several sequence groups (`sequence_groups`)."""
    cfg = model.config
    return B >= 1 and cfg.same_length and cfg.mem_len > 0 and cfg.d_head in (32, 64, 128) and cfg.d_model % 8 == 0 and cfg.d_inner % 8 == 0


_PDL = os.environ.get('TXL_DECODE_PDL', '1') != '0'       # programmatic dependent launch of the step's kernels
_ATTN_SPLITS = int(os.environ.get('TXL_DECODE_ATTN_SPLITS', '0'))   # 0 = automatic
# per layer: MB of the next attention kernel's ring that the six kernels before it ask into L2 (cp.async.bulk.prefetch.L2).  Measured at cfg4:
# 0 MB 618.8 us/step, 48 MB 623.1, 80 MB 641.8, 112 MB 692.3 - the small kernels slow down by more than the attention kernel gains: off.
_PREFETCH_MB = float(os.environ.get('TXL_DECODE_PREFETCH_MB', '0'))
_SPLIT_COLS = int(os.environ.get('TXL_DEC_SPLIT_COLS', '1000000'))   # A/B switch: 512 makes every Linear fit two ring stages (see TXL_DEC_LINEAR_2STAGE)
_ABL = int(os.environ.get('TXL_DECODE_ABL', '0'))       # timing ablations (results are garbage): 1 = no attention launch, 2 = no Linear / LayerNorm launches


_PERSIST = os.environ.get('TXL_DECODE_PERSIST', '0') == '1'       # grid-barrier variant: correct, measured slower than the launch chain (profiles/r02_decode_ab.txt)


def persist_supported(model, B):
    """The persistent one-kernel step (csrc/decode_persist.cu) takes this model and `B` sequences per chain."""
    cfg = model.config
    if not _PERSIST or model._E.dtype != torch.bfloat16 or not cfg.same_length or cfg.mem_len <= 0:
        return False
    Vx = cfg.vocab_size + len(getattr(cfg, 'cutoffs', []) or [])
    return bool(load().txl_decode_persist_supported(int(B), cfg.n_head, cfg.d_head, cfg.d_model, cfg.d_inner, cfg.mem_len, cfg.n_layer, Vx))


_CLUSTER = os.environ.get('TXL_DECODE_CLUSTER', '1') != '0'
CL_MAXB = 64       # sequences per chain of the cluster engine: 8 per cluster x the co-resident clusters of 8 SMs (>= 8 on a B200)
# The cluster engine is the default up to this many sequences per GPU (measured, profiles/r02_decode_ab.txt: 378 vs 416 us/step at 8 sequences,
# 412 vs 414 at 16; at 32 / 64 the launch chain with sequence groups still wins: 515 vs 422, 624 vs 543).  TXL_DECODE_CLUSTER_MAXB overrides.
CL_AUTO_MAXB = int(os.environ.get('TXL_DECODE_CLUSTER_MAXB', '16'))


def cluster_supported(model, B):
    """The cluster engine (csrc/decode_cluster.cu: 8-CTA clusters, one per attention head, DSMEM hand-over) takes this model and `B` sequences."""
    cfg = model.config
    if not _CLUSTER or B > CL_AUTO_MAXB or model._E.dtype != torch.bfloat16 or not cfg.same_length or cfg.mem_len <= 0:
        return False
    Vx = cfg.vocab_size + len(getattr(cfg, 'cutoffs', []) or [])
    return bool(load().txl_decode_cluster_supported(int(B), cfg.n_head, cfg.d_head, cfg.d_model, cfg.d_inner, cfg.mem_len, cfg.n_layer, Vx))


def _dec_linear_ok(A, W):
    return (A.dtype == torch.bfloat16 and A.shape[1] % 32 == 0 and A.stride(1) == 1 and W.stride(1) == 1 and A.stride(0) % 8 == 0
            and W.stride(0) % 8 == 0)


def _skinny(A, W, bias=None, relu=False, out=None, pf=(None, 0)):
    """y[B, N] = x[B, K] W[N, K]^T (+bias)(ReLU).  bf16: `txl_dec_linear` (8-16 output features per CTA, mma.sync fragments out of a cp.async
    ring, so 64-150 CTAs stream the weights); fp32 parity mode: the exact-FMA weight-streaming kernel."""
    M, K = A.shape
    N = W.shape[0]
    if _dec_linear_ok(A, W):
        if out is None:
            out = torch.empty(M, N, dtype=A.dtype, device=A.device)
        assert out.dtype in (torch.bfloat16, torch.float32) and out.stride(1) == 1
        check(load().txl_dec_linear(ptr(A), A.stride(0), ptr(W), W.stride(0), ptr(bias), ptr(out), out.stride(0), M, N, K, int(relu),
                                    int(out.dtype == torch.float32), 1, pf[0], pf[1], stream_ptr()), 'dec_linear')
        return out
    if A.dtype == torch.bfloat16 and K >= 16:
        return ops.gemm(W, A, transB=True, bias=bias, relu=relu, out=out, bias_row=bias is not None, transpose_out=True)
    if out is None:
        out = torch.empty(M, N, dtype=A.dtype, device=A.device)
    check(load().txl_skinny_gemm(ptr(A), A.stride(0), ptr(W), W.stride(0), ptr(bias), ptr(out), out.stride(0), M, N, K, int(relu), dtype_code(A.dtype),
                                 stream_ptr()), 'skinny_gemm')
    return out


_SMS = {}


def _sm_count(dev):
    if dev not in _SMS:
        _SMS[dev] = torch.cuda.get_device_properties(dev).multi_processor_count
    return _SMS[dev]


def _linear_add_ln(x, A, W, bias, gamma, beta, eps, pf0=(None, 0), pf1=(None, 0)):
    """LayerNorm(x + A W^T + bias): split-K `txl_dec_linear` leaves fp32 partial planes, `txl_dec_add_ln` sums them with the bias and the
    residual and normalises (one launch for what was GEMM epilogue + add + LayerNorm)."""
    M, K = A.shape
    N = W.shape[0]
    lib = load()
    sms = _sm_count(A.device)
    splits = max(1, min(K // 256, sms // ((N + 7) // 8)), K // _SPLIT_COLS)       # also: at most _SPLIT_COLS columns per CTA
    while splits > 1 and K % (32 * splits):
        splits -= 1
    part = torch.empty(splits, M, N, dtype=torch.float32, device=A.device)
    check(lib.txl_dec_linear(ptr(A), A.stride(0), ptr(W), W.stride(0), None, ptr(part), N, M, N, K, 0, 1, splits, pf0[0], pf0[1], stream_ptr()), 'dec_linear')
    y = torch.empty(M, N, dtype=torch.bfloat16, device=A.device)
    check(lib.txl_dec_add_ln(ptr(x), ptr(part), splits, ptr(bias), ptr(gamma), ptr(beta), ptr(y), M, N, float(eps), pf1[0], pf1[1], stream_ptr()), 'dec_add_ln')
    return y


def sequence_groups(model, B, requested=None):
    """How many independent sequence groups generate() decodes as parallel graph branches (GroupedDecoder).  bf16 second-generation path only."""
    if requested is None:
        requested = int(os.environ.get('TXL_DECODE_GROUPS', '0')) or None
    if hasattr(model, '_W') and cluster_supported(model, B):
        # clusters own their sequences end to end: one chain, no sequence groups
        return max(1, min(int(requested), B)) if requested is not None else 1
    need = (B + SK_MAXM - 1) // SK_MAXM              # a Decoder (one chain of kernels) takes at most SK_MAXM sequences
    if hasattr(model, '_W') and persist_supported(model, (B + need - 1) // need):
        # the persistent step occupies every SM: sequence groups would only serialise, so as few chains as the 64-row limit allows
        return max(need, min(int(requested), B)) if requested is not None else need
    if model._E.dtype != torch.bfloat16 or os.environ.get('TXL_DECODE_TAIL', '1') == '0' or model.config.vocab_size > 8192 or getattr(model.config, 'cutoffs', None):
        return need
    if requested is not None:
        return max(need, min(int(requested), B))
    if B > SK_MAXM:
        return (B + 15) // 16
    # groups of 16 sequences: the Linears then work on ONE 16-row MMA tile, and 16 x 8 heads = 128 attention CTAs sit one per SM.  Measured at 64
    # sequences (us/step): 1 group 613, 2 groups 601, 3 groups 710, 4 groups 546, 5 groups 567, 8 groups 701.
    return min(4, B // 16) if B >= 32 else 1


class Decoder:
    """Device-resident generation state for `B` sequences.  Built from the mems the prompt forward returned."""

    def __init__(self, model, mems, out_ids, col0, *, do_sample, temperature, top_k, top_p, eos_token_id, pad_token_id, seed=0, seq_offset=0,
                 use_graph=True, attn_splits=None, persist=True):
        cfg = model.config
        self.model, self.cfg = model, cfg
        bm = mems._bm if hasattr(mems, '_bm') else model._mems_to_bm(mems, out_ids.shape[0])
        B, ML, d = bm[0].shape
        if ML != cfg.mem_len:
            raise ValueError('decode cache needs mems of exactly mem_len rows')
        self.B, self.ML, self.d = B, ML, d
        H, dh = cfg.n_head, cfg.d_head
        dt, dev = model._E.dtype, model._E.device
        self.dt, self.dev = dt, dev
        lib = load()
        # per-layer ring caches of projected keys / values + cached r tables (weights are frozen during generation)
        pos_tab = ops.posemb_table(ML + 1, cfg.clamp_len, d, dt, dev)
        self.kc, self.vc, self.kvc, self.r, self.r_hm = [], [], [], [], []
        self.pipe_attn = dt == torch.bfloat16
        self.cluster = persist and cluster_supported(model, B)
        self.persist = self.cluster or (persist and persist_supported(model, B))       # both run over the hidden-state ring
        if self.persist:
            # the cache IS the hidden-state mems (chronological rows: slot 0 = oldest = the first one to be overwritten); private copies,
            # the step writes into them.  K projection transposed per head once: wkT[h, c, e] = W_k[h*64 + e, c].
            self.ring = [t.contiguous().clone() for t in bm]
            self.wkT = [w.qkv[d:2 * d].view(H, dh, d).transpose(1, 2).contiguous() for w in model._W]
            self.r = [ops.gemm(pos_tab, w.r, transB=True) for w in model._W]                          # (ML+1, d): row x <-> distance ML - x
        for li, w in enumerate(model._W if not self.persist else []):
            kv = ops.gemm(bm[li].reshape(B * ML, d).contiguous(), w.qkv[d:], transB=True)          # (B*ML, 2d)
            if self.pipe_attn:
                kvc = torch.empty(B, H, ML, 2 * dh, dtype=dt, device=dev)      # a key's k row then its v row: one contiguous run per stage of keys
                check(lib.txl_decode_cache_init_kv(ptr(kv), kv.stride(0), ptr(kvc), B, H, ML, dh, stream_ptr()), 'decode_cache_init_kv')
                self.kvc.append(kvc)
            else:
                kc = torch.empty(B, H, ML, dh, dtype=dt, device=dev)
                vc = torch.empty(B, H, ML, dh, dtype=dt, device=dev)
                check(lib.txl_decode_cache_init(ptr(kv), kv.stride(0), ptr(kc), ptr(vc), B, H, ML, dh, dtype_code(dt), stream_ptr()), 'decode_cache_init')
                self.kc.append(kc)
                self.vc.append(vc)
            self.r.append(ops.gemm(pos_tab, w.r, transB=True))                                       # (ML+1, d)
            if self.pipe_attn:
                rh = torch.empty(H, ML + 1, dh, dtype=dt, device=dev)                                # per-head rows contiguous for the bulk copies
                check(lib.txl_decode_rtab_head_major(ptr(self.r[-1]), ptr(rh), ML + 1, H, dh, stream_ptr()), 'decode_rtab_head_major')
                self.r_hm.append(rh)
        # ring splits of the attention kernel (several CTAs per (sequence, head), merged by the last to arrive).  Measured at cfg4, 64 sequences:
        # 1 split 667 us/step, 2 splits 703, 4 splits 720 — the merge costs more than the better SM balance returns; with few sequences per GPU
        # (8 x 8 heads = 64 CTAs for 592 slots) a CTA streaming its whole ring alone is latency-bound, so the ring is cut to fill the chip.
        auto = max(1, min(8, 512 // max(1, B * H), (ML * dh) // (4 * 2048)))
        self.attn_splits = _ATTN_SPLITS if _ATTN_SPLITS > 0 else (attn_splits or auto)
        # stage geometry of the attention kernel: with ring splits a CTA only sees mem_len / splits keys, so it needs short stages to pipeline at
        # all (8 sequences, 8 splits: 32 keys x 4 stages 421 us/step, 128 keys x 2 stages 689); unsplit rings take the library default (128 x 2)
        self.attn_cfg = 0 if (self.attn_splits > 1 and 'TXL_DECODE_ATTN_CFG' not in os.environ) else None
        self.attn_ws = self.attn_cnt = None
        self.pf_bytes = int(min(B * H * ML * 2 * dh * 2, _PREFETCH_MB * 2 ** 20)) if self.pipe_attn else 0
        self._abl_qkv = torch.zeros(B, 3 * d, dtype=dt, device=dev) if _ABL else None
        self.gen2_ln = self.pipe_attn and d % 32 == 0 and cfg.d_inner % 32 == 0 and d <= 1024      # split-K Linear + fused add/LayerNorm
        if self.pipe_attn and self.attn_splits > 1:
            self.attn_ws = torch.empty(lib.txl_decode_attn_pipe_ws_bytes(B, H, dh, self.attn_splits), dtype=torch.uint8, device=dev)
            self.attn_cnt = torch.zeros(B * H, dtype=torch.int32, device=dev)
        # fused step tail (log-softmax + keyed uniform + sampler + eos/pad bookkeeping + next embedding + step counter in one kernel)
        self.gen2_tail = self.pipe_attn and os.environ.get('TXL_DECODE_TAIL', '1') != '0' and cfg.vocab_size <= 8192 and not cfg.cutoffs
        self.x0 = torch.empty(B, d, dtype=dt, device=dev) if self.gen2_tail else None
        self.tail_arrive = torch.zeros(1, dtype=torch.int32, device=dev) if self.gen2_tail else None
        self.pos = torch.zeros(1, dtype=torch.int32, device=dev)
        self.tok = torch.empty(B, dtype=torch.int64, device=dev)
        self.next = torch.empty(B, dtype=torch.int64, device=dev)
        self.u = torch.zeros(B, dtype=torch.float32, device=dev)
        self.unfinished = torch.ones(B, dtype=torch.int64, device=dev)
        self.out_ids, self.col0 = out_ids, col0
        self.do_sample, self.temperature, self.top_k, self.top_p = bool(do_sample), float(temperature), int(top_k or 0), float(top_p)
        self.eos, self.pad = eos_token_id, pad_token_id
        self.seed, self.seq_offset = int(seed), int(seq_offset)
        self.V = cfg.vocab_size
        self.Vx = self.V + len(cfg.cutoffs)                # LM-head columns: token logits + one per adaptive-softmax cluster
        self.Vp = (self.Vx + 7) // 8 * 8
        self.logits = torch.zeros(B, self.Vp, dtype=torch.float32 if dt == torch.bfloat16 else dt, device=dev)
        if self.persist:
            self._build_persist()
        self.scores = torch.zeros(B, cfg.vocab_size, dtype=torch.float32, device=dev) if self.gen2_tail else None      # log-probs of the last step
        self.graph = None
        self.use_graph = use_graph
        self.steps_done = 0

    def _build_persist(self):
        """Per-layer pointer tables + workspace of the persistent step kernel; uploads the table (one synchronising call)."""
        m, cfg, lib = self.model, self.cfg, load()
        L = cfg.n_layer

        def arr(ts):
            return (C.c_void_p * L)(*[t.data_ptr() for t in ts])
        W = m._W
        self._persist_keep = [[w.qkv for w in W], self.wkT, [w.o for w in W], [w.w1 for w in W], [w.w2 for w in W], self.r, [w.b1 for w in W],
                              [w.b2 for w in W], [w.rwb for w in W], [w.rrb for w in W], [w.ln1_w for w in W], [w.ln1_b for w in W],
                              [w.ln2_w for w in W], [w.ln2_b for w in W], self.ring]
        self._persist_arrays = [arr(ts) for ts in self._persist_keep]
        ws_bytes = lib.txl_decode_cluster_ws_bytes if self.cluster else lib.txl_decode_persist_ws_bytes
        nbytes = ws_bytes(self.B, cfg.n_head, cfg.d_head, self.d, cfg.d_inner, self.ML, L, self.Vx)
        self._persist_ws = torch.zeros(nbytes + 256, dtype=torch.uint8, device=self.dev)
        off = (-self._persist_ws.data_ptr()) % 256
        self._persist_ws_ptr = self._persist_ws.data_ptr() + off
        if self.x0 is None:
            self.x0 = torch.empty(self.B, self.d, dtype=self.dt, device=self.dev)
        self._persist_call(1)

    def _persist_call(self, build):
        cfg = self.cfg
        step = load().txl_decode_cluster_step if self.cluster else load().txl_decode_persist_step
        check(step(*self._persist_arrays, ptr(self.model._E_ext), ptr(self.model._out_bias_ext), ptr(self.x0), ptr(self.pos),
                   ptr(self.logits), self.logits.stride(0), self._persist_ws_ptr, int(build), self.B, cfg.n_head, cfg.d_head,
                   self.d, cfg.d_inner, self.ML, cfg.n_layer, self.Vx, float(cfg.layer_norm_epsilon), stream_ptr()),
              'decode_cluster_step' if self.cluster else 'decode_persist_step')

    # one decode step: every line is a kernel launch on the current stream
    def _step_kernels(self):
        m, cfg, lib = self.model, self.cfg, load()
        B, d, H, dh, ML = self.B, self.d, cfg.n_head, cfg.d_head, self.ML
        if self.persist:
            if not self.gen2_tail:                     # (adaptive-softmax / large vocabularies: the embedding is a separate launch)
                self.x0.copy_(ops.embed_fwd(self.tok, m._E, math.sqrt(d)))
            self._persist_call(0)
            return self._finish_step(self.logits)
        pdl_old = lib.txl_set_pdl(1) if (self.pipe_attn and _PDL) else None
        cfg_old = lib.txl_decode_attn_pipe_config(self.attn_cfg) if (self.pipe_attn and self.attn_cfg is not None) else None
        try:
            # second generation: x0 already holds the embedding of the current token (run() for the first step, the tail kernel afterwards)
            x = self.x0 if self.gen2_tail else ops.embed_fwd(self.tok, m._E, math.sqrt(d))
            nl = len(m._W)

            def pf(li_next, k):
                """k-th sixth of the L2-prefetch window of layer `li_next`'s ring (the six kernels between two attention kernels share it)."""
                if not self.pipe_attn or self.pf_bytes <= 0:
                    return (None, 0)
                sz = self.pf_bytes // 6 // 128 * 128
                return (self.kvc[li_next % nl].data_ptr() + k * sz, sz)

            for li, w in enumerate(m._W):
                qkv = (_skinny(x, w.qkv, pf=pf(li, 5)) if li > 0 else _skinny(x, w.qkv)) if not (_ABL & 2) else self._abl_qkv
                vec = torch.empty(B, d, dtype=self.dt, device=self.dev)
                if _ABL & 1:
                    pass
                elif self.pipe_attn:
                    check(lib.txl_decode_attn_pipe(ptr(qkv), ptr(self.kvc[li]), ptr(self.r_hm[li]), ptr(w.rwb), ptr(w.rrb), ptr(vec),
                                                   ptr(self.pos), B, H, ML, dh, self.attn_splits, ptr(self.attn_ws), ptr(self.attn_cnt), stream_ptr()),
                          'decode_attn_pipe')
                else:
                    check(lib.txl_decode_attn(ptr(qkv), ptr(self.kc[li]), ptr(self.vc[li]), ptr(self.r[li]), ptr(w.rwb), ptr(w.rrb), ptr(vec), ptr(self.pos),
                                              B, H, ML, dh, dtype_code(self.dt), stream_ptr()), 'decode_attn')
                if _ABL & 2:
                    continue
                if self.gen2_ln:
                    y1 = _linear_add_ln(x, vec, w.o, None, w.ln1_w, w.ln1_b, cfg.layer_norm_epsilon, pf(li + 1, 0), pf(li + 1, 1))
                    hdn = _skinny(y1, w.w1, bias=w.b1, relu=True, pf=pf(li + 1, 2))
                    x = _linear_add_ln(y1, hdn, w.w2, w.b2, w.ln2_w, w.ln2_b, cfg.layer_norm_epsilon, pf(li + 1, 3), pf(li + 1, 4))
                    continue
                ao = _skinny(vec, w.o)
                y1, _, _, _ = ops.add_ln_fwd(x, ao, w.ln1_w, w.ln1_b, cfg.layer_norm_epsilon, save=False)
                hdn = _skinny(y1, w.w1, bias=w.b1, relu=True)
                f = _skinny(hdn, w.w2, bias=w.b2)
                x, _, _, _ = ops.add_ln_fwd(y1, f, w.ln2_w, w.ln2_b, cfg.layer_norm_epsilon, save=False)
            _skinny(x, m._E_ext, bias=m._out_bias_ext, out=self.logits[:, :self.Vx], pf=pf(0, 5))      # the next step's first attention kernel
        finally:
            if pdl_old is not None:
                lib.txl_set_pdl(pdl_old)
            if cfg_old is not None:
                lib.txl_decode_attn_pipe_config(cfg_old)
        return self._finish_step(self.logits)

    def _finish_step(self, logits):
        lib, B = load(), self.B
        if self.gen2_tail and logits is self.logits:
            use_eos = self.eos is not None
            pdl_old = lib.txl_set_pdl(1) if _PDL else None
            try:
                check(lib.txl_decode_tail(ptr(logits), logits.stride(0), ptr(self.scores), B, self.V, int(self.do_sample), self.temperature, self.top_k,
                                          self.top_p, self.seed, self.seq_offset, ptr(self.tok), ptr(self.unfinished), ptr(self.out_ids),
                                          self.out_ids.stride(0), self.col0, ptr(self.pos), ptr(self.tail_arrive), int(self.eos if use_eos else 0),
                                          int(self.pad if self.pad is not None else 0), int(use_eos), ptr(self.model._E), ptr(self.x0), self.d,
                                          math.sqrt(self.d), stream_ptr()), 'decode_tail')
            finally:
                if pdl_old is not None:
                    lib.txl_set_pdl(pdl_old)
            return
        if self.cfg.cutoffs:
            _, _, logprobs, _ = ops.adaptive_lsm_nll_fwd(logits, self.V, self.cfg.cutoffs, None, want_logprobs=True)
        else:
            _, _, logprobs, _ = ops.logsoftmax_nll_fwd(logits, self.V, None, want_logprobs=True)
        self.scores = logprobs
        if self.do_sample:
            check(lib.txl_decode_uniform(ptr(self.u), B, self.seed, self.seq_offset, ptr(self.pos), stream_ptr()), 'decode_uniform')
        check(lib.txl_sample(ptr(logprobs), B, self.V, int(self.do_sample), self.temperature, self.top_k, self.top_p, ptr(self.u), ptr(self.next), None, None,
                             stream_ptr()), 'sample')
        use_eos = self.eos is not None
        check(lib.txl_decode_commit(ptr(self.next), ptr(self.tok), ptr(self.unfinished), ptr(self.out_ids), self.out_ids.stride(0), self.col0, ptr(self.pos), B,
                                    int(self.eos if use_eos else 0), int(self.pad if self.pad is not None else 0), int(use_eos), stream_ptr()), 'decode_commit')

    def set_unfinished(self, unfinished):
        self.unfinished.copy_(unfinished)

    def last_scores(self):
        """(B, V) fp32 log-probs the last step sampled from (what HF's `scores` holds before the warpers)."""
        return self.scores

    def run(self, first_token, n_steps, poll_every=64):
        """Feed `first_token` (B,) and generate `n_steps` tokens into out_ids[:, col0:col0+n_steps].  Returns steps actually run."""
        self.tok.copy_(first_token)
        done = 0
        if n_steps <= 0:
            return 0
        if self.gen2_tail:
            self.x0.copy_(ops.embed_fwd(self.tok, self.model._E, math.sqrt(self.d)))
        self._step_kernels()                      # eager first step: warms caches, sets kernel attributes
        done += 1
        if self.use_graph and n_steps > 1:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                # capture one step; nothing here depends on host-side values that change between steps.  capture_begin / capture_end directly:
                # the torch.cuda.graph context manager also runs gc.collect() and torch.cuda.empty_cache(), a fixed cost per generate() call
                g.capture_begin()
                try:
                    self._step_kernels()
                finally:
                    g.capture_end()
            torch.cuda.current_stream().wait_stream(side)
            self.graph = g
            # the capture itself does not execute: pos / caches are untouched
        while done < n_steps:
            if self.graph is not None:
                self.graph.replay()
            else:
                self._step_kernels()
            done += 1
            if self.eos is not None and done % poll_every == 0 and int(self.unfinished.max().item()) == 0:
                break
        self.steps_done = done
        return done


def make_decoder(model, mems, out_ids, col0, groups=None, **kw):
    """The decode engine generate() uses for `out_ids.shape[0]` sequences: one chain of kernels (Decoder) or several sequence groups as
    parallel graph branches (GroupedDecoder).  Both expose run(first_token, n_steps), set_unfinished(mask) and last_scores()."""
    B = out_ids.shape[0]
    g = sequence_groups(model, B, groups)
    if g > 1:
        return GroupedDecoder(model, mems, out_ids, col0, g, **kw)
    return Decoder(model, mems, out_ids, col0, **kw)


class _BmSlice:
    """Batch slice of the batch-major mems (what Decoder reads through `._bm`)."""

    def __init__(self, bm, lo, hi):
        self._bm = [t[lo:hi] for t in bm]


class GroupedDecoder:
    """The sequences of one GPU cut into `groups` independent Decoders whose steps are captured as PARALLEL branches of one CUDA graph.
    A step is a chain of one HBM-bound kernel (attention over the ring) and six latency-bound ones per layer; sequences never interact
    (batched sampling shards by sequence with no collective), so while one group's attention kernel streams its ring the other
    groups run their Linears: HBM stays busy and the per-layer latency chain of a group is hidden behind the other groups' traffic.
    Draws are keyed on the global sequence index, so the tokens do not depend on the grouping."""

    def __init__(self, model, mems, out_ids, col0, groups, *, seq_offset=0, **kw):
        B = out_ids.shape[0]
        bm = mems._bm if hasattr(mems, '_bm') else model._mems_to_bm(mems, B)
        per, rem = divmod(B, groups)
        self.bounds, lo = [], 0
        for g in range(groups):
            hi = lo + per + (1 if g < rem else 0)
            self.bounds.append((lo, hi))
            lo = hi
        # the other groups keep the chip busy while one group's attention kernel runs: no ring splits (4 splits: 727.6 us/step, none: 546.0)
        kw.setdefault('attn_splits', 1)
        kw.setdefault('persist', False)       # sequence groups are the launch chain's way to fill the chip: its sub-decoders never pick a one-kernel engine
        self.decs = [Decoder(model, _BmSlice(bm, lo, hi), out_ids[lo:hi], col0, seq_offset=seq_offset + lo, **kw) for lo, hi in self.bounds]
        self.eos = self.decs[0].eos
        self.use_graph = self.decs[0].use_graph
        self.graph = None
        self.steps_done = 0

    def set_unfinished(self, unfinished):
        for d, (lo, hi) in zip(self.decs, self.bounds):
            d.unfinished.copy_(unfinished[lo:hi])

    def last_scores(self):
        return torch.cat([d.scores for d in self.decs], 0)

    def _all_unfinished_max(self):
        return max(int(d.unfinished.max().item()) for d in self.decs)

    def run(self, first_token, n_steps, poll_every=64):
        if n_steps <= 0:
            return 0
        for d, (lo, hi) in zip(self.decs, self.bounds):
            d.tok.copy_(first_token[lo:hi])
            if d.gen2_tail:
                d.x0.copy_(ops.embed_fwd(d.tok, d.model._E, math.sqrt(d.d)))
            d._step_kernels()                         # eager first step: warms caches, sets kernel attributes
        done = 1
        if self.use_graph and n_steps > 1:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            main = torch.cuda.Stream()
            branches = [main] + [torch.cuda.Stream() for _ in self.decs[1:]]
            main.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(main):
                g.capture_begin()                     # not torch.cuda.graph(): that also runs gc.collect() + empty_cache() on every call
                try:
                    fork = torch.cuda.Event()
                    fork.record(main)
                    for st, d in zip(branches, self.decs):
                        if st is not main:
                            st.wait_event(fork)
                        with torch.cuda.stream(st):
                            d._step_kernels()
                    for st in branches[1:]:
                        join = torch.cuda.Event()
                        join.record(st)
                        main.wait_event(join)
                finally:
                    g.capture_end()
            torch.cuda.current_stream().wait_stream(main)
            self.graph = g
        while done < n_steps:
            if self.graph is not None:
                self.graph.replay()
            else:
                for d in self.decs:
                    d._step_kernels()
            done += 1
            if self.eos is not None and done % poll_every == 0 and self._all_unfinished_max() == 0:
                break
        self.steps_done = done
        return done
