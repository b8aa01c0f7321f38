"""Mems-cached decode step for `generate`: one new token per sequence per step, projected-K/V ring cache, every per-step op a C-ABI
kernel, the whole step replayed as a CUDA graph (no host round trip between tokens).

Follows HF 4.25 `sample` / `greedy_search` step for step: embed last token -> L x (qkv, relative-position attention
over [mems ; current] with the same_length band, o_net + LN, FF + LN) -> log-softmax -> warpers -> draw -> eos/pad bookkeeping -> append.
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import ops
from ._lib import check, dtype_code, load, ptr, stream_ptr

SK_MAXM = 64


def supported(model, B):
    cfg = model.config
    return B <= SK_MAXM and cfg.same_length and cfg.mem_len > 0 and cfg.d_head in (32, 64, 128) and cfg.d_model % 8 == 0 and cfg.d_inner % 8 == 0


def _skinny(A, W, bias=None, relu=False, out=None):
    """y[B, N] = x[B, K] W[N, K]^T (+bias)(ReLU).  bf16: the tcgen05 GEMM computes y^T = W x^T (output features fill the 128-row MMA tile,
    the <=64 sequences are the MMA N) and stores it transposed; fp32 parity mode: the exact-FMA weight-streaming kernel."""
    if A.dtype == torch.bfloat16 and A.shape[1] >= 16:
        return ops.gemm(W, A, transB=True, bias=bias, relu=relu, out=out, bias_row=bias is not None, transpose_out=True)
    M, K = A.shape
    N = W.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=A.dtype, device=A.device)
    check(load().txl_skinny_gemm(ptr(A), A.stride(0), ptr(W), W.stride(0), ptr(bias), ptr(out), out.stride(0), M, N, K, int(relu), dtype_code(A.dtype),
                                 stream_ptr()), 'skinny_gemm')
    return out


class Decoder:
    """Device-resident generation state for `B` sequences.  Built from the mems the prompt forward returned."""

    def __init__(self, model, mems, out_ids, col0, *, do_sample, temperature, top_k, top_p, eos_token_id, pad_token_id, seed=0, seq_offset=0,
                 use_graph=True, use_fused=False):
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
        self.kc, self.vc, self.r = [], [], []
        for li, w in enumerate(model._W):
            kv = ops.gemm(bm[li].reshape(B * ML, d).contiguous(), w.qkv[d:], transB=True)          # (B*ML, 2d)
            kc = torch.empty(B, H, ML, dh, dtype=dt, device=dev)
            vc = torch.empty(B, H, ML, dh, dtype=dt, device=dev)
            check(lib.txl_decode_cache_init(ptr(kv), kv.stride(0), ptr(kc), ptr(vc), B, H, ML, dh, dtype_code(dt), stream_ptr()), 'decode_cache_init')
            self.kc.append(kc)
            self.vc.append(vc)
            self.r.append(ops.gemm(pos_tab, w.r, transB=True))                                       # (ML+1, d)
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
        self.Vp = (self.V + 7) // 8 * 8
        self.logits = torch.zeros(B, self.Vp, dtype=torch.float32 if dt == torch.bfloat16 else dt, device=dev)
        self.scores = None
        self.graph = None
        self.use_graph = use_graph
        self.steps_done = 0
        self.fused = None
        if use_fused:
            self._build_fused()

    def _build_fused(self):
        """Pointer tables + workspace of the persistent fused step kernel (csrc/decode_fused.cu)."""
        m, cfg, lib = self.model, self.cfg, load()
        L = cfg.n_layer

        def arr(ts):
            return (C.c_void_p * L)(*[t.data_ptr() for t in ts])
        W = m._W
        self._fused_keep = [[w.qkv for w in W], [w.o for w in W], [w.w1 for w in W], [w.w2 for w in W], self.r, [w.b1 for w in W], [w.b2 for w in W],
                            [w.rwb for w in W], [w.rrb for w in W], [w.ln1_w for w in W], [w.ln1_b for w in W], [w.ln2_w for w in W], [w.ln2_b for w in W],
                            self.kc, self.vc]
        self._fused_arrays = [arr(ts) for ts in self._fused_keep]
        nbytes = lib.txl_decode_fused_workspace(self.B, self.d, cfg.d_inner, self.V, L, dtype_code(self.dt))
        self._fused_ws = torch.zeros(nbytes, dtype=torch.uint8, device=self.dev)
        self.logits32 = torch.zeros(self.B, self.Vp, dtype=torch.float32, device=self.dev)
        self._fused_call(1)
        self.fused = True

    def _fused_call(self, build):
        cfg = self.cfg
        check(load().txl_decode_fused_step(*self._fused_arrays, ptr(self.model._E), ptr(self.model._out_bias), ptr(self.tok), ptr(self.pos),
                                           ptr(self.logits32), ptr(self._fused_ws), int(build), self.B, cfg.n_head, cfg.d_head, self.d, cfg.d_inner,
                                           self.ML, cfg.n_layer, self.V, self.Vp, float(cfg.layer_norm_epsilon), dtype_code(self.dt), stream_ptr()),
              'decode_fused_step')

    # one decode step: every line is a kernel launch on the current stream
    def _step_kernels(self):
        m, cfg, lib = self.model, self.cfg, load()
        B, d, H, dh, ML = self.B, self.d, cfg.n_head, cfg.d_head, self.ML
        if self.fused:
            self._fused_call(0)
            return self._finish_step(self.logits32)
        x = ops.embed_fwd(self.tok, m._E, math.sqrt(d))
        for li, w in enumerate(m._W):
            qkv = _skinny(x, w.qkv)
            vec = torch.empty(B, d, dtype=self.dt, device=self.dev)
            check(lib.txl_decode_attn(ptr(qkv), ptr(self.kc[li]), ptr(self.vc[li]), ptr(self.r[li]), ptr(w.rwb), ptr(w.rrb), ptr(vec), ptr(self.pos),
                                      B, H, ML, dh, dtype_code(self.dt), stream_ptr()), 'decode_attn')
            ao = _skinny(vec, w.o)
            y1, _, _, _ = ops.add_ln_fwd(x, ao, w.ln1_w, w.ln1_b, cfg.layer_norm_epsilon, save=False)
            hdn = _skinny(y1, w.w1, bias=w.b1, relu=True)
            f = _skinny(hdn, w.w2, bias=w.b2)
            x, _, _, _ = ops.add_ln_fwd(y1, f, w.ln2_w, w.ln2_b, cfg.layer_norm_epsilon, save=False)
        _skinny(x, m._E, bias=m._out_bias, out=self.logits[:, :self.V])
        return self._finish_step(self.logits)

    def _finish_step(self, logits):
        lib, B = load(), self.B
        _, _, logprobs, _ = ops.logsoftmax_nll_fwd(logits, self.V, None, want_logprobs=True)
        self.scores = logprobs
        if self.do_sample:
            check(lib.txl_decode_uniform(ptr(self.u), B, self.seed, self.seq_offset, ptr(self.pos), stream_ptr()), 'decode_uniform')
        check(lib.txl_sample(ptr(logprobs), B, self.V, int(self.do_sample), self.temperature, self.top_k, self.top_p, ptr(self.u), ptr(self.next), None, None,
                             stream_ptr()), 'sample')
        use_eos = self.eos is not None
        check(lib.txl_decode_commit(ptr(self.next), ptr(self.tok), ptr(self.unfinished), ptr(self.out_ids), self.out_ids.stride(0), self.col0, ptr(self.pos), B,
                                    int(self.eos if use_eos else 0), int(self.pad if self.pad is not None else 0), int(use_eos), stream_ptr()), 'decode_commit')

    def run(self, first_token, n_steps, poll_every=64):
        """Feed `first_token` (B,) and generate `n_steps` tokens into out_ids[:, col0:col0+n_steps].  Returns steps actually run."""
        self.tok.copy_(first_token)
        done = 0
        if n_steps <= 0:
            return 0
        self._step_kernels()                      # eager first step: warms caches, sets kernel attributes
        done += 1
        if self.use_graph and n_steps > 1:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                # capture one step; nothing here depends on host-side values that change between steps
                with torch.cuda.graph(g, stream=side):
                    self._step_kernels()
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
