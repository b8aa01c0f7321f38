"""Drop-in boundary: `MyTransfoXLConfig`, `MyTransfoXLLMHeadModel`, `TransfoXLLMHeadModelOutput`.

Mirrors reference `musicnlp/models/transformer_xl.py` (config :15-77, output dataclass :81-124, forward :130-221,
prepare_inputs_for_generation :223-241) — same names, keyword arguments, outputs, side effects and errors — with the
HF 4.25.1 modules underneath replaced by the sm_100a kernels of `libtxl_b200.so`.
"""
from __future__ import annotations

import json
import os
from collections import OrderedDict
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn

from . import engine, ops
from ._lib import TxlError

PT_LOSS_PAD = -100   # musicnlp/util/train/train_util_wrap.py:22

__all__ = ['MyTransfoXLConfig', 'MyTransfoXLLMHeadModel', 'TransfoXLLMHeadModelOutput', 'TxlMems']


# ============================================================================= config
class MyTransfoXLConfig:
    """reference transformer_xl.py:15-77; every field not set there keeps the HF `TransfoXLConfig` default
    (confirmed by the logged dump, notebook/train/transformer-xl.ipynb:492-513)."""
    model_type = 'transfo-xl'
    presets = {
        'debug': dict(d_model=128, n_head=8, n_layer=4),
        'debug-large': dict(d_model=128, n_head=8, n_layer=4),
        'tiny': dict(d_model=256, n_head=8, n_layer=6),
        'small': dict(d_model=512, n_head=8, n_layer=12),
        'base': dict(d_model=768, n_head=12, n_layer=12),
        'large': dict(d_model=1024, n_head=16, n_layer=18),
    }
    size2max_length = {'debug': 64, 'debug-large': 128, 'tiny': 512, 'small': 1024, 'base': 2048, 'large': 2048}
    for _k, _c in presets.items():
        _d, _h = _c['d_model'], _c['n_head']
        assert _d % _h == 0
        if 'debug' in _k:
            _m, _cl = 64, 64
        else:
            _m, _cl = max(128, size2max_length[_k] // 8), max(1024, size2max_length[_k] // 2)
        _c.update(dict(d_embed=_d, d_inner=_d * 4, d_head=_d // _h, mem_len=_m, clamp_len=_cl, div_val=1))
    del _k, _c, _d, _h, _m, _cl

    _hf_defaults = dict(
        vocab_size=267735, cutoffs=[20000, 40000, 200000], d_model=1024, d_embed=1024, n_head=16, d_head=64, d_inner=4096,
        div_val=4, pre_lnorm=False, n_layer=18, mem_len=1600, clamp_len=1000, same_length=True, proj_share_all_but_first=True,
        attn_type=0, sample_softmax=-1, adaptive=True, dropout=0.1, dropatt=0.0, untie_r=True, init='normal', init_range=0.01,
        proj_init_std=0.01, init_std=0.02, layer_norm_epsilon=1e-5, eos_token_id=0, pad_token_id=None, bos_token_id=None,
        tie_word_embeddings=True, use_return_dict=True, output_attentions=False, output_hidden_states=False,
        top_k=50, top_p=1.0, temperature=1.0, do_sample=False, max_length=20,
    )

    def __init__(self, model_size: str = 'base', tokenizer=None, max_length: int = None, **kwargs):
        config = dict(MyTransfoXLConfig.presets[model_size])
        if tokenizer is not None:
            vsz = config['vocab_size'] = tokenizer.vocab_size
            if vsz >= 32768 * 8:
                config['cutoffs'] = [20000, 40000, 200000]
            elif vsz >= 32768:
                config['cutoffs'] = [10000]
            elif vsz >= 16384:
                config['cutoffs'] = [5000]
            elif vsz >= 1000:
                config['cutoffs'] = [1000]
            else:
                config['cutoffs'] = []
        config.update(kwargs)
        # B200 extension: arithmetic mode of the kernels ('bf16' tensor-core path, 'fp32' exact-FMA parity mode)
        self.compute_dtype = config.pop('compute_dtype', 'bf16')
        self.model_size = model_size
        merged = dict(MyTransfoXLConfig._hf_defaults)
        merged.update(config)
        for k, v in merged.items():
            setattr(self, k, v)
        self.cutoffs = list(self.cutoffs)
        self.tie_projs = [False] + [True] * len(self.cutoffs)
        self.n_token = self.vocab_size
        self.max_length_ = max_length or MyTransfoXLConfig.size2max_length[model_size]
        if self.compute_dtype not in ('bf16', 'fp32'):
            raise ValueError(f"compute_dtype must be 'bf16' or 'fp32', got {self.compute_dtype!r}")

    @property
    def model_meta(self) -> Dict[str, Any]:
        return dict(n_layer=self.n_layer, hidden_size=self.d_embed, ff_size=self.d_inner, seg_len=self.mem_len,
                    max_len=self.max_length_, vocab_size=self.vocab_size)

    # HF aliases used by Trainer / generate
    @property
    def hidden_size(self):
        return self.d_model

    @property
    def num_hidden_layers(self):
        return self.n_layer

    @property
    def num_attention_heads(self):
        return self.n_head

    def to_dict(self) -> Dict[str, Any]:
        out = {k: v for k, v in self.__dict__.items() if not k.startswith('_')}
        out['model_type'] = self.model_type
        return out

    def to_json_string(self) -> str:
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + '\n'

    def save_pretrained(self, path):
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, 'config.json'), 'w') as f:
            f.write(self.to_json_string())

    @classmethod
    def from_pretrained(cls, path):
        with open(os.path.join(path, 'config.json')) as f:
            dic = json.load(f)
        size = dic.pop('model_size', 'base')
        max_length = dic.pop('max_length_', None)
        gen_max_length = dic.pop('max_length', None)     # HF generation default, not the ctor's `max_length`
        for k in ('model_type', 'tie_projs', 'n_token'):
            dic.pop(k, None)
        cfg = cls(model_size=size, max_length=max_length, **dic)
        if gen_max_length is not None:
            cfg.max_length = gen_max_length
        return cfg


# ============================================================================= outputs
class TransfoXLLMHeadModelOutput(OrderedDict):
    """reference transformer_xl.py:81-124 — field order `losses, prediction_scores, mems, hidden_states, attentions, loss`;
    behaves like HF `ModelOutput`: attribute access, `out["loss"]`, integer indexing over the non-None fields."""
    _fields = ('losses', 'prediction_scores', 'mems', 'hidden_states', 'attentions', 'loss')

    def __init__(self, losses=None, prediction_scores=None, mems=None, hidden_states=None, attentions=None, loss=None):
        super().__init__()
        for k, v in zip(self._fields, (losses, prediction_scores, mems, hidden_states, attentions, loss)):
            object.__setattr__(self, k, v)
            if v is not None:
                super().__setitem__(k, v)

    def __getitem__(self, k):
        if isinstance(k, str):
            return super().__getitem__(k)
        return self.to_tuple()[k]

    def to_tuple(self):
        return tuple(self[k] for k in self.keys())

    @property
    def logits(self):
        # log-probabilities, not logits: reference transformer_xl.py:117-124
        return self.prediction_scores


class TxlMems(list):
    """`mems` as the model returns them: a list of L time-major (mem_len, B, d) tensors (HF contract, Appendix A.8'),
    backed by batch-major storage so that feeding them straight back into `forward` costs no layout conversion."""

    def __init__(self, bm: List[torch.Tensor]):
        super().__init__(t.transpose(0, 1) for t in bm)    # (B, M, d) -> (M, B, d) views
        self._bm = bm


# ============================================================================= parameter containers (HF state_dict names, A.8)
class _Holder(nn.Module):
    pass


def _named_param_specs(cfg):
    """(state_dict name, shape, kind) in flat-buffer order.  kind: 'mat' (cast to compute dtype) or 'vec' (kept fp32)."""
    d, di, V, H, dh = cfg.d_model, cfg.d_inner, cfg.vocab_size, cfg.n_head, cfg.d_head
    nc = len(cfg.cutoffs)
    specs = [('transformer.word_emb.emb_layers.0.weight', (V, d), 'mat')]
    if nc:      # adaptive-softmax clusters: packed right behind the embedding / output bias, so that [embedding ; cluster_weight] is ONE (V+nc, d)
        specs += [('crit.cluster_weight', (nc, d), 'mat+')]          # matrix and [out bias ; cluster_bias] one vector ('+' = no alignment gap)
    specs += [('crit.out_layers.0.bias', (V,), 'vec')]
    if nc:
        specs += [('crit.cluster_bias', (nc,), 'vec+')]
    for i in range(cfg.n_layer):
        a, f = f'transformer.layers.{i}.dec_attn.', f'transformer.layers.{i}.pos_ff.'
        specs += [
            (a + 'qkv_net.weight', (3 * H * dh, d), 'mat'), (a + 'r_net.weight', (H * dh, d), 'mat'), (a + 'o_net.weight', (d, H * dh), 'mat'),
            (a + 'r_r_bias', (H, dh), 'vec'), (a + 'r_w_bias', (H, dh), 'vec'),
            (a + 'layer_norm.weight', (d,), 'vec'), (a + 'layer_norm.bias', (d,), 'vec'),
            (f + 'CoreNet.0.weight', (di, d), 'mat'), (f + 'CoreNet.0.bias', (di,), 'vec'),
            (f + 'CoreNet.3.weight', (d, di), 'mat'), (f + 'CoreNet.3.bias', (d,), 'vec'),
            (f + 'layer_norm.weight', (d,), 'vec'), (f + 'layer_norm.bias', (d,), 'vec'),
        ]
    return specs


_ALIGN = 128   # elements; keeps every bf16 shadow matrix 256-byte aligned (TMA needs 16)


class _TxlStep(torch.autograd.Function):
    """One fused forward(+loss); backward is the hand-scheduled pipeline in engine.backward."""

    @staticmethod
    def forward(ctx, model, ids, labels_shift, mems_bm, seed, drop_p, want_logprobs, *params):
        out = engine.forward(model.config, model._W, model._E_ext, model._out_bias_ext, ids, mems_bm, labels_shift, drop_p=drop_p, seed=seed,
                             save=True, want_logprobs=want_logprobs, want_argmax=model.monitor_greedy, zero_kvm=model._zero_kvm)
        ctx.model, ctx.sv = model, out['saved']
        ctx.set_materialize_grads(False)
        model._last = out
        B, T = ids.shape
        losses, ctx.perm = model._returned_losses(out['losses'], labels_shift, B, T)
        ctx.mark_non_differentiable(*[t for t in (out['logprobs'],) if t is not None])
        lp = out['logprobs'] if out['logprobs'] is not None else torch.empty(0, device=ids.device)
        return out['loss'], losses, lp

    @staticmethod
    def backward(ctx, g_loss, g_losses, _g_lp):
        model, sv = ctx.model, ctx.sv
        if sv is None or sv.logits is None:
            raise TxlError('backward called twice on the same step (activations are released after the first backward)')
        B, T = sv.B, sv.T
        losses = sv.losses
        grow = torch.zeros(B, T, dtype=torch.float32, device=losses.device)
        if g_loss is not None:
            grow += (losses.view(B, T) != 0).to(torch.float32) * (g_loss.to(torch.float32) / sv.count)
        if g_losses is not None:
            if ctx.perm is None:
                grow[:, :T - 1] += g_losses.to(torch.float32)
            else:       # packed loss order (adaptive softmax, HF keep_order=False): route each entry back to its position; no host sync
                grow.view(-1).index_add_(0, ctx.perm.clamp(min=0), g_losses.reshape(-1).to(torch.float32) * (ctx.perm >= 0))
        gflat = model._fresh_grad_buffer()
        model._backwards_since_step += 1
        hook = model._grad_hook
        engine.backward(model.config, model._W, model._G, model._E_ext, model._gE, model._g_out_bias, sv, grow.view(-1),
                        on_layer_done=(lambda li: hook('layer', li)) if hook else None)
        if hook:
            hook('embed', -1)
        ctx.sv = None
        grads = tuple(gflat[o:o + n].view(shape) for (o, n, shape) in model._slots)
        return (None, None, None, None, None, None, None) + grads


# ============================================================================= model
class MyTransfoXLLMHeadModel(nn.Module):
    cls_name = 'TransformerXl'
    trainer_compatible = False     # absent from the reference's config => HF default False (Appendix A.1)

    def __init__(self, config: MyTransfoXLConfig, device=None):
        super().__init__()
        if len(config.cutoffs) > 4 or any(not (0 < c < config.vocab_size) for c in config.cutoffs) or sorted(set(config.cutoffs)) != list(config.cutoffs):
            raise NotImplementedError('cutoffs must be at most 4 increasing values inside (0, vocab_size)')
        if config.div_val != 1 or config.d_embed != config.d_model or config.pre_lnorm or not config.untie_r or config.attn_type != 0:
            raise NotImplementedError('only div_val=1, d_embed=d_model, post-LN, untie_r=True, attn_type=0 are on the reference path')
        if config.n_head * config.d_head != config.d_model:
            raise NotImplementedError('n_head * d_head must equal d_model (reference presets guarantee it)')
        if config.same_length and config.mem_len <= 0:
            raise NotImplementedError('same_length with mem_len<=0 masks every key')
        if getattr(config, 'dropatt', 0.0):
            raise NotImplementedError('dropatt > 0 (dropout on the attention probabilities) is not implemented: the reference keeps the HF '
                                      'default dropatt=0.0 (notebook/train/transformer-xl.ipynb:492-513)')
        self.config = config
        self._specs = _named_param_specs(config)
        self._slots, off, end = [], 0, 0
        for _, shape, kind in self._specs:
            n = 1
            for s in shape:
                n *= s
            if kind.endswith('+'):
                off = end                     # glued to the previous tensor
            self._slots.append((off, n, shape))
            end = off + n
            off = (end + _ALIGN - 1) // _ALIGN * _ALIGN
        self._flat_numel = off
        self._n_head_slots = 2 + 2 * (1 if config.cutoffs else 0)       # slots before the first layer's
        flat = torch.zeros(off, dtype=torch.float32)
        self._build_modules(flat)
        self._init_weights()
        self._engine_ready_for = None
        self._grad_hook = None
        self._last = None
        self._zeros = {}
        self._step_seed = 0
        # with monitor_greedy the LM-head kernel also emits the greedy id of every position (`last_greedy`), so the trainer's
        # `outputs.logits.argmax(-1)` (train_util_wrap.py:106) / `preprocess_logits_for_metrics` (train.py:248) needs no (B,T,V) tensor
        self.monitor_greedy = False
        self.last_greedy = None
        # optional range check (HF raises an index error for ids / labels outside the vocabulary; the kernels here would embed zeros / score
        # loss 0): with check_ranges offending ids and labels are counted on the device (the labels by the label-shift kernel), `assert_ranges_ok()` reads the counters
        self.check_ranges = False
        self._bad_labels = None
        self._bad_ids = None
        self.last_generate_path = None
        self._backwards_since_step = 0
        self._grad_accum_base = None
        if device is not None:
            self.to(device)

    # ------------------------------------------------------------------ module tree with HF names
    def _build_modules(self, flat):
        cfg = self.config
        params = {}
        for (name, shape, _), (o, n, _s) in zip(self._specs, self._slots):
            params[name] = nn.Parameter(flat[o:o + n].view(shape))
        self._flat = flat
        tr = _Holder()
        tr.word_emb = _Holder()
        emb0 = _Holder()
        emb0.weight = params['transformer.word_emb.emb_layers.0.weight']
        tr.word_emb.emb_layers = nn.ModuleList([emb0])
        tr.pos_emb = _Holder()
        tr.pos_emb.register_buffer('inv_freq', 1 / (10000 ** (torch.arange(0.0, cfg.d_model, 2.0) / cfg.d_model)))
        layers = []
        for i in range(cfg.n_layer):
            a, f = f'transformer.layers.{i}.dec_attn.', f'transformer.layers.{i}.pos_ff.'
            lay, att, ff = _Holder(), _Holder(), _Holder()
            for sub in ('qkv_net', 'r_net', 'o_net'):
                h = _Holder()
                h.weight = params[a + sub + '.weight']
                setattr(att, sub, h)
            att.r_r_bias, att.r_w_bias = params[a + 'r_r_bias'], params[a + 'r_w_bias']
            ln = _Holder()
            ln.weight, ln.bias = params[a + 'layer_norm.weight'], params[a + 'layer_norm.bias']
            att.layer_norm = ln
            c0, c3 = _Holder(), _Holder()
            c0.weight, c0.bias = params[f + 'CoreNet.0.weight'], params[f + 'CoreNet.0.bias']
            c3.weight, c3.bias = params[f + 'CoreNet.3.weight'], params[f + 'CoreNet.3.bias']
            ff.CoreNet = nn.ModuleList([c0, nn.Identity(), nn.Identity(), c3, nn.Identity()])
            ln2 = _Holder()
            ln2.weight, ln2.bias = params[f + 'layer_norm.weight'], params[f + 'layer_norm.bias']
            ff.layer_norm = ln2
            lay.dec_attn, lay.pos_ff = att, ff
            layers.append(lay)
        tr.layers = nn.ModuleList(layers)
        self.transformer = tr
        crit = _Holder()
        out0 = _Holder()
        out0.weight = params['transformer.word_emb.emb_layers.0.weight']     # tie_word_embeddings
        out0.bias = params['crit.out_layers.0.bias']
        crit.out_layers = nn.ModuleList([out0])
        if cfg.cutoffs:
            crit.cluster_weight = params['crit.cluster_weight']
            crit.cluster_bias = params['crit.cluster_bias']
        self.crit = crit
        self._param_by_name = params

    def _init_weights(self):
        """Appendix A.9: Linear/Embedding ~N(0, init_std), LayerNorm weight ~N(1, init_std), biases 0, r_*_bias ~N(0, init_std)."""
        std = self.config.init_std
        with torch.no_grad():
            for name, p in self._param_by_name.items():
                if name.endswith('layer_norm.weight'):
                    p.normal_(1.0, std)
                elif (name.endswith('.bias') and 'r_' not in name.rsplit('.', 1)[-1]) or name == 'crit.cluster_bias':
                    p.zero_()
                else:
                    p.normal_(0.0, std)

    # ------------------------------------------------------------------ HF-protocol helpers
    def get_output_embeddings(self):
        return self.crit.out_layers[0]

    def get_input_embeddings(self):
        return self.transformer.word_emb.emb_layers[0]

    def num_parameters(self, only_trainable=False):
        return sum(p.numel() for p in self.parameters())

    def reset_memory_length(self, mem_len):
        self.config.mem_len = mem_len

    def init_mems(self, bsz):
        """HF `TransfoXLModel.init_mems` (Appendix A.2-1): zero hidden-state mems, attended like real ones."""
        cfg = self.config
        if cfg.mem_len <= 0:
            return None
        dev = self._flat.device
        return TxlMems([torch.zeros(bsz, cfg.mem_len, cfg.d_model, dtype=self._act_dtype(), device=dev) for _ in range(cfg.n_layer)])

    @staticmethod
    def _reorder_cache(mems, beam_idx):
        return [layer_past.index_select(1, beam_idx.to(layer_past.device)) for layer_past in mems]

    def _act_dtype(self):
        return torch.bfloat16 if self.config.compute_dtype == 'bf16' else torch.float32

    # ------------------------------------------------------------------ engine state
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._engine_ready_for = None          # parameters may have moved: re-flatten lazily
        return out

    def _ensure_engine(self):
        """(Re)build flat fp32 storage on the parameters' device, the compute-dtype shadow, and per-layer views."""
        p0 = self._param_by_name['transformer.word_emb.emb_layers.0.weight']
        dev = p0.device
        if dev.type != 'cuda':
            raise TxlError('MyTransfoXLLMHeadModel runs on a B200 only: move the model to CUDA (there is no CPU fallback; '
                           'the CPU restatement lives in oracle/ and is test infrastructure)')
        key = (dev, p0.data_ptr(), self.config.compute_dtype)
        if self._engine_ready_for == key:
            return
        ops.device_ok()
        flat_ok = self._flat.device == dev and all(
            p.data_ptr() == self._flat.data_ptr() + 4 * o for p, (o, n, s) in zip(self._param_by_name.values(), self._slots))
        if not flat_ok:
            flat = torch.zeros(self._flat_numel, dtype=torch.float32, device=dev)
            with torch.no_grad():
                for p, (o, n, shape) in zip(self._param_by_name.values(), self._slots):
                    flat[o:o + n].copy_(p.detach().reshape(-1).to(torch.float32))
                    p.data = flat[o:o + n].view(shape)
            self._flat = flat
        self._gflat = None
        self._shadow = torch.empty(self._flat_numel, dtype=torch.bfloat16, device=dev) if self.config.compute_dtype == 'bf16' else None
        self._shadow_version = None
        self._engine_ready_for = (dev, p0.data_ptr(), self.config.compute_dtype)
        self._zeros = {}
        self._bind_views()

    def _view(self, buf, i):
        o, n, shape = self._slots[i]
        return buf[o:o + n].view(shape)

    def _bind_views(self):
        """Per-layer weight views: matrices from the compute-dtype buffer, vectors from fp32 master."""
        mat_src = self._shadow if self._shadow is not None else self._flat
        names = [s[0] for s in self._specs]
        idx = {n: i for i, n in enumerate(names)}

        def mat(n):
            return self._view(mat_src, idx[n])

        def vec(n):
            return self._view(self._flat, idx[n]).reshape(-1)

        self._E = mat('transformer.word_emb.emb_layers.0.weight')
        self._out_bias = vec('crit.out_layers.0.bias')
        # LM-head operands: with adaptive-softmax clusters the cluster rows / biases sit right behind, so one GEMM yields token + cluster logits
        V, d, nc = self.config.vocab_size, self.config.d_model, len(self.config.cutoffs)
        o_e, o_b = self._slots[idx['transformer.word_emb.emb_layers.0.weight']][0], self._slots[idx['crit.out_layers.0.bias']][0]
        self._E_ext = mat_src[o_e:o_e + (V + nc) * d].view(V + nc, d)
        self._out_bias_ext = self._flat[o_b:o_b + V + nc]
        self._W = []
        for i in range(self.config.n_layer):
            a, f = f'transformer.layers.{i}.dec_attn.', f'transformer.layers.{i}.pos_ff.'
            self._W.append(engine.LayerW(
                qkv=mat(a + 'qkv_net.weight'), r=mat(a + 'r_net.weight'), o=mat(a + 'o_net.weight'),
                rrb=vec(a + 'r_r_bias'), rwb=vec(a + 'r_w_bias'), ln1_w=vec(a + 'layer_norm.weight'), ln1_b=vec(a + 'layer_norm.bias'),
                w1=mat(f + 'CoreNet.0.weight'), b1=vec(f + 'CoreNet.0.bias'), w2=mat(f + 'CoreNet.3.weight'), b2=vec(f + 'CoreNet.3.bias'),
                ln2_w=vec(f + 'layer_norm.weight'), ln2_b=vec(f + 'layer_norm.bias')))
        self._idx = idx

    def _bind_grad_views(self, gflat):
        idx = self._idx

        def gv(n):
            return self._view(gflat, idx[n])

        V, d, nc = self.config.vocab_size, self.config.d_model, len(self.config.cutoffs)
        o_e, o_b = self._slots[idx['transformer.word_emb.emb_layers.0.weight']][0], self._slots[idx['crit.out_layers.0.bias']][0]
        self._gE = gflat[o_e:o_e + (V + nc) * d].view(V + nc, d)          # [embedding ; cluster_weight] gradients
        self._g_out_bias = gflat[o_b:o_b + V + nc]
        self._G = []
        for i in range(self.config.n_layer):
            a, f = f'transformer.layers.{i}.dec_attn.', f'transformer.layers.{i}.pos_ff.'
            self._G.append(engine.LayerW(
                qkv=gv(a + 'qkv_net.weight'), r=gv(a + 'r_net.weight'), o=gv(a + 'o_net.weight'),
                rrb=gv(a + 'r_r_bias').reshape(-1), rwb=gv(a + 'r_w_bias').reshape(-1),
                ln1_w=gv(a + 'layer_norm.weight'), ln1_b=gv(a + 'layer_norm.bias'),
                w1=gv(f + 'CoreNet.0.weight'), b1=gv(f + 'CoreNet.0.bias'), w2=gv(f + 'CoreNet.3.weight'), b2=gv(f + 'CoreNet.3.bias'),
                ln2_w=gv(f + 'layer_norm.weight'), ln2_b=gv(f + 'layer_norm.bias')))

    def _fresh_grad_buffer(self):
        """Zeroed flat fp32 gradient buffer.  A new one is allocated whenever autograd may still hold views of the previous
        one as `.grad` (gradient accumulation), so earlier gradients are never clobbered."""
        reuse = self._gflat is not None
        if reuse:
            lo, hi = self._gflat.data_ptr(), self._gflat.data_ptr() + 4 * self._flat_numel
            for p in self._param_by_name.values():
                if p.grad is not None and lo <= p.grad.data_ptr() < hi:
                    reuse = False
                    break
        if reuse:
            self._gflat.zero_()
        else:
            if self._gflat is not None and self._grad_accum_base is None:
                self._grad_accum_base = self._gflat      # `.grad`s alias it: autograd accumulates the later backwards INTO this buffer
            self._gflat = torch.zeros(self._flat_numel, dtype=torch.float32, device=self._flat.device)
        self._bind_grad_views(self._gflat)
        return self._gflat

    def zero_grad(self, set_to_none: bool = True):
        super().zero_grad(set_to_none=set_to_none)
        self._grad_accum_base = None
        self._backwards_since_step = 0

    def flat_grads(self):
        """The flat fp32 gradient the optimiser must consume = what the parameters' `.grad` hold (HF Trainer semantics: gradient
        accumulation over several backwards, external clipping / unscaling of `.grad`).  Normally the `.grad`s are views of one flat buffer
        (the first backward's; later backwards are accumulated into it by autograd) and that buffer is returned as is; if any `.grad` was
        replaced by the user, the gradients are gathered into a scratch buffer."""
        params = list(self._param_by_name.values())
        g0 = params[0].grad
        if g0 is None:
            raise RuntimeError('no gradients: call loss.backward() first')
        base = g0.data_ptr() - 4 * self._slots[0][0]
        for cand in (self._grad_accum_base, self._gflat):
            if cand is not None and cand.data_ptr() == base and all(
                    p.grad is not None and p.grad.data_ptr() == base + 4 * o for p, (o, n, s) in zip(params, self._slots)):
                return cand
        scratch = torch.zeros(self._flat_numel, dtype=torch.float32, device=self._flat.device)
        for p, (o, n, shape) in zip(params, self._slots):
            if p.grad is not None:
                scratch[o:o + n].copy_(p.grad.reshape(-1))
        return scratch

    def layer_param_ranges(self):
        """[(start, end)] element ranges of the flat buffers: index 0 = embedding + output bias, 1.. = layers (for gradient buckets)."""
        per, h0 = 13, self._n_head_slots
        out = [(self._slots[0][0], self._slots[h0][0] if len(self._slots) > h0 else self._flat_numel)]
        for i in range(self.config.n_layer):
            s = self._slots[h0 + per * i][0]
            e = self._slots[h0 + per * (i + 1)][0] if i + 1 < self.config.n_layer else self._flat_numel
            out.append((s, e))
        return out

    def _refresh_shadow(self):
        if self._shadow is None:
            return
        ver = tuple(p._version for p in self._param_by_name.values())
        if ver != self._shadow_version:
            ops.cast_f32_to_bf16(self._flat, self._shadow)
            self._shadow_version = ver

    def _zero_kvm(self, rows, cols, dtype, dev):
        key = (rows, cols, dtype)
        z = self._zeros.get(key)
        if z is None:
            self._zeros = {key: torch.zeros(rows, cols, dtype=dtype, device=dev)}
            z = self._zeros[key]
        return z

    # ------------------------------------------------------------------ mems conversion at the boundary
    def _mems_to_bm(self, mems, B):
        if mems is None:
            return None
        if isinstance(mems, TxlMems):
            bm = mems._bm
        else:
            mems = list(mems)
            if len(mems) != self.config.n_layer:
                raise ValueError(f'mems must hold {self.config.n_layer} tensors, got {len(mems)}')
            bm = None
        act = self._act_dtype()
        out = []
        for i in range(self.config.n_layer):
            if bm is not None:
                t = bm[i]
                t = t if (t.dtype == act and t.is_contiguous()) else t.to(act).contiguous()
            else:
                m = mems[i]
                if m.dim() != 3 or m.shape[1] != B or m.shape[2] != self.config.d_model:
                    raise ValueError(f'mems[{i}] must be (mlen, {B}, {self.config.d_model}), got {tuple(m.shape)}')
                if not m.is_cuda:
                    m = m.to(self._flat.device)
                t = ops.tm_to_bm(m if m.dtype in (torch.float32, torch.bfloat16) else m.float(), act)
            out.append(t)
        return out

    def _new_mems(self, mems_bm, hid_in, B, T):
        """HF `_update_mems` (Appendix A.8'): last mem_len rows of cat([mems, layer input])."""
        ML = self.config.mem_len
        if ML <= 0:
            return None
        d = self.config.d_model
        new = []
        for i, h in enumerate(hid_in):
            h = h.view(B, T, d)
            if T >= ML:
                new.append(h[:, T - ML:].detach())
            else:
                old = mems_bm[i] if mems_bm is not None else torch.zeros(B, ML, d, dtype=h.dtype, device=h.device)
                new.append(torch.cat([old, h], dim=1)[:, -ML:].detach())
        return TxlMems(new)

    # ------------------------------------------------------------------ forward (reference transformer_xl.py:130-221)
    def forward(self, key_scores=None, input_ids: Optional[torch.LongTensor] = None, mems=None, head_mask=None, inputs_embeds=None,
                labels: Optional[torch.LongTensor] = None, output_attentions=None, output_hidden_states=None, return_dict=None):
        return_dict = return_dict if return_dict is not None else self.config.use_return_dict
        if input_ids is None and inputs_embeds is None:
            raise ValueError('You have to specify either input_ids or inputs_embeds')
        if inputs_embeds is not None or head_mask is not None or output_attentions or output_hidden_states:
            raise NotImplementedError('inputs_embeds / head_mask / output_attentions / output_hidden_states are not on the '
                                      'reference hot path (never passed by train.py / eval.py) and are not implemented')
        self._ensure_engine()
        self._refresh_shadow()
        dev = self._flat.device
        ids = input_ids.to(dev, non_blocking=True)
        if ids.dtype != torch.int64:
            ids = ids.long()
        ids = ids.contiguous()
        bsz, tgt_len = ids.shape
        if self.check_ranges:          # counted on the device (no host wait); HF's embedding lookup would raise for these
            bad = ((ids < 0) | (ids >= self.config.vocab_size)).sum(dtype=torch.int32).view(1)
            self._bad_ids = bad if self._bad_ids is None else self._bad_ids + bad
        labels_shift = None
        if labels is not None:
            if tuple(labels.shape) != (bsz, tgt_len):
                raise RuntimeError('Input and labels should have the same size in the batch dimension.')
            # reference :176-182 — in-place fix-up of an all-pad first row.  Host labels: checked on the host (no device sync);
            # device labels: checked and patched by the label-shift kernel itself, so the training loop never waits for the GPU here.
            if not labels.is_cuda:
                miss_valid_label = labels[0, 1:].sum() == (labels.size(1) - 1) * -100
                if miss_valid_label:
                    labels[0, 1] = self.config.eos_token_id
            lab = labels.to(dev, non_blocking=True)
            direct = lab.dtype == torch.int64 and lab.stride(1) == 1
            lab_k = lab if direct else lab.long().contiguous()
            if self.check_ranges and self._bad_labels is None:
                self._bad_labels = torch.zeros(1, dtype=torch.int32, device=dev)
            labels_shift = ops.shift_labels(lab_k, self.config.eos_token_id if self.config.eos_token_id is not None else 0,
                                            self.config.vocab_size, self._bad_labels if self.check_ranges else None)
            if labels.is_cuda and not direct:
                labels[0, 1] = lab_k[0, 1]          # carry the in-place side effect back to the caller's tensor (device-side copy)
        mems_bm = self._mems_to_bm(mems, bsz)
        in_eval = not self.training
        want_logprobs = labels is None or in_eval
        drop_p = float(self.config.dropout) if self.training else 0.0
        self._step_seed += 1
        seed = (torch.initial_seed() * 1000003 + self._step_seed) & ((1 << 62) - 1)
        need_grad = torch.is_grad_enabled() and labels is not None and any(p.requires_grad for p in self._param_by_name.values())
        if need_grad:
            loss, losses, lp = _TxlStep.apply(self, ids, labels_shift, mems_bm, seed, drop_p, want_logprobs, *self._param_by_name.values())
            out = self._last
            logprobs = lp if want_logprobs else None
        else:
            with torch.no_grad():
                out = engine.forward(self.config, self._W, self._E_ext, self._out_bias_ext, ids, mems_bm, labels_shift, drop_p=drop_p, seed=seed,
                                     save=False, want_logprobs=want_logprobs, want_argmax=self.monitor_greedy, zero_kvm=self._zero_kvm)
            loss = out['loss']
            losses = self._returned_losses(out['losses'], labels_shift, bsz, tgt_len)[0] if labels is not None else None
            logprobs = out['logprobs']
        new_mems = self._new_mems(mems_bm, out['hid_in'], bsz, tgt_len)
        # greedy ids of every position, straight from the LM-head kernel (what `logits.argmax(-1)` is in train_util_wrap.py:106 / train.py:248)
        self.last_greedy = out['argmax'].view(bsz, tgt_len) if out.get('argmax') is not None else None
        self._last = None
        prediction_scores = logprobs.view(bsz, tgt_len, -1) if want_logprobs else ()
        if labels is None:
            losses, loss = None, None
        if not return_dict:
            tail = (new_mems,)
            if self.trainer_compatible:
                output = (prediction_scores, losses) if losses is not None else (prediction_scores,)
                output += tail
                return ((loss,) + output) if loss is not None else output
            output = (prediction_scores, *tail)
            output = ((losses,) + output) if losses is not None else output
            return (output + (loss,)) if loss is not None else output
        return TransfoXLLMHeadModelOutput(loss=loss, prediction_scores=prediction_scores, losses=losses, mems=new_mems,
                                          hidden_states=None, attentions=None)

    def assert_ranges_ok(self):
        """Two host reads: raises if any input id or label seen since `check_ranges = True` was outside [0, vocab_size) (labels: and not -100)."""
        if self._bad_ids is not None and int(self._bad_ids.item()) != 0:
            raise IndexError(f'{int(self._bad_ids.item())} input id(s) outside [0, {self.config.vocab_size}) were passed to forward')
        if self._bad_labels is not None and int(self._bad_labels.item()) != 0:
            raise IndexError(f'{int(self._bad_labels.item())} label(s) outside [0, {self.config.vocab_size}) were passed to forward')

    def mark_params_dirty(self):
        """Call after writing parameters through raw storage (`model._flat`, `.data`, a broadcast): such writes do not bump the autograd
        version counters `_refresh_shadow` keys on, so the bf16 shadow would silently go stale."""
        self._shadow_version = None

    def _returned_losses(self, pos_losses, labels_shift, B, T):
        """`losses` as the reference returns them: (B, T-1) in position order without clusters; with adaptive-softmax clusters HF's criterion
        (built with keep_order=False) returns them PACKED cluster by cluster with the ignored labels as trailing zeros (Appendix A.6) — same
        multiset, so `losses[losses != 0].mean()` is unchanged.  Returns (losses, perm or None)."""
        if not self.config.cutoffs:
            return pos_losses.view(B, T)[:, :T - 1], None
        return ops.pack_losses(pos_losses, labels_shift, B, T, self.config.vocab_size, self.config.cutoffs)

    def ntp_acc_counts(self, labels, out=None):
        """int64[2] device tensor (matches, non-pad positions) of next-token prediction for the last forward (needs monitor_greedy=True):
        preds[:, :-1] vs labels[:, 1:] over labels != -100 — the reference's `ntp_acc` = out[0] / out[1]
        (train_util_wrap.py:113-120, train.py:279-284).  Accumulates into `out` when given; no host sync."""
        if self.last_greedy is None:
            raise TxlError('ntp_acc_counts: set model.monitor_greedy = True before the forward')
        lab = labels.to(self.last_greedy.device, non_blocking=True).long().contiguous()
        if tuple(lab.shape) != tuple(self.last_greedy.shape):
            raise RuntimeError('labels must have the shape of the last input_ids')
        return ops.ntp_acc(self.last_greedy, lab, out)

    # ------------------------------------------------------------------ generation surface (reference :223-241)
    def prepare_inputs_for_generation(self, input_ids, past=None, **model_kwargs):
        """HF generation hook with the 4.25-era `past` keyword (reference transformer_xl.py:223-241): without a cache the whole prompt is fed,
        with one only the newest token plus the mems.  Contrastive search hands `past` back as one list of tensors per layer; those are
        stacked again, in place as the reference does."""
        if not past:
            return {'input_ids': input_ids}
        if not isinstance(past, list):
            raise TypeError(f'past must be a list of per-layer mems, got {type(past).__name__}')
        for layer, entry in enumerate(past):
            if isinstance(entry, (list, tuple)):
                if not all(isinstance(t, torch.Tensor) for t in entry):
                    raise TypeError('nested past entries must be tensors')
                past[layer] = torch.stack(list(entry), dim=0)
        return {'mems': past, 'input_ids': input_ids[:, -1:]}

    def generate(self, input_ids=None, **kwargs):
        from .generation import generate
        return generate(self, input_ids, **kwargs)

    # ------------------------------------------------------------------ persistence (HF layout: config.json + pytorch_model.bin)
    def save_pretrained(self, path):
        os.makedirs(path, exist_ok=True)
        self.config.save_pretrained(path)
        sd = {k: v.detach().cpu().clone() for k, v in self.state_dict().items() if not k.startswith('_')}
        torch.save(sd, os.path.join(path, 'pytorch_model.bin'))

    @classmethod
    def from_pretrained(cls, path, **kw):
        cfg = MyTransfoXLConfig.from_pretrained(path)
        for k, v in kw.items():
            setattr(cfg, k, v)
        model = cls(cfg)
        sd = torch.load(os.path.join(path, 'pytorch_model.bin'), map_location='cpu')
        model.load_state_dict(sd)
        model.eval()
        return model
