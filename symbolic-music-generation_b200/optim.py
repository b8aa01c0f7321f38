"""Fused optimiser step for the flat parameter buffer: global-norm clip + AdamW + bf16 shadow refresh in one pass.

Mirrors what the reference's HF `Trainer` does around the hot path (musicnlp/trainer/train.py:166-190: AdamW betas
(0.9, 0.999), eps 1e-8, max_grad_norm 1, cosine schedule with warm-up; decay is skipped for LayerNorm weights and every
parameter whose name contains "bias", which includes r_w_bias / r_r_bias — HF `get_parameter_names` rule).  SURVEY §8f-1.
"""
from __future__ import annotations

import math

import torch

from . import ops


def cosine_with_warmup(step, total_steps, warmup_steps, base_lr):
    if step < warmup_steps:
        return base_lr * step / max(1, warmup_steps)
    prog = (step - warmup_steps) / max(1, total_steps - warmup_steps)
    return base_lr * max(0.0, 0.5 * (1.0 + math.cos(math.pi * prog)))


class FusedAdamW:
    def __init__(self, model, lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_grad_norm=1.0):
        model._ensure_engine()
        self.model = model
        self.lr, self.betas, self.eps, self.wd, self.max_grad_norm = lr, betas, eps, weight_decay, max_grad_norm
        n, dev = model._flat_numel, model._flat.device
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        mask = torch.zeros(n, dtype=torch.uint8)
        for (name, shape, kind), (o, cnt, _s) in zip(model._specs, model._slots):
            if 'bias' not in name and 'layer_norm' not in name:
                mask[o:o + cnt] = 1
        self.decay_mask = mask.to(dev)
        self.t = 0
        self._sq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.last_grad_norm = None

    def step(self, lr=None):
        """Consumes the parameters' `.grad` (`model.flat_grads()`): with gradient accumulation (several backwards before one step, as HF
        Trainer's `gradient_accumulation_steps` does) autograd has summed the micro-batches into the first backward's flat buffer, which is
        what is read here — not the buffer of the last backward alone."""
        model = self.model
        g = model.flat_grads()
        self.t += 1
        scale = None
        if self.max_grad_norm and self.max_grad_norm > 0:
            self._sq.zero_()
            ops.sumsq(g, self._sq)
            norm = self._sq.sqrt()
            self.last_grad_norm = norm
            scale = (self.max_grad_norm / (norm + 1e-6)).clamp(max=1.0)        # torch.nn.utils.clip_grad_norm_ rule
        ops.adamw_step(model._flat, g, self.m, self.v, self.decay_mask, self.lr if lr is None else lr, self.betas[0], self.betas[1], self.eps,
                       self.wd, self.t, scale, model._shadow)
        # parameters were updated through raw pointers: autograd version counters did not move and the shadow is already fresh

    def zero_grad(self, set_to_none=True):
        self.model.zero_grad(set_to_none=True)
