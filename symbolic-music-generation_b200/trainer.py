"""The per-step contract of the reference's HF-Trainer glue around the hot path, on the device:

  `MyTrainer.compute_loss` (musicnlp/util/train/train_util_wrap.py:88-144): `outputs = model(**inputs)`, `loss = outputs["loss"]`, and — when
  training with `monitor_ntp_acc` — next-token accuracy of `outputs.logits.argmax(-1)` against the shifted, pad-masked labels;
  HF `Trainer.training_step` + optimizer: backward, `clip_grad_norm_(max_grad_norm=1)`, AdamW (betas .9/.999, eps 1e-8, no decay on LayerNorm
  weights and on every parameter whose name contains "bias"), cosine schedule with `ceil(warmup_ratio * total_steps)` warm-up steps
  (musicnlp/trainer/train.py:166-190 defaults, :79-110 presets: lr 3e-4, weight_decay 1e-2, warmup_ratio 0.1).

Not a Trainer re-implementation (callbacks, evaluation loops, checkpoint rotation and logging sinks are out of scope): the arithmetic of one
optimisation step and its logged quantities, so that the reference's loss / ntp_acc trajectory can be reproduced.  The greedy ids come fused
from the LM-head kernel (`model.monitor_greedy`), the accuracy counts stay on the device until `log()` reads them.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional

import torch

from . import optim
from .dist import GradBucketer


def warmup_steps(total_steps: int, warmup_ratio: float) -> int:
    return math.ceil(total_steps * warmup_ratio)          # HF TrainingArguments.get_warmup_steps


class TxlTrainer:
    def __init__(self, model, total_steps: int, learning_rate: float = 3e-4, weight_decay: float = 1e-2, warmup_ratio: float = 0.1,
                 max_grad_norm: float = 1.0, betas=(0.9, 0.999), eps: float = 1e-8, monitor_ntp_acc: bool = True, bucket_mb: float = 25.0):
        self.model = model
        self.total_steps, self.base_lr = int(total_steps), float(learning_rate)
        self.warmup = warmup_steps(self.total_steps, warmup_ratio)
        self.opt = optim.FusedAdamW(model, lr=learning_rate, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        self.bucketer = GradBucketer(model, bucket_mb=bucket_mb) if torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1 else None
        if self.bucketer is not None:
            # DDP's construction-time contract: every replica starts from rank 0's parameters (replicas built under different seeds or from
            # different checkpoints would otherwise diverge silently while their averaged gradients look healthy)
            model._ensure_engine()
            torch.distributed.broadcast(model._flat, 0)
            model.mark_params_dirty()
        self.monitor_ntp_acc = bool(monitor_ntp_acc)
        self.step = 0                                          # optimizer steps taken == scheduler steps taken
        self._acc = torch.zeros(2, dtype=torch.int64, device=model._flat.device) if model._flat is not None and model._flat.is_cuda else None
        self.history: List[Dict[str, float]] = []

    def lr_at(self, step: int) -> float:
        """Learning rate used BY optimizer step number `step` (0-based): HF steps the scheduler after the optimizer, so step 0 runs at lambda(0)."""
        return optim.cosine_with_warmup(step, self.total_steps, self.warmup, self.base_lr)

    def compute_loss(self, inputs: Dict[str, torch.Tensor], return_outputs: bool = False):
        model = self.model
        monitor = self.monitor_ntp_acc and model.training and 'labels' in inputs
        model.monitor_greedy = monitor
        outputs = model(**inputs)
        if monitor:
            if self._acc is None:
                self._acc = torch.zeros(2, dtype=torch.int64, device=model._flat.device)
            model.ntp_acc_counts(inputs['labels'], self._acc)
        if isinstance(outputs, dict) and 'loss' not in outputs:
            raise ValueError('The model did not return a loss from the inputs, only the following keys: ' + ','.join(outputs.keys()))
        loss = outputs['loss'] if isinstance(outputs, dict) else outputs[0]
        return (loss, outputs) if return_outputs else loss

    def training_step(self, inputs: Dict[str, torch.Tensor]) -> torch.Tensor:
        self.model.train()
        loss = self.compute_loss(inputs)
        loss.backward()
        self.opt.step(lr=self.lr_at(self.step))
        self.opt.zero_grad()
        self.step += 1
        return loss.detach()

    def log(self, loss: torch.Tensor) -> Dict[str, float]:
        """One host read per call: loss, lr of the step just taken, gradient norm before clipping, ntp_acc since the last log.
        (HF Trainer logs `lr_scheduler.get_last_lr()` AFTER the scheduler stepped, i.e. the NEXT step's rate: `learning_rate_next` holds that.)"""
        entry = dict(step=self.step, loss=float(loss.item()), learning_rate=self.lr_at(self.step - 1), learning_rate_next=self.lr_at(self.step),
                     grad_norm=float(self.opt.last_grad_norm.item()) if self.opt.last_grad_norm is not None else float('nan'))
        if self.monitor_ntp_acc and self._acc is not None:
            hit, cnt = self._acc.tolist()
            entry['ntp_acc'] = hit / cnt if cnt else float('nan')
            self._acc.zero_()
        self.history.append(entry)
        return entry

    def train(self, batches: Iterable, logging_steps: int = 1, max_steps: Optional[int] = None) -> List[Dict[str, float]]:
        """`batches` yields dicts with `input_ids` / `labels` (e.g. built from io.DeviceBatchPipeline)."""
        limit = self.total_steps if max_steps is None else min(max_steps, self.total_steps)
        for inputs in batches:
            if self.step >= limit:
                break
            if not isinstance(inputs, dict):
                inputs = dict(input_ids=inputs[0], labels=inputs[1])
            loss = self.training_step(inputs)
            if logging_steps and self.step % logging_steps == 0:
                self.log(loss)
        return self.history
