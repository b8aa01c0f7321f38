"""Data-parallel training plumbing: one process per GPU, bucketed gradient all-reduce launched from inside the
hand-scheduled backward so that NCCL traffic over NVLink overlaps the remaining layers' compute.

The reference runs single-GPU through HF Trainer (DDP would be its multi-GPU path: 25 MB buckets, gradient averaging);
this reproduces DDP's arithmetic — every replica computes `losses[losses != 0].mean()` on its shard, gradients are averaged.
Batched sampling needs no collective: sequences are sharded, each rank owns its mems.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class GradBucketer:
    """Installs `model._grad_hook`.  Buckets are contiguous slices of the flat gradient buffer: layers are laid out in order,
    backward finishes them last-to-first, so a bucket = a run of adjacent layers; the embedding/bias slice goes last."""

    def __init__(self, model, bucket_mb: float = 25.0, group=None):
        model._ensure_engine()
        self.model, self.group = model, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        ranges = model.layer_param_ranges()          # [emb+bias, layer0, ..., layerL-1]
        cap = int(bucket_mb * 1024 * 1024 / 4)
        L = len(ranges) - 1
        self.buckets = []                            # (trigger_layer, start, end)
        hi = L
        while hi >= 1:
            lo = hi
            size = ranges[hi][1] - ranges[hi][0]
            while lo - 1 >= 1 and size + (ranges[lo - 1][1] - ranges[lo - 1][0]) <= cap:
                lo -= 1
                size += ranges[lo][1] - ranges[lo][0]
            self.buckets.append((lo - 1, ranges[lo][0], ranges[hi][1]))   # fires when layer index lo-1 (0-based) is done
            hi = lo - 1
        self.emb_range = ranges[0]
        self._by_layer = {b[0]: b for b in self.buckets}
        self._handles = []
        self.launched = 0
        model._grad_hook = self

    def __call__(self, kind, li):
        if self.world == 1:
            return
        g = self.model._gflat
        if kind == 'layer':
            b = self._by_layer.get(li)
            if b is not None:
                self._launch(g[b[1]:b[2]])
        else:
            self._launch(g[self.emb_range[0]:self.emb_range[1]])
            for h in self._handles:
                h.wait()
            self._handles = []

    def _launch(self, t):
        t.div_(self.world) if not _has_avg() else None
        op = dist.ReduceOp.AVG if _has_avg() else dist.ReduceOp.SUM
        self._handles.append(dist.all_reduce(t, op=op, group=self.group, async_op=True))
        self.launched += 1

    def bucket_sizes_mb(self):
        return [round((e - s) * 4 / 2 ** 20, 2) for _, s, e in self.buckets] + [round((self.emb_range[1] - self.emb_range[0]) * 4 / 2 ** 20, 2)]


def _has_avg():
    return dist.is_initialized() and dist.get_backend() == 'nccl'


def shard_sequences(n_total: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of `n_total` independent sequences for batched sampling; no collective on the data path."""
    per, rem = divmod(n_total, world)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)
