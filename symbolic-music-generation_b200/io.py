"""Formats either side of the hot path: what reaches `forward` and what leaves `generate`.

In:  the reference feeds the model through HF `DataCollatorForLanguageModeling(tokenizer, mlm=False)` (musicnlp/trainer/train.py:360) over
     examples the tokenizer already padded to `max_length` (musicnlp/preprocess/dataset.py:361): `labels = input_ids` with pad -> -100.
     At > 1 M tokens/s a host-side collator plus a synchronous copy would stall the step, so `DeviceBatchPipeline` stages the id matrix in
     pinned memory, copies it on its own stream one batch ahead, and builds the labels on the device (`txl_clm_labels`).
Out: `MusicGenerator._truncate_last_bar` (musicnlp/trainer/eval.py:178-185) cuts a generated sequence at its last start-of-bar token before
     decoding; `truncate_last_bar` finds the cut points of a whole batch on the device (`txl_last_index_of`).
"""
from __future__ import annotations

from typing import Iterable, Iterator, List, Tuple

import torch

from ._lib import check, load, ptr, stream_ptr

PT_LOSS_PAD = -100


def clm_labels(input_ids: torch.Tensor, pad_token_id: int) -> torch.Tensor:
    """labels of the causal-LM collator: a copy of `input_ids` (int64, CUDA) with `pad_token_id` replaced by -100."""
    if input_ids.dtype != torch.int64:
        raise TypeError('input_ids must be int64')
    ids = input_ids.contiguous()
    labels = torch.empty_like(ids)
    if ids.numel():
        check(load().txl_clm_labels(ptr(ids), ptr(labels), ids.numel(), int(pad_token_id), stream_ptr()), 'clm_labels')
    return labels


def last_index_of(ids: torch.Tensor, token: int) -> torch.Tensor:
    """int64 [B]: last column of `token` in every row of the int64 CUDA matrix `ids` (rows may be strided), -1 where absent."""
    if ids.dtype != torch.int64 or ids.dim() != 2 or ids.stride(1) != 1:
        raise TypeError('ids must be an int64 [B, T] matrix with unit column stride')
    out = torch.empty(ids.shape[0], dtype=torch.int64, device=ids.device)
    if ids.shape[0] and ids.shape[1]:
        check(load().txl_last_index_of(ptr(ids), ids.stride(0), ids.shape[0], ids.shape[1], int(token), ptr(out), stream_ptr()), 'last_index_of')
    else:
        out.fill_(-1)
    return out


def truncate_last_bar(ids: torch.Tensor, sob_token_id: int) -> List[List[int]]:
    """`[row[:last start-of-bar].tolist() for row in ids]` — reference eval.py:178-185 for a batch; like the reference it refuses a row
    without any start-of-bar token."""
    cut = last_index_of(ids, sob_token_id).tolist()
    if any(c < 0 for c in cut):
        raise AssertionError('No start of bar token found when truncate_to_sob enabled')
    host = ids.cpu()
    return [host[b, :c].tolist() for b, c in enumerate(cut)]


class DeviceBatchPipeline:
    """Iterates `(input_ids, labels)` CUDA batches from an iterable of host id matrices (int64 [B, T], padded with `pad_token_id`).

    Two pinned staging buffers and a copy stream keep the next batch's host-to-device copy and label kernel in flight while the current
    batch trains; the consumer's stream waits on an event, never on the host."""

    def __init__(self, batches: Iterable[torch.Tensor], pad_token_id: int, device=None):
        self.batches, self.pad = batches, int(pad_token_id)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self._pinned = [None, None]

    def _stage(self, slot: int, host_ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.cuda.Event]:
        host_ids = torch.as_tensor(host_ids)
        if host_ids.dtype != torch.int64:
            host_ids = host_ids.long()
        buf = self._pinned[slot]
        if buf is None or buf.shape != host_ids.shape:
            buf = self._pinned[slot] = torch.empty(host_ids.shape, dtype=torch.int64).pin_memory()
        buf.copy_(host_ids)
        with torch.cuda.stream(self.copy_stream):
            dev_ids = buf.to(self.device, non_blocking=True)
            labels = clm_labels(dev_ids, self.pad)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return dev_ids, labels, ev

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        it = iter(self.batches)
        slot = 0
        try:
            nxt = self._stage(slot, next(it))
        except StopIteration:
            return
        reuse = [None, None]                      # event after which a pinned slot may be overwritten
        while nxt is not None:
            cur, cur_slot = nxt, slot
            slot ^= 1
            try:
                host = next(it)
                if reuse[slot] is not None:
                    reuse[slot].synchronize()     # the copy that read this pinned buffer two batches ago has finished
                nxt = self._stage(slot, host)
            except StopIteration:
                nxt = None
            ids, labels, ev = cur
            torch.cuda.current_stream().wait_event(ev)
            ids.record_stream(torch.cuda.current_stream())
            labels.record_stream(torch.cuda.current_stream())
            reuse[cur_slot] = ev
            yield ids, labels
