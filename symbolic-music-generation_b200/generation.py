"""`generate` surface of the boundary: greedy search and sampling (temperature / top-k / top-p, renormalised), the
modes `musicnlp/trainer/eval.py:277-333` drives, following HF 4.25 `GenerationMixin.greedy_search` / `sample`
(SURVEY.md Appendix A.7): one forward per new token fed through `prepare_inputs_for_generation(input_ids, past=mems)`,
log-prob scores of the last position, warpers in HF order, eos/pad bookkeeping, stop at `max_length`.

The device-resident decode step (ring cache, CUDA graph; decode.py) is the DEFAULT for every call the reference makes
(`eval.py:333`: HF kwargs only).  Sampling draws are keyed (seed, global sequence index, step): when the caller passes no
`seed`, one is drawn from torch's global generator, so `torch.manual_seed(s)` makes a run reproducible exactly as it does for
HF's `torch.multinomial`, and consecutive calls differ.  Passing a `torch.Generator` (`generator=`) or asking for per-step
scores selects the one-forward-per-token host loop.  Any batch size: more than 64 sequences are decoded as further sequence groups.

`typical_p` and `repetition_penalty` (accepted by the reference's `sample` strategy, eval.py:279) run on the host loop as well: the
penalty is a processor on the raw step scores, typical filtering sits between the top-p filter and the renormalisation (HF 4.25 order),
both as a few torch ops on the device around the sampling kernel.  Beam search and contrastive search are outside the path and raise.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import decode, ops


def repetition_penalty_scores(scores, prev_ids, penalty: float):
    """HF RepetitionPenaltyLogitsProcessor: scores of tokens already in `prev_ids` are divided by `penalty` when positive, multiplied when negative."""
    s = torch.gather(scores, 1, prev_ids)
    s = torch.where(s < 0, s * penalty, s / penalty)
    return scores.scatter(1, prev_ids, s)


def typical_filter_scores(scores, mass: float, min_tokens_to_keep: int = 1):
    """HF TypicalLogitsWarper: keep the tokens whose surprisal is nearest the entropy until their mass reaches `mass`; -inf elsewhere."""
    logp = torch.log_softmax(scores, dim=-1)
    p = logp.exp()
    ent = -(logp * p).nansum(-1, keepdim=True)
    dist = torch.abs(-logp - ent)
    sorted_dist, order = torch.sort(dist, descending=False)
    cum = scores.gather(-1, order).softmax(dim=-1).cumsum(dim=-1)
    last = (cum < mass).sum(dim=1).clamp_(max=scores.shape[-1] - 1)
    remove_sorted = sorted_dist > sorted_dist.gather(1, last.view(-1, 1))
    remove_sorted[..., :min_tokens_to_keep] = False
    remove = remove_sorted.scatter(1, order, remove_sorted)
    return scores.masked_fill(remove, -float('inf'))


def generate(model, input_ids=None, max_length: Optional[int] = None, max_new_tokens: Optional[int] = None, do_sample: Optional[bool] = None,
             early_stopping=None, top_k: Optional[int] = None, top_p: Optional[float] = None, temperature: Optional[float] = None,
             typical_p: Optional[float] = None, repetition_penalty: Optional[float] = None, renormalize_logits: Optional[bool] = None,
             num_beams: Optional[int] = None, num_beam_groups: Optional[int] = None, diversity_penalty=None, penalty_alpha=None,
             num_return_sequences: Optional[int] = None, eos_token_id='config', pad_token_id='config', generator=None,
             return_step_scores: bool = False, use_decode_cache: bool = True, seed: Optional[int] = None, seq_offset: int = 0, use_cuda_graph: bool = True,
             decode_groups: Optional[int] = None, **unused):
    cfg = model.config
    if input_ids is None:
        raise ValueError('generate needs input_ids (the reference always passes a tokenised prompt, eval.py:276)')
    if (num_beams or 1) > 1 or (num_beam_groups or 1) > 1 or penalty_alpha:
        raise NotImplementedError('beam / diverse-beam / contrastive search are not on the measured path (SURVEY §8b)')
    typical = typical_p is not None and typical_p < 1.0
    if typical and not (0.0 < typical_p):
        raise ValueError(f'typical_p has to be in (0, 1), got {typical_p}')
    penalise = repetition_penalty is not None and repetition_penalty != 1.0
    if penalise and not repetition_penalty > 0:
        raise ValueError(f'repetition_penalty has to be > 0, got {repetition_penalty}')
    do_sample = bool(cfg.do_sample if do_sample is None else do_sample)
    top_k = cfg.top_k if top_k is None else top_k             # HF: config default top_k=50 applies when the caller passes none
    top_p = cfg.top_p if top_p is None else top_p
    temperature = cfg.temperature if temperature is None else temperature
    nret = num_return_sequences or 1
    if eos_token_id == 'config':
        eos_token_id = cfg.eos_token_id
    if pad_token_id == 'config':
        pad_token_id = cfg.pad_token_id
    if pad_token_id is None and eos_token_id is not None:
        pad_token_id = eos_token_id                            # HF: "Setting pad_token_id to eos_token_id"
    model._ensure_engine()
    dev = model._flat.device
    ids = input_ids.to(dev).long()
    if nret > 1:
        ids = ids.repeat_interleave(nret, dim=0)
    if max_length is None:
        max_length = ids.shape[1] + max_new_tokens if max_new_tokens is not None else cfg.max_length
    B = ids.shape[0]
    was_training = model.training
    model.eval()
    unfinished = torch.ones(B, dtype=torch.int64, device=dev)
    out_ids = torch.empty(B, max_length, dtype=torch.int64, device=dev)
    cur = ids.shape[1]
    out_ids[:, :cur] = ids
    step_scores = []
    past = None
    fast = (use_decode_cache and not return_step_scores and generator is None and not typical and not penalise and decode.supported(model, B)
            and max_length - cur >= 2)
    if fast and do_sample and seed is None:
        # the reference's call carries no seed (eval.py:277-333): key the draws on torch's global generator, as HF's multinomial is
        seed = int(torch.randint(1, 2 ** 62, (1,)).item())
    seed = int(seed or 0)
    model.last_generate_path = 'decode_cache' if fast else 'forward_per_token'
    try:
        with torch.no_grad():
            while cur < max_length:
                inputs = model.prepare_inputs_for_generation(out_ids[:, :cur], past=past)
                out = model(**inputs, return_dict=True)
                scores = out.logits[:, -1, :].contiguous()
                past = out.mems
                if fast:
                    u = None
                    if do_sample:      # first draw uses the same keyed stream as the device loop (step index -1 -> pos 0 is the next one)
                        u = torch.empty(B, dtype=torch.float32, device=dev)
                        from ._lib import check as _chk, load as _ld, ptr as _ptr, stream_ptr as _sp
                        neg = torch.full((1,), -1, dtype=torch.int32, device=dev)
                        _chk(_ld().txl_decode_uniform(_ptr(u), B, int(seed), int(seq_offset), _ptr(neg), _sp()), 'decode_uniform')
                    nxt, _, _ = ops.sample(scores, do_sample, temperature, top_k, top_p, u)
                    if eos_token_id is not None:
                        nxt = nxt * unfinished + pad_token_id * (1 - unfinished)
                        unfinished = unfinished * (nxt != eos_token_id).long()
                    out_ids[:, cur] = nxt
                    cur += 1
                    if cur < max_length:
                        dkw = dict(do_sample=do_sample, temperature=temperature, top_k=top_k, top_p=top_p, eos_token_id=eos_token_id,
                                   pad_token_id=pad_token_id, seed=seed, seq_offset=seq_offset, use_graph=use_cuda_graph)
                        dec = decode.make_decoder(model, past, out_ids, cur, groups=decode_groups, **dkw)
                        dec.set_unfinished(unfinished)
                        cur += dec.run(nxt, max_length - cur)
                    break
                u = torch.rand(B, device=dev, generator=generator) if do_sample else None
                if penalise:         # processor on the raw scores (greedy search applies it too); HF renormalises after it when asked to
                    scores = repetition_penalty_scores(scores, out_ids[:, :cur], float(repetition_penalty))
                    if not do_sample and renormalize_logits:
                        scores = torch.log_softmax(scores, dim=-1)
                if do_sample and typical:
                    # temperature / top-k / top-p in the kernel (its draw is discarded), typical filter on the warped scores, second call samples
                    _, _, warped = ops.sample(scores, True, temperature, top_k, top_p, u, want_warped=True)
                    nxt, _, warped = ops.sample(typical_filter_scores(warped, float(typical_p)), True, 1.0, 0, 1.0, u, want_warped=return_step_scores)
                else:
                    nxt, _, warped = ops.sample(scores, do_sample, temperature, top_k, top_p, u, want_warped=return_step_scores and do_sample)
                if return_step_scores:
                    step_scores.append(warped if do_sample else scores)
                if eos_token_id is not None:
                    nxt = nxt * unfinished + pad_token_id * (1 - unfinished)
                    unfinished = unfinished * (nxt != eos_token_id).long()
                out_ids[:, cur] = nxt
                cur += 1
                if eos_token_id is not None and (cur % 64 == 0 or cur == max_length) and int(unfinished.max().item()) == 0:
                    break
    finally:
        if was_training:
            model.train()
    result = out_ids[:, :cur]
    if eos_token_id is not None and cur > ids.shape[1]:
        # HF stops right after the step at which every sequence has emitted eos; we only poll every 64 steps, so trim.
        gen = result[:, ids.shape[1]:]
        is_eos = gen == eos_token_id
        if bool(is_eos.any(dim=1).all().item()):
            first = torch.where(is_eos, torch.arange(gen.shape[1], device=dev).expand_as(gen), gen.shape[1]).min(dim=1).values
            result = result[:, :ids.shape[1] + int(first.max().item()) + 1]
    return (result, step_scores) if return_step_scores else result
