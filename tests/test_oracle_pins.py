"""Independent pins of the oracle (CPU).

The reference holds no golden vector for this path and HF `transformers==4.25.1` (the file the arithmetic lives in) cannot be installed here,
so `oracle/txl_ref.py` cannot be pinned against the reference itself ("parity unpinned").  What CAN be done here is to check the
restatement against third-party implementations of the same published algorithms that ARE installed, none of which shares code with the oracle:

  * `transformers.models.xlnet.modeling_xlnet.XLNetLayer` (transformers 5.5) — XLNet's relative attention is Transformer-XL's, by the same
    authors: q/k/v/r/o projections, `r_w_bias` / `r_r_bias`, AC + BD, a pad-free formulation of the relative shift (`rel_shift_bnij`),
    1/sqrt(d_head), softmax over keys, P.V, o projection, residual + post-LayerNorm, then the position-wise FF with residual + post-LayerNorm.
    Fed the oracle's weights, TXL's band mask and TXL's (clamped) position sequence, it must reproduce the oracle's decoder layer.
  * `torch.nn.AdaptiveLogSoftmaxWithLoss` — the two-level adaptive softmax (Grave et al.) HF's `ProjectedAdaptiveLogSoftmax` implements.
  * `transformers.generation.logits_process` (5.5) — temperature / top-k / top-p / renormalisation warpers of `generate`.
  * `transformers.get_cosine_schedule_with_warmup`, `Trainer.get_decay_parameter_names` — the optimiser-side rules of §8f-1.
"""
import math

import pytest
import torch

from oracle.txl_ref import (RefConfig, RefTransfoXLLMHeadModel, _Crit, hf_param_groups, literal_attn_mask)


def _xlnet_layer_from(ref_layer, cfg):
    from transformers.models.xlnet.configuration_xlnet import XLNetConfig
    from transformers.models.xlnet.modeling_xlnet import XLNetLayer
    d, H, dh = cfg.d_model, cfg.n_head, cfg.d_head
    xc = XLNetConfig(vocab_size=8, d_model=d, n_layer=1, n_head=H, d_inner=cfg.d_inner, ff_activation='relu', untie_r=True, attn_type='uni',
                     layer_norm_eps=cfg.layer_norm_epsilon, dropout=0.0)
    assert xc.d_head == dh
    xl = XLNetLayer(xc).eval()
    att = ref_layer.dec_attn
    wq, wk, wv = torch.chunk(att.qkv_net.weight.detach(), 3, dim=0)           # (H*dh, d) each
    with torch.no_grad():
        for name, w in (('q', wq), ('k', wk), ('v', wv), ('r', att.r_net.weight.detach())):
            getattr(xl.rel_attn, name).copy_(w.t().reshape(d, H, dh))         # einsum('ibh,hnd->ibnd'): [h_in, n, d] = W[n*dh+d, h_in]
        xl.rel_attn.o.copy_(att.o_net.weight.detach().reshape(d, H, dh))      # einsum('ibnd,hnd->ibh'): [h_out, n, d] = W_o[h_out, n*dh+d]
        xl.rel_attn.r_w_bias.copy_(att.r_w_bias)
        xl.rel_attn.r_r_bias.copy_(att.r_r_bias)
        xl.rel_attn.r_s_bias.zero_()
        xl.rel_attn.seg_embed.zero_()
        xl.rel_attn.layer_norm.load_state_dict(att.layer_norm.state_dict())
        ff = ref_layer.pos_ff
        xl.ff.layer_1.load_state_dict(ff.CoreNet[0].state_dict())
        xl.ff.layer_2.load_state_dict(ff.CoreNet[3].state_dict())
        xl.ff.layer_norm.load_state_dict(ff.layer_norm.state_dict())
    return xl


@pytest.mark.parametrize('T,mlen,mem_len,clamp', [(7, 5, 5, 3), (6, 6, 6, 100), (1, 8, 8, 4), (5, 0, 4, 2), (4, 2, 6, 3), (16, 16, 16, 8)])
def test_decoder_layer_equals_hf_xlnet_layer(T, mlen, mem_len, clamp):
    torch.manual_seed(T * 100 + mlen)
    cfg = RefConfig(vocab_size=11, d_model=32, n_head=4, n_layer=1, d_head=8, d_inner=48, d_embed=32, mem_len=mem_len, clamp_len=clamp, dropout=0.0)
    ref = RefTransfoXLLMHeadModel(cfg).eval()
    layer = ref.transformer.layers[0]
    with torch.no_grad():            # non-trivial LayerNorm / bias values
        for p in layer.parameters():
            p.add_(0.05 * torch.randn_like(p))
    B, d = 3, cfg.d_model
    klen = T + mlen
    w = torch.randn(T, B, d)
    mems = torch.randn(mlen, B, d) if mlen > 0 else None
    mask = literal_attn_mask(T, mlen, mem_len, True)
    # --- oracle: TXL position sequence klen-1 .. 0 (clamped), literal pad/view _rel_shift
    pos_seq = torch.arange(klen - 1, -1, -1.0).clamp(max=clamp)
    r_txl = ref.transformer.pos_emb(pos_seq)                                   # (klen, 1, d)
    want = layer(w, r_txl, mask, mems)
    # --- XLNet: position sequence klen .. -T+1 (its rel_shift_bnij drops the first row instead of padding a zero column)
    pos_xl = torch.arange(klen, -T, -1.0).clamp(max=clamp)
    r_xl = ref.transformer.pos_emb(pos_xl).expand(-1, B, -1)                   # (klen+T, B, d)
    xl = _xlnet_layer_from(layer, cfg)
    got = xl(w, None, mask[:, :, None, None].float(), None, r_xl, None, mems=mems)[0]
    assert torch.allclose(got, want, rtol=1e-5, atol=2e-6), (got - want).abs().max()


def test_adaptive_softmax_equals_torch_adaptive_logsoftmax():
    torch.manual_seed(5)
    d, V, cutoffs = 24, 57, [20, 41]
    crit = _Crit(V, d, cutoffs).double()
    with torch.no_grad():
        crit.out_layers[0].weight.normal_(0, 0.3)
        crit.out_layers[0].bias.normal_(0, 0.3)
        crit.out_layers[0].bias[cutoffs[0]:] = 0          # torch's tails have no bias; the shortlist and cluster biases are exercised
        crit.cluster_weight.normal_(0, 0.3)
        crit.cluster_bias.normal_(0, 0.3)
    tr = torch.nn.AdaptiveLogSoftmaxWithLoss(d, V, cutoffs, div_value=1.0, head_bias=True).double()
    with torch.no_grad():
        tr.head.weight.copy_(torch.cat([crit.out_layers[0].weight[:cutoffs[0]], crit.cluster_weight], 0))
        tr.head.bias.copy_(torch.cat([crit.out_layers[0].bias[:cutoffs[0]], crit.cluster_bias], 0))
        ends = cutoffs + [V]
        for i in range(2):
            tr.tail[i][0].weight.copy_(torch.eye(d, dtype=torch.float64))
            tr.tail[i][1].weight.copy_(crit.out_layers[0].weight[ends[i]:ends[i + 1]])
    hidden = torch.randn(2, 9, d, dtype=torch.float64)
    full = crit(hidden, None)                                                        # (18, V) log-probs
    assert torch.allclose(full, tr.log_prob(hidden.view(-1, d)), atol=1e-10)
    assert torch.allclose(full.exp().sum(-1), torch.ones(18, dtype=torch.float64), atol=1e-10)
    labels = torch.randint(0, V, (2, 9))
    nll_pos = crit(hidden, labels, keep_order=True)                                  # (16,) in position order
    want = -tr(hidden[:, :-1].reshape(-1, d), labels[:, 1:].reshape(-1)).output
    assert torch.allclose(nll_pos, want, atol=1e-10)
    # keep_order=False (what TransfoXLLMHeadModel uses): the same values packed cluster by cluster, ignored labels as trailing zeros
    labels[0, 3] = -100
    lab = labels[:, 1:].reshape(-1)
    nll_pos = crit(hidden, labels, keep_order=True)
    packed = crit(hidden, labels)
    want = torch.cat([nll_pos[(lab >= a) & (lab < b)] for a, b in zip([0] + cutoffs, cutoffs + [V])])
    assert torch.equal(packed[:want.numel()], want) and torch.equal(packed[want.numel():], torch.zeros(1, dtype=torch.float64))
    assert nll_pos[2] == 0          # the ignored position


def test_warpers_equal_hf_logits_processors():
    from transformers.generation.logits_process import LogitNormalization, TemperatureLogitsWarper, TopKLogitsWarper, TopPLogitsWarper
    torch.manual_seed(9)
    scores = torch.log_softmax(2.0 * torch.randn(5, 97), -1)
    ids = torch.zeros(5, 1, dtype=torch.long)
    for temperature, top_k, top_p in [(1.0, 8, 1.0), (0.9, 32, 0.9), (1.3, 0, 0.75), (1.0, 50, 0.5), (0.7, 1, 1.0)]:
        want = scores.clone()
        if temperature != 1.0:
            want = TemperatureLogitsWarper(temperature)(ids, want)
        if top_k:
            want = TopKLogitsWarper(top_k)(ids, want)
        if top_p < 1.0:
            want = TopPLogitsWarper(top_p)(ids, want)
        want = LogitNormalization()(ids, want)
        got = RefTransfoXLLMHeadModel.warp_scores(scores, temperature, top_k, top_p, True)
        assert torch.equal(torch.isinf(got), torch.isinf(want)), (temperature, top_k, top_p)
        keep = ~torch.isinf(want)
        assert torch.allclose(got[keep], want[keep], atol=1e-6)


def test_repetition_penalty_and_typical_equal_hf_processors():
    """`repetition_penalty` / `typical_p` (accepted keys of the reference's `sample` strategy, musicnlp/trainer/eval.py:279): the oracle's
    row-by-row restatements and the product's batched torch versions (generation.py, host-loop path) == HF's RepetitionPenaltyLogitsProcessor /
    TypicalLogitsWarper on the same scores."""
    import importlib
    from transformers.generation.logits_process import RepetitionPenaltyLogitsProcessor, TypicalLogitsWarper
    gen = importlib.import_module('symbolic-music-generation_b200.generation')
    torch.manual_seed(11)
    scores = torch.log_softmax(2.5 * torch.randn(6, 83), -1)
    scores[:, :5] += 3.0                      # some positive scores: the penalty divides those and multiplies negative ones
    prev = torch.randint(0, 83, (6, 19))
    prev[:, 3] = prev[:, 7]                   # repeated tokens are penalised once
    for pen in (1.2, 0.8, 2.0):
        want = RepetitionPenaltyLogitsProcessor(pen)(prev, scores.clone())
        assert torch.allclose(RefTransfoXLLMHeadModel.repetition_penalty(scores, prev, pen), want, atol=1e-6)
        assert torch.allclose(gen.repetition_penalty_scores(scores, prev, pen), want, atol=1e-6)
    filt = scores.clone()
    filt[:, 40:] = -float('inf')              # as it arrives from the top-k / top-p filters
    for mass in (0.2, 0.5, 0.9, 0.95):
        for sc in (scores, filt):
            want = TypicalLogitsWarper(mass)(prev, sc.clone())
            for got in (RefTransfoXLLMHeadModel.typical_filter(sc, mass), gen.typical_filter_scores(sc, mass)):
                assert torch.equal(torch.isinf(got), torch.isinf(want)), mass
                keep = ~torch.isinf(want)
                assert torch.equal(got[keep], want[keep]) and int(keep.sum(1).min()) >= 1


def test_schedule_and_decay_groups_equal_hf():
    import importlib
    import transformers
    optim = importlib.import_module('symbolic-music-generation_b200.optim')
    p = torch.nn.Parameter(torch.zeros(1))
    for total, warm in [(6, 2), (50, 5), (10, 0)]:
        opt = torch.optim.SGD([p], lr=3e-4)
        sch = transformers.get_cosine_schedule_with_warmup(opt, warm, total)
        for step in range(total):
            assert abs(opt.param_groups[0]['lr'] - optim.cosine_with_warmup(step, total, warm, 3e-4)) < 1e-12
            opt.step()
            sch.step()
    cfg = RefConfig(vocab_size=1190, d_model=32, n_head=4, n_layer=2, d_head=8, d_inner=48, d_embed=32, mem_len=4, clamp_len=8, cutoffs=[1000])
    model = RefTransfoXLLMHeadModel(cfg)
    names = set(transformers.Trainer.get_decay_parameter_names(None, model))
    decay = {id(q) for q in hf_param_groups(model, 0.01)[0]['params']}
    by_name = dict(model.named_parameters(remove_duplicate=False))
    assert {id(by_name[n]) for n in names} == decay
    assert 'crit.cluster_weight' in names and not any('bias' in n or 'layer_norm' in n for n in names)


def test_param_count_kats():
    from oracle.txl_ref import expected_param_count
    assert expected_param_count(12, 768, 3072, 418) == 92_435_362                      # notebook/train/transformer-xl.ipynb:491 ("92.4M")
    cfg = RefConfig(vocab_size=1190, d_model=64, n_head=4, n_layer=2, d_head=16, d_inner=128, d_embed=64, mem_len=8, clamp_len=8, cutoffs=[1000])
    m = RefTransfoXLLMHeadModel(cfg)
    assert m.num_parameters() == expected_param_count(2, 64, 128, 1190, 1)
    assert {'crit.cluster_weight', 'crit.cluster_bias', 'crit.out_layers.0.bias'} <= set(m.state_dict())
