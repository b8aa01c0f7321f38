"""CPU tests of the oracle (oracle/txl_ref.py): known-answer facts the reference logs, closed forms vs the literal HF
constructions, and the committed golden fixture (regression pin of the oracle itself — parity with HF 4.25.1 is unpinned)."""
import os

import pytest
import torch

from oracle.txl_ref import (RefConfig, RefTransfoXLLMHeadModel, expected_param_count, literal_attn_mask, literal_index_maps,
                            literal_rel_shift)

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'oracle_tiny.pt')


def test_param_count_kat():
    # notebook/train/transformer-xl.ipynb:491 logs "92.4M" for 12 L / d 768 / di 3072 / V 418
    assert expected_param_count(12, 768, 3072, 418) == 92_435_362
    cfg = RefConfig(vocab_size=37, d_model=32, n_head=4, n_layer=3, d_head=8, d_inner=48, d_embed=32, mem_len=4, clamp_len=8)
    assert RefTransfoXLLMHeadModel(cfg).num_parameters() == expected_param_count(3, 32, 48, 37)


def test_config_preset_small():
    # reference transformer_xl.py:20,26-49 with train.py:521-527 overrides
    c = RefConfig.from_preset('small', vocab_size=1190, max_length=1024, mem_len=1024, cutoffs=[])
    assert (c.d_model, c.n_head, c.n_layer, c.d_head, c.d_inner, c.mem_len, c.clamp_len) == (512, 8, 12, 64, 2048, 1024, 1024)
    c = RefConfig.from_preset('small', vocab_size=1190)
    assert c.mem_len == 128 and c.cutoffs == [1000]


def test_state_dict_names():
    cfg = RefConfig(vocab_size=37, d_model=32, n_head=4, n_layer=1, d_head=8, d_inner=48, d_embed=32, mem_len=4, clamp_len=8)
    keys = set(RefTransfoXLLMHeadModel(cfg).state_dict())
    want = {'transformer.word_emb.emb_layers.0.weight', 'transformer.pos_emb.inv_freq', 'crit.out_layers.0.weight', 'crit.out_layers.0.bias'}
    a, f = 'transformer.layers.0.dec_attn.', 'transformer.layers.0.pos_ff.'
    want |= {a + n for n in ('qkv_net.weight', 'o_net.weight', 'r_net.weight', 'r_r_bias', 'r_w_bias', 'layer_norm.weight', 'layer_norm.bias')}
    want |= {f + n for n in ('CoreNet.0.weight', 'CoreNet.0.bias', 'CoreNet.3.weight', 'CoreNet.3.bias', 'layer_norm.weight', 'layer_norm.bias')}
    assert keys == want


GRID = [(7, 5, 5, 3), (6, 6, 6, 100), (1, 8, 8, 4), (5, 0, 4, 2), (4, 2, 6, 3), (9, 4, 4, 6), (3, 9, 4, 0), (16, 16, 16, 8), (1, 1, 1, 0),
        (33, 31, 31, 1024), (12, 0, 12, 5)]


@pytest.mark.parametrize('T,M,ML,C', GRID)
def test_closed_forms_match_literal(T, M, ML, C):
    """live set and BD row index in closed form == pad/reshape + triu/tril construction."""
    mask, ridx = literal_index_maps(T, M, ML, C)
    klen = M + T
    mask_len = klen - ML
    msl = T - mask_len if mask_len > 0 else T
    for i in range(T):
        for j in range(klen):
            masked = (j > i + M) or (j <= i - msl)
            assert bool(mask[i, j]) == masked
            if not masked:
                p = M + i - j
                assert int(ridx[i, j]) == (min(p, C) if C > 0 else p)


def test_rel_shift_example():
    x = torch.arange(12.).view(3, 4)
    y = literal_rel_shift(x)
    # row i keeps BD0[i, j + T-1-i] for j <= mlen + i (mlen = klen - T = 1)
    assert y[0, :2].tolist() == [2., 3.] and y[1, :3].tolist() == [5., 6., 7.] and y[2].tolist() == [8., 9., 10., 11.]


def test_live_keys_per_row_same_length():
    m = literal_attn_mask(8, 8, 8, True)
    assert ((m == 0).sum(1) == 8).all()      # exactly mem_len live keys per query when mlen == mem_len
    m = literal_attn_mask(5, 0, 4, True)
    assert (m == 0).sum(1).tolist() == [1, 2, 3, 4, 4]


def test_zero_mems_take_softmax_mass():
    """mems=None means zero mems that ARE attended (Appendix A.5): output differs from a mem_len-free run."""
    torch.manual_seed(0)
    cfg = RefConfig(vocab_size=37, d_model=32, n_head=4, n_layer=1, d_head=8, d_inner=48, d_embed=32, mem_len=6, clamp_len=8, dropout=0.0)
    m = RefTransfoXLLMHeadModel(cfg).eval()
    ids = torch.randint(0, 37, (2, 5))
    with torch.no_grad():
        a = m(input_ids=ids).logits
        b = m(input_ids=ids, mems=[torch.zeros(6, 2, 32)]).logits
    assert torch.equal(a, b)


def test_golden_fixture():
    g = torch.load(GOLD)
    m = RefTransfoXLLMHeadModel(RefConfig(**g['cfg'])).eval()
    m.load_state_dict(g['state_dict'])
    with torch.no_grad():
        o1 = m(input_ids=g['ids'], labels=g['labels'].clone())
        o2 = m(input_ids=g['ids'][:, :4], mems=o1.mems)
    torch.testing.assert_close(o1.losses, g['losses'], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(o1.logits, g['logits'], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(o2.logits, g['logits2'], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(o1.mems[0], g['mems0'], rtol=1e-5, atol=1e-6)
    assert torch.equal(m.generate(g['ids'][:, :3], max_length=14), g['greedy'])
    for key, (mask, ridx) in g['maps'].items():
        T, M, ML, C = map(int, key.split(','))
        mk, rx = literal_index_maps(T, M, ML, C)
        assert torch.equal(mk, mask) and torch.equal(rx, ridx)


def test_loss_ignores_pad_and_fixup():
    torch.manual_seed(1)
    cfg = RefConfig(vocab_size=37, d_model=32, n_head=4, n_layer=1, d_head=8, d_inner=48, d_embed=32, mem_len=4, clamp_len=8, dropout=0.0)
    m = RefTransfoXLLMHeadModel(cfg).train()
    ids = torch.randint(1, 37, (2, 6))
    labels = ids.clone()
    labels[0, 1:] = -100
    o = m(input_ids=ids, labels=labels)
    assert labels[0, 1] == cfg.eos_token_id          # reference transformer_xl.py:176-182 mutates the caller's labels
    assert o.prediction_scores == () and o.losses.shape == (2, 5)
    nz = o.losses[o.losses != 0]
    torch.testing.assert_close(o.loss, nz.mean())


def test_warpers_keep_sets():
    s = torch.log_softmax(torch.tensor([[2.0, 1.0, 1.0, 0.5, -1.0, -3.0]]), -1)
    w = RefTransfoXLLMHeadModel.warp_scores(s, top_k=2)
    assert (w > -float('inf')).sum() == 3            # ties at the k-th value are kept
    w = RefTransfoXLLMHeadModel.warp_scores(s, top_k=0, top_p=0.5)
    assert (w > -float('inf')).sum() == 2            # first token crossing p is kept
    torch.testing.assert_close(w.exp().sum(), torch.tensor(1.0))


def test_ntp_acc_counts_known_answer():
    """KAT by hand for the reference's shifted, pad-masked accuracy (train_util_wrap.py:113-120)."""
    from oracle.txl_ref import ntp_acc_counts
    preds = torch.tensor([[5, 6, 7, 0], [1, 1, 1, 1]])
    labels = torch.tensor([[9, 5, 9, 7], [1, 1, -100, -100]])
    # row 0: preds[:3] = 5,6,7 vs labels[1:] = 5,9,7 -> 2 of 3; row 1: preds[:3] = 1,1,1 vs 1,-100,-100 -> 1 of 1
    assert ntp_acc_counts(preds, labels) == (3, 4)
    assert ntp_acc_counts(preds[:, :1], labels[:, :1]) == (0, 0)


def test_clm_collate_and_truncate_known_answers():
    """KATs by hand for the two format steps either side of the path (train.py:360, eval.py:178-185)."""
    from oracle.txl_ref import clm_collate, truncate_last_bar
    ids = torch.tensor([[7, 3, 1, 1], [1, 5, 6, 2]])
    out = clm_collate(ids, pad_token_id=1)
    assert out['labels'].tolist() == [[7, 3, -100, -100], [-100, 5, 6, 2]] and out['input_ids'] is ids
    assert truncate_last_bar(torch.tensor([4, 9, 5, 9, 6]), 9) == [4, 9, 5]
    assert truncate_last_bar(torch.tensor([9, 5]), 9) == []
    with pytest.raises(AssertionError):
        truncate_last_bar(torch.tensor([4, 5]), 9)


def test_oracle_train_steps_schedule_and_groups():
    """HF rules restated in the oracle's training loop: ceil warm-up, lambda(0) = 0 at the first step, no decay on biases / LayerNorm."""
    from oracle.txl_ref import hf_param_groups, train_steps
    torch.manual_seed(0)
    m = RefTransfoXLLMHeadModel(RefConfig(vocab_size=50, d_model=32, d_embed=32, n_head=2, d_head=16, d_inner=64, n_layer=1, mem_len=8, dropout=0.0))
    groups = hf_param_groups(m, 0.01)
    n_decay, n_nodecay = sum(p.numel() for p in groups[0]['params']), sum(p.numel() for p in groups[1]['params'])
    assert n_decay + n_nodecay == sum(p.numel() for p in m.parameters())
    names_nd = {n for n, p in m.named_parameters() if any(p is q for q in groups[1]['params'])}
    assert 'transformer.layers.0.dec_attn.r_w_bias' in names_nd and 'transformer.layers.0.pos_ff.layer_norm.weight' in names_nd
    assert 'transformer.layers.0.dec_attn.qkv_net.weight' not in names_nd
    g = torch.Generator().manual_seed(1)
    batches = [(torch.randint(0, 50, (2, 12), generator=g),) * 2 for _ in range(5)]
    logs = train_steps(m, batches, total_steps=5, warmup_ratio=0.3)       # ceil(1.5) = 2 warm-up steps
    lrs = [l['learning_rate'] for l in logs]
    assert lrs[0] == 0.0 and abs(lrs[1] - 1.5e-4) < 1e-12 and abs(lrs[2] - 3e-4) < 1e-12 and lrs[3] < lrs[2] and lrs[4] < lrs[3]
    assert logs[-1]['loss'] < logs[1]['loss']


@pytest.mark.parametrize('ML,clamp', [(8, 1024), (8, 3), (33, 16), (128, 1024), (5, 5)])
def test_decode_ring_index_map_equals_literal_shift(ML, clamp):
    """Integer indexing of the T=1 decode step, bit-exact on the CPU: the ring-slot -> r-row closed form of csrc/decode_stream.cu / decode.cu
    (slot s <= cur: ML - cur + s, else s - cur; the bulk-copy split of a stage at the wrap point) addresses exactly the relative positions that
    HF's pad/reshape `_rel_shift` + same_length mask give for qlen=1, mlen=mem_len=ML, and key 0 (the slot being overwritten) is the masked one."""
    from oracle.txl_ref import literal_index_maps, literal_pos_seq
    masked, ridx = literal_index_maps(1, ML, ML, clamp, True)
    assert masked[0].tolist() == [1] + [0] * ML                   # cat(mems, cur): only the oldest memory is outside the band
    table_pos = literal_pos_seq(ML + 1, clamp)                    # relative-position value of r row x (what posemb_table(ML+1) encodes)
    for pos in list(range(0, 2 * ML + 3)):
        cur = pos % ML
        xs = []
        for s in range(ML):
            x = ML - cur + s if s <= cur else s - cur            # the kernels' closed form
            dist = (cur - s) % ML
            assert x == ML - dist
            xs.append(x)
            # ring slot s holds the token `dist` steps back = key j = ML - dist of cat(mems, cur) (slot cur = the new token = key ML)
            j = ML - dist
            assert masked[0, j] == 0 and int(ridx[0, j]) == int(table_pos[x])
        assert sorted(xs) == list(range(1, ML + 1)) and xs[cur] == ML
        # the producer's two bulk copies per stage reproduce the same rows (stage sizes of every geometry the kernel has)
        for CKS in (32, 64, 128):
            for s0 in range(0, ML, CKS):
                n = min(CKS, ML - s0)
                n1 = max(0, min(n, cur + 1 - s0))
                rows = [ML - cur + s0 + i for i in range(n1)] + [s0 + n1 - cur + i for i in range(n - n1)]
                assert rows == xs[s0:s0 + n]
