"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, and exports exactly what include/txl_b200.h declares;
the Python binding table matches it; and the product fails loudly without a GPU (no CPU / oracle fallback)."""
import ctypes
import importlib
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = 'symbolic-music-generation_b200'


def _declared():
    txt = open(os.path.join(ROOT, 'include', 'txl_b200.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return set(re.findall(r'\b(txl_[a-z0-9_]+)\s*\(', txt))


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/txl_b200.h but not exported'


def test_binding_table_matches_header(built_lib):
    L = importlib.import_module(PKG + '._lib')
    assert set(L.SIGNATURES) == _declared()
    lib = L.load()
    assert lib.txl_version() >= 100


def test_sass_is_sm100a(built_lib):
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', built_lib], capture_output=True, text=True).stdout
    assert 'sm_100a' in out


def test_product_does_not_import_oracle():
    pkgdir = os.path.join(ROOT, PKG)
    for fn in os.listdir(pkgdir):
        if fn.endswith('.py'):
            src = open(os.path.join(pkgdir, fn)).read()
            assert 'import oracle' not in src and 'from oracle' not in src, fn


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_fails_loudly_without_gpu(pkg, built_lib):
    cfg = pkg.MyTransfoXLConfig('debug', vocab_size=50, cutoffs=[])
    m = pkg.MyTransfoXLLMHeadModel(cfg)
    with pytest.raises(pkg.TxlError):
        m(input_ids=torch.zeros(1, 4, dtype=torch.long))


def test_config_mirrors_reference_presets(pkg):
    c = pkg.MyTransfoXLConfig('small', vocab_size=1190, max_length=1024, mem_len=1024, cutoffs=[])
    assert (c.d_model, c.n_head, c.n_layer, c.d_head, c.d_inner, c.mem_len, c.clamp_len, c.div_val) == (512, 8, 12, 64, 2048, 1024, 1024, 1)
    assert c.same_length and c.untie_r and not c.pre_lnorm and c.eos_token_id == 0 and c.dropout == 0.1
    assert c.max_length_ == 1024 and c.model_meta['seg_len'] == 1024

    class Tok:
        vocab_size = 1190
    assert pkg.MyTransfoXLConfig('small', tokenizer=Tok()).cutoffs == [1000]
    Tok.vocab_size = 422
    assert pkg.MyTransfoXLConfig('small', tokenizer=Tok()).cutoffs == []
    # the reference's default criterion for V >= 1000: adaptive-softmax cluster path, HF state_dict names, cluster rows glued behind the embedding
    m = pkg.MyTransfoXLLMHeadModel(pkg.MyTransfoXLConfig('debug', vocab_size=1190, cutoffs=[1000]))
    sd = m.state_dict()
    assert sd['crit.cluster_weight'].shape == (1, 128) and sd['crit.cluster_bias'].shape == (1,)
    emb, cw = m.transformer.word_emb.emb_layers[0].weight, m.crit.cluster_weight
    assert cw.data_ptr() == emb.data_ptr() + emb.numel() * 4 and m.crit.cluster_bias.data_ptr() == m.crit.out_layers[0].bias.data_ptr() + 1190 * 4
    assert float(m.crit.cluster_bias.abs().sum()) == 0 and float(cw.abs().sum()) > 0
    with pytest.raises(NotImplementedError):
        pkg.MyTransfoXLLMHeadModel(pkg.MyTransfoXLConfig('debug', vocab_size=1190, cutoffs=[1000, 500]))


@pytest.mark.parametrize('cutoffs', [[], [20], [10, 25]])
def test_state_dict_roundtrip_with_oracle(pkg, tmp_path, cutoffs):
    from oracle.txl_ref import RefConfig, RefTransfoXLLMHeadModel
    kw = dict(vocab_size=37, d_model=32, n_head=4, n_layer=2, d_head=8, d_inner=48, mem_len=4, clamp_len=8)
    ref = RefTransfoXLLMHeadModel(RefConfig(d_embed=32, cutoffs=cutoffs, **kw))
    m = pkg.MyTransfoXLLMHeadModel(pkg.MyTransfoXLConfig('debug', cutoffs=cutoffs, d_embed=32, **kw))
    assert set(m.state_dict()) == set(ref.state_dict())
    assert m.num_parameters() == ref.num_parameters()
    m.load_state_dict(ref.state_dict())
    for k, v in ref.state_dict().items():
        assert torch.equal(m.state_dict()[k], v), k
    m.save_pretrained(tmp_path)
    m2 = pkg.MyTransfoXLLMHeadModel.from_pretrained(tmp_path)
    for k, v in m.state_dict().items():
        assert torch.equal(m2.state_dict()[k], v), k
    assert m2.config.mem_len == 4 and m2.config.max_length_ == m.config.max_length_


def test_output_container(pkg):
    o = pkg.TransfoXLLMHeadModelOutput(loss=torch.tensor(1.0), prediction_scores=(), losses=torch.zeros(2, 3), mems=[1])
    assert o['loss'] == 1.0 and o.logits == () and list(o.keys()) == ['losses', 'prediction_scores', 'mems', 'loss']
    assert o[0] is o.losses


def test_prepare_inputs_for_generation(pkg):
    m = pkg.MyTransfoXLLMHeadModel(pkg.MyTransfoXLConfig('debug', vocab_size=50, cutoffs=[]))
    ids = torch.arange(10).view(2, 5)
    assert m.prepare_inputs_for_generation(ids)['input_ids'] is ids
    past = [torch.zeros(3, 2, 4)]
    out = m.prepare_inputs_for_generation(ids, past=past)
    assert out['mems'] is past and out['input_ids'].tolist() == [[4], [9]]


def test_decode_sequence_groups_rule(pkg):
    """Host logic of the grouped decode (no GPU): groups of 16 sequences from 32 up; explicit requests are clamped to [1, B]; fp32 models
    and large vocabularies stay on one chain of kernels per 64 sequences; any batch size is accepted (a chain takes at most 64 sequences)."""
    import importlib
    import types
    import torch
    decode = importlib.import_module('symbolic-music-generation_b200.decode')

    def fake(dtype, V=1190):
        return types.SimpleNamespace(_E=torch.empty(1, dtype=dtype), config=types.SimpleNamespace(vocab_size=V))
    m = fake(torch.bfloat16)
    assert [decode.sequence_groups(m, B) for B in (1, 8, 16, 31, 32, 47, 48, 64, 100, 256)] == [1, 1, 1, 1, 2, 2, 3, 4, 7, 16]
    assert decode.sequence_groups(m, 7, requested=3) == 3 and decode.sequence_groups(m, 2, requested=5) == 2 and decode.sequence_groups(m, 9, requested=0) == 1
    assert decode.sequence_groups(fake(torch.float32), 64) == 1 and decode.sequence_groups(fake(torch.bfloat16, V=40000), 64) == 1
    assert decode.sequence_groups(fake(torch.float32), 65) == 2 and decode.sequence_groups(fake(torch.float32), 200) == 4 and decode.sequence_groups(m, 100, requested=1) == 2
    bounds = []
    per, rem = divmod(7, 3)
    lo = 0
    for g in range(3):
        hi = lo + per + (1 if g < rem else 0)
        bounds.append((lo, hi))
        lo = hi
    assert bounds == [(0, 3), (3, 5), (5, 7)]


def test_trainer_schedule_is_hf_cosine_with_ceil_warmup(pkg):
    """trainer.warmup_steps / optim.cosine_with_warmup == HF get_cosine_schedule_with_warmup with ceil(ratio * steps) warm-up steps
    (the oracle's LambdaLR restates the same rule; GPU trajectory test compares the two loops step by step)."""
    import importlib
    import math
    trainer = importlib.import_module('symbolic-music-generation_b200.trainer')
    optim = importlib.import_module('symbolic-music-generation_b200.optim')
    assert trainer.warmup_steps(6, 0.3) == 2 and trainer.warmup_steps(1000, 0.1) == 100 and trainer.warmup_steps(10, 0.01) == 1
    for total, warm in [(6, 2), (1000, 100), (10, 0)]:
        for step in range(total + 1):
            if step < warm:
                want = 3e-4 * step / max(1, warm)
            else:
                want = 3e-4 * max(0.0, 0.5 * (1.0 + math.cos(math.pi * (step - warm) / max(1, total - warm))))
            assert abs(optim.cosine_with_warmup(step, total, warm, 3e-4) - want) < 1e-15


def test_argument_errors_are_reported_before_any_launch(built_lib):
    """Error behaviour of the boundary (0 / negative return codes + txl_last_error(), nothing silently different): bad shapes and
    null pointers are refused with TXL_EINVAL and a message by the argument checks that precede every launch — callable without a GPU."""
    L = importlib.import_module(PKG + '._lib')
    lib = L.load()
    EINVAL = -1
    one = ctypes.c_void_p(256)           # a non-null, 16-byte aligned pointer value: the checks below never dereference it

    def refused(rc, needle):
        msg = lib.txl_last_error().decode()
        assert rc == EINVAL and needle in msg, (rc, msg)
    refused(lib.txl_dec_linear(one, 64, one, 64, None, one, 64, 65, 8, 64, 0, 0, 1, None, 0, None), 'dec_linear')          # M > 64
    refused(lib.txl_dec_linear(one, 64, one, 64, None, one, 64, 4, 8, 40, 0, 0, 1, None, 0, None), 'dec_linear')           # K % 32 != 0
    refused(lib.txl_dec_linear(one, 64, one, 64, None, one, 64, 4, 8, 64, 1, 1, 2, None, 0, None), 'split-K')              # ReLU with split-K
    refused(lib.txl_dec_linear(None, 64, one, 64, None, one, 64, 4, 8, 64, 0, 0, 1, None, 0, None), 'dec_linear')          # null operand
    refused(lib.txl_dec_add_ln(one, None, 2, None, one, one, one, 4, 64, 1e-5, None, 0, None), 'dec_add_ln')               # planes missing
    refused(lib.txl_dec_add_ln(one, one, 1, None, one, one, one, 4, 2048, 1e-5, None, 0, None), 'dec_add_ln')              # d > 1024
    refused(lib.txl_decode_attn_pipe(one, one, one, one, one, one, one, 2, 2, 16, 48, 1, None, None, None), 'd_head')       # d_head not in {32,64,128}
    refused(lib.txl_decode_attn_pipe(one, one, one, one, one, one, one, 2, 2, 16, 64, 2, None, None, None), 'splits')       # splits without workspace
    refused(lib.txl_decode_tail(one, 8, None, 2, 16, 1, 0.0, 8, 1.0, 1, 0, one, one, one, 8, 0, one, one, 0, 0, 1, one, one, 8, 1.0, None), 'decode_tail')   # ldl < V / temperature 0
    refused(lib.txl_ntp_acc(one, 3, one, 8, 2, 8, one, None), 'ntp_acc')                                                    # row pitch < T
    refused(lib.txl_clm_labels(one, None, 8, 1, None), 'clm_labels')
    refused(lib.txl_last_index_of(one, 4, 2, 8, 9, one, None), 'last_index_of')                                             # row pitch < T
    refused(lib.txl_sample(one, 2, 16, 1, 0.0, 8, 1.0, one, one, None, None, None), 'sample')                               # temperature 0
    handled = ctypes.c_int(7)
    refused(lib.txl_gemm_add_ln_fwd(None, one, None, one, one, one, one, None, None, None, 512, 512, 512, 512, 512, 1e-5, 0.0, 0, 0, None, ctypes.byref(handled)),
            'gemm_add_ln_fwd')                                                                                               # null operand
    assert handled.value == 0
    refused(lib.txl_gemm_add_ln_fwd(one, one, None, one, one, one, one, one, None, None, 512, 512, 512, 512, 512, 1e-5, 0.0, 0, 0, None, ctypes.byref(handled)),
            'saved together')                                                                                                # z without mean / rstd
    refused(lib.txl_gemm_add_ln_fwd(one, one, None, one, one, one, one, None, None, None, 512, 512, 512, 512, 512, 1e-5, 1.0, 0, 0, None, ctypes.byref(handled)),
            'drop_p')                                                                                                        # drop_p = 1
    assert lib.txl_decode_attn_pipe_ws_bytes(2, 3, 64, 1) == 0 and lib.txl_decode_attn_pipe_ws_bytes(2, 3, 64, 4) == 2 * 3 * 4 * 66 * 4
