import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
PKG = 'symbolic-music-generation_b200'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu under gpurun)')


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests need a B200 and the built library: skip them (instead of failing with 'no NVIDIA driver') on CPU boxes."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason='needs a B200 (run with -m gpu under gpurun)')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def pkg():
    return importlib.import_module(PKG)


@pytest.fixture(scope='session')
def ops(pkg):
    return importlib.import_module(PKG + '.ops')


@pytest.fixture(scope='session')
def built_lib():
    """Build (or reuse) the in-tree library; CPU tests only load it and look at its symbols."""
    b = importlib.import_module(PKG + '.build')
    return b.build_library()


def make_pair(pkg, mode, seed=77, **kw):
    """(oracle model on CPU, B200 model on cuda) sharing one state_dict."""
    import torch
    from oracle.txl_ref import RefConfig, RefTransfoXLLMHeadModel
    torch.manual_seed(seed)
    base = dict(vocab_size=422, d_model=128, n_head=4, n_layer=2, d_head=32, d_inner=256, mem_len=32, clamp_len=1024, dropout=0.0)
    base.update(kw)
    base.setdefault('cutoffs', [])
    ref = RefTransfoXLLMHeadModel(RefConfig(d_embed=base['d_model'], **base))
    cfg = pkg.MyTransfoXLConfig('debug', compute_dtype=mode, d_embed=base['d_model'], **base)
    model = pkg.MyTransfoXLLMHeadModel(cfg)
    missing = model.load_state_dict(ref.state_dict(), strict=True)
    model.to('cuda')
    return ref, model
