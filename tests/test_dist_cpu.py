"""Host-side logic of the data-parallel path on CPU: gradient bucket construction and the overlapped bucketed all-reduce
(world_size 2, gloo), and the sequence sharding used by batched sampling.  The GPU arithmetic is covered by the -m gpu tests."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

PKG = 'symbolic-music-generation_b200'


class FakeModel:
    """Stands in for MyTransfoXLLMHeadModel's flat-gradient interface (no CUDA needed)."""

    def __init__(self, n_layer=6, per_layer=1000, emb=300):
        self.n_layer, self.per, self.emb = n_layer, per_layer, emb
        self._flat_numel = emb + n_layer * per_layer
        self._gflat = torch.zeros(self._flat_numel)
        self._grad_hook = None

    def _ensure_engine(self):
        pass

    def layer_param_ranges(self):
        out = [(0, self.emb)]
        for i in range(self.n_layer):
            out.append((self.emb + i * self.per, self.emb + (i + 1) * self.per))
        return out


def test_bucket_construction():
    pdist = importlib.import_module(PKG + '.dist')
    m = FakeModel(n_layer=6, per_layer=1000, emb=300)
    b = pdist.GradBucketer(m, bucket_mb=2500 * 4 / 2 ** 20)       # capacity 2500 floats => 2 layers per bucket
    assert [(t, e - s) for t, s, e in b.buckets] == [(4, 2000), (2, 2000), (0, 2000)]
    covered = sorted((s, e) for _, s, e in b.buckets)
    assert covered[0][0] == 300 and covered[-1][1] == m._flat_numel
    assert all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
    assert b.emb_range == (0, 300)
    big = pdist.GradBucketer(FakeModel(), bucket_mb=1e-6)          # smaller than one layer: one layer per bucket
    assert len(big.buckets) == 6


def test_shard_sequences():
    pdist = importlib.import_module(PKG + '.dist')
    for n, w in ((64, 8), (10, 4), (3, 8)):
        spans = [pdist.shard_sequences(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _worker(rank, world, port, ret):
    os.environ['MASTER_ADDR'], os.environ['MASTER_PORT'] = '127.0.0.1', str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    pdist = importlib.import_module(PKG + '.dist')
    m = FakeModel()
    bucketer = pdist.GradBucketer(m, bucket_mb=2500 * 4 / 2 ** 20)
    torch.manual_seed(100 + rank)
    local = torch.randn(m._flat_numel)
    # emulate the hand-scheduled backward: layer L-1 .. 0 finish in order, then the embedding
    ranges = m.layer_param_ranges()
    for li in range(m.n_layer - 1, -1, -1):
        s, e = ranges[li + 1]
        m._gflat[s:e] = local[s:e]
        m._grad_hook('layer', li)
    m._gflat[0:m.emb] = local[0:m.emb]
    m._grad_hook('embed', -1)
    all_local = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(all_local, local)
    want = torch.stack(all_local).mean(0)
    ok = torch.allclose(m._gflat, want, atol=1e-6) and bucketer.launched == len(bucketer.buckets) + 1
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ret[0] and ret[1]
