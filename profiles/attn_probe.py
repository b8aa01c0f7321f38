"""Runs the relative-position attention forward + backward kernels once (plus warm-up) on the cfg2 shape, for ncu captures:
  ncu --set full --clock-control none --import-source on -k regex:relattn -s 8 -c 4 -o gpurun_out/attn python profiles/attn_probe.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ops = importlib.import_module('symbolic-music-generation_b200.ops')

B, T, M, H, dh = int(os.environ.get('PROBE_B', 32)), 1024, 1024, 8, 64
d = H * dh
torch.manual_seed(0)
dt = torch.bfloat16
qkv = (0.5 * torch.randn(B * T, 3 * d, device='cuda')).to(dt)
kvm = (0.5 * torch.randn(B * M, 2 * d, device='cuda')).to(dt)
r = (0.5 * torch.randn(T + M, d, device='cuda')).to(dt)
rwb, rrb = 0.1 * torch.randn(d, device='cuda'), 0.1 * torch.randn(d, device='cuda')
band = ops.make_band(T, M, M, 1024, 1)
dout = torch.randn(B * T, d, device='cuda').to(dt)
dqkv, dkvm = torch.empty_like(qkv), torch.empty_like(kvm)
dr, drwb, drrb = torch.zeros(T + M, d, device='cuda'), torch.zeros(d, device='cuda'), torch.zeros(d, device='cuda')
SAVE = os.environ.get('PROBE_SAVE', '1') == '1'      # forward leaves its P~ tiles for the backward (the training path) vs. full recompute
for it in range(3):
    out, lse, saved = ops.relattn_fwd(qkv[:, :d], kvm[:, :d], kvm[:, d:], qkv[:, d:2 * d], qkv[:, 2 * d:], r, rwb, rrb, B, T, H, dh, band, save=True)
    if not SAVE:
        saved = None
    ops.relattn_bwd(qkv[:, :d], kvm[:, :d], kvm[:, d:], qkv[:, d:2 * d], qkv[:, 2 * d:], r, rwb, rrb, out, lse, dout, dqkv[:, :d], dkvm[:, :d], dkvm[:, d:],
                    dqkv[:, d:2 * d], dqkv[:, 2 * d:], dr, drwb, drrb, B, T, H, dh, band, saved=saved)
torch.cuda.synchronize()
if os.environ.get('PROBE_TIME'):
    def tm(fn, n=5):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    f = lambda: ops.relattn_fwd(qkv[:, :d], kvm[:, :d], kvm[:, d:], qkv[:, d:2 * d], qkv[:, 2 * d:], r, rwb, rrb, B, T, H, dh, band)
    fs = lambda: ops.relattn_fwd(qkv[:, :d], kvm[:, :d], kvm[:, d:], qkv[:, d:2 * d], qkv[:, 2 * d:], r, rwb, rrb, B, T, H, dh, band, save=True)
    b = lambda: ops.relattn_bwd(qkv[:, :d], kvm[:, :d], kvm[:, d:], qkv[:, d:2 * d], qkv[:, 2 * d:], r, rwb, rrb, out, lse, dout, dqkv[:, :d], dkvm[:, :d], dkvm[:, d:],
                                dqkv[:, d:2 * d], dqkv[:, 2 * d:], dr, drwb, drrb, B, T, H, dh, band, saved=saved)
    print(f"TXL_DBG={os.environ.get('TXL_DBG', '0')} saved={saved is not None}  fwd {tm(f):8.1f} us  fwd+save {tm(fs):8.1f} us   bwd(all passes) {tm(b):8.1f} us")
print('probe done')
