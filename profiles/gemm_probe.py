"""Times the cfg2 GEMM shapes of one layer (fwd / dgrad / wgrad) alone with CUDA events: TFLOP/s per shape.
   python profiles/gemm_probe.py"""
import importlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ops = importlib.import_module('symbolic-music-generation_b200.ops')
N, d, di = 32768, 512, 2048
bf = torch.bfloat16
torch.manual_seed(0)


def t(*shape):
    return (0.1 * torch.randn(*shape, device='cuda')).to(bf)


x, x3, xh = t(N, d), t(N, 3 * d), t(N, di)
Wqkv, Wo, W1, W2 = t(3 * d, d), t(d, d), t(di, d), t(d, di)
b1, b2 = torch.randn(di, device='cuda'), torch.randn(d, device='cuda')
g3, g1, gh = torch.zeros(3 * d, d, device='cuda'), torch.zeros(d, d, device='cuda'), torch.zeros(di, d, device='cuda')
cs = torch.zeros(di, device='cuda')
gam, bet = torch.ones(d, device='cuda'), torch.zeros(d, device='cuda')
cases = {
    'fwd qkv   x[N,512] Wqkv^T -> [N,1536]': (lambda: ops.gemm(x, Wqkv, transB=True), 2 * N * d * 3 * d),
    'fwd o     x[N,512] Wo^T   -> [N,512]': (lambda: ops.gemm(x, Wo, transB=True), 2 * N * d * d),
    'fwd ff1   +bias+relu+dropout -> [N,2048]': (lambda: ops.gemm(x, W1, transB=True, bias=b1, relu=True, drop_p=0.1, seed=1, site=1), 2 * N * d * di),
    'fwd ff1   +bias+relu          -> [N,2048]': (lambda: ops.gemm(x, W1, transB=True, bias=b1, relu=True), 2 * N * d * di),
    'fwd ff2   h[N,2048] W2^T +bias -> [N,512]': (lambda: ops.gemm(xh, W2, transB=True, bias=b2), 2 * N * d * di),
    'fwd o   + add_ln (two kernels)': (lambda: ops.add_ln_fwd(x, ops.gemm(x, Wo, transB=True), gam, bet, 1e-5, 0.1, 1, 2, True), 2 * N * d * d),
    'fwd o   + dropout+residual+LN fused': (lambda: ops.gemm_add_ln_fwd(x, Wo, None, x, gam, bet, 1e-5, 0.1, 1, 2, True), 2 * N * d * d),
    'fwd o   + fused, eval (no z / stats, p = 0)': (lambda: ops.gemm_add_ln_fwd(x, Wo, None, x, gam, bet, 1e-5, 0.0, 1, 2, False), 2 * N * d * d),
    'fwd ff2 + add_ln (two kernels)': (lambda: ops.add_ln_fwd(x, ops.gemm(xh, W2, transB=True, bias=b2), gam, bet, 1e-5, 0.1, 1, 2, True), 2 * N * d * di),
    'fwd ff2 + dropout+residual+LN fused': (lambda: ops.gemm_add_ln_fwd(xh, W2, b2, x, gam, bet, 1e-5, 0.1, 1, 2, True), 2 * N * d * di),
    'dgrad ff2 df[N,512] W2 +mask+dropout+colsum -> [N,2048]': (lambda: ops.gemm(x, W2, mask_pos_aux=xh, colsum=cs, drop_p=0.1, seed=1, site=1), 2 * N * d * di),
    'dgrad ff2 +mask+dropout (no colsum)': (lambda: ops.gemm(x, W2, mask_pos_aux=xh, drop_p=0.1, seed=1, site=1), 2 * N * d * di),
    'dgrad ff2 +mask only': (lambda: ops.gemm(x, W2, mask_pos_aux=xh), 2 * N * d * di),
    'dgrad ff2 plain': (lambda: ops.gemm(x, W2), 2 * N * d * di),
    'colsum [N,2048]': (lambda: ops.colsum(xh, cs), 1),
    'dgrad ff1 dh[N,2048] W1 -> [N,512]': (lambda: ops.gemm(xh, W1), 2 * N * d * di),
    'dgrad qkv dqkv[N,1536] Wqkv -> [N,512]': (lambda: ops.gemm(x3, Wqkv), 2 * N * d * 3 * d),
    'wgrad qkv dqkv^T x -> [1536,512] fp32 +=': (lambda: ops.gemm(x3, x, transA=True, out=g3, accumulate=True), 2 * N * d * 3 * d),
    'wgrad o   dao^T vec -> [512,512] fp32 +=': (lambda: ops.gemm(x, x, transA=True, out=g1, accumulate=True), 2 * N * d * d),
    'wgrad ff1 dh^T y1 -> [2048,512] fp32 +=': (lambda: ops.gemm(xh, x, transA=True, out=gh, accumulate=True), 2 * N * d * di),
}
res = {}
for name, (fn, flops) in cases.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res[name] = dict(us=round(ms * 1e3, 1), tflops=round(flops / ms / 1e9, 1))
    print(f'{ms * 1e3:9.1f} us  {flops / ms / 1e9:8.1f} TFLOP/s  {name}')
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'gemm_probe.json'), 'w'), indent=1)
