"""One launch of the FF2-forward GEMM shape and one of the FF1-forward (bias+ReLU+dropout) shape for an ncu --set full capture."""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ops = importlib.import_module('symbolic-music-generation_b200.ops')
N, d, di = 32768, 512, 2048
bf = torch.bfloat16
x, xh = (0.1 * torch.randn(N, d, device='cuda')).to(bf), (0.1 * torch.randn(N, di, device='cuda')).to(bf)
W1, W2 = (0.1 * torch.randn(di, d, device='cuda')).to(bf), (0.1 * torch.randn(d, di, device='cuda')).to(bf)
b1, b2 = torch.randn(di, device='cuda'), torch.randn(d, device='cuda')
for _ in range(3):
    ops.gemm(x, W1, transB=True, bias=b1, relu=True, drop_p=0.1, seed=1, site=1)
    ops.gemm(xh, W2, transB=True, bias=b2)
torch.cuda.synchronize()
# round 2: the fused Linear + dropout + residual + LayerNorm launches (o_net shape K = 512, FF2 shape K = 2048), launches 6 and 7 of tc_gemm_kernel:
#   ncu --set full --clock-control none -k regex:tc_gemm_kernel -s 6 -c 2 -f -o gpurun_out/r02_gemm_ln python profiles/gemm_ncu_probe.py
Wo = (0.1 * torch.randn(d, d, device='cuda')).to(bf)
gam, bet = torch.ones(d, device='cuda'), torch.zeros(d, device='cuda')
ops.gemm_add_ln_fwd(x, Wo, None, x, gam, bet, 1e-5, 0.1, 1, 2, True)
ops.gemm_add_ln_fwd(xh, W2, b2, x, gam, bet, 1e-5, 0.1, 1, 2, True)
torch.cuda.synchronize()
