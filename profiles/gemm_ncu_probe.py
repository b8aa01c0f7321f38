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
