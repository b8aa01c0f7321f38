"""Times the LayerNorm(+residual, +dropout) forward / backward kernels alone on the cfg2 activation shape [32768, 512] bf16.
   python profiles/ln_probe.py"""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ops = importlib.import_module('symbolic-music-generation_b200.ops')
N, d = 32768, 512
bf = torch.bfloat16
x, r, dy, dy2 = [(0.5 * torch.randn(N, d, device='cuda')).to(bf) for _ in range(4)]
g, b = torch.ones(d, device='cuda'), torch.zeros(d, device='cuda')
dg, db = torch.zeros(d, device='cuda'), torch.zeros(d, device='cuda')
y, z, mean, rstd = ops.add_ln_fwd(x, r, g, b, 1e-5, 0.1, 1, 3, True)
def tm(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
f = tm(lambda: ops.add_ln_fwd(x, r, g, b, 1e-5, 0.1, 1, 3, True))
bw = tm(lambda: ops.add_ln_bwd(dy, z, g, mean, rstd, dg, db, drop_p=0.1, seed=1, site=3, dy2=dy2))
mb = N * d * 2 / 1e6
print(f'add_ln_fwd {f:7.1f} us  ({4 * mb / f:5.2f} TB/s of x,r -> y,z)   add_ln_bwd {bw:7.1f} us  ({5 * mb / bw:5.2f} TB/s of dy,dy2,z -> dx,dr)')
