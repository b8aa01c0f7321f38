import sys, importlib, torch
import os; ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
pkg = importlib.import_module('symbolic-music-generation_b200')
from conftest import make_pair
torch.manual_seed(0)
ref, model = make_pair(pkg, 'bf16', vocab_size=1190, d_model=512, n_head=8, d_head=64, d_inner=2048, n_layer=1, mem_len=1024, clamp_len=1024)
g = torch.Generator().manual_seed(77)
ids = torch.randint(0, 1190, (2, 1024), generator=g); labels = ids.clone()
mems = [0.5 * torch.randn(1024, 2, 512)]
ref.eval(); model.eval()
with torch.no_grad():
    ro = ref(input_ids=ids, mems=mems, labels=labels.clone())
    out = model(input_ids=ids.cuda(), mems=[m.cuda() for m in mems], labels=labels.cuda())
a, b = out.losses.cpu(), ro.losses
rel = (a - b).abs() / b.abs()
print('loss: mean rel', rel.mean().item(), 'p99', rel.flatten().kthvalue(int(0.99 * rel.numel())).values.item(), 'max', rel.max().item())
w = rel.flatten().argmax().item(); bi, ti = divmod(w, 1023)
print('worst token', bi, ti, 'ours', a[bi, ti].item(), 'ref', b[bi, ti].item())
# position profile of error
e = (a - b).abs().mean(0)
print('abs err by position block of 128:', [round(e[i:i + 128].mean().item(), 4) for i in range(0, 1023, 128)])
lg, lr = out.logits.float().cpu(), ro.logits
d = (lg - lr).abs()
print('logprob abs err: mean', d.mean().item(), 'max', d.max().item(), 'at', divmod(d.flatten().argmax().item(), 1190))
# hidden check through mems output of layer 0 input is embedding only; compare logits row-wise cosine
import os
os.environ['X'] = '1'
ref2, model32 = make_pair(pkg, 'fp32', vocab_size=1190, d_model=512, n_head=8, d_head=64, d_inner=2048, n_layer=1, mem_len=1024, clamp_len=1024)
model32.load_state_dict(ref.state_dict()); model32.cuda().eval()
with torch.no_grad():
    o32 = model32(input_ids=ids.cuda(), mems=[m.cuda() for m in mems], labels=labels.cuda())
r32 = (o32.losses.cpu() - b).abs() / b.abs()
print('fp32 mode: max rel', r32.max().item())
