"""Short batched generation (cfg4 shapes, no CUDA graph so that ncu sees every kernel):
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/decode_launches.csv python profiles/decode_probe.py"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('symbolic-music-generation_b200')
B, NEW = int(os.environ.get('PROBE_B', 64)), int(os.environ.get('PROBE_NEW', 6))
cfg = pkg.MyTransfoXLConfig('small', vocab_size=1190, max_length=1024, mem_len=1024, cutoffs=[], dropout=0.0)
torch.manual_seed(77)
model = pkg.MyTransfoXLLMHeadModel(cfg).cuda().eval()
prompt = torch.randint(1, 1190, (B, 16)).cuda()
out = model.generate(input_ids=prompt, max_length=16 + NEW, do_sample=True, top_k=8, eos_token_id=None, seed=77, use_cuda_graph=False)
torch.cuda.synchronize()
print(out.shape)

if os.environ.get('PROBE_STAGES'):
    L_ = importlib.import_module('symbolic-music-generation_b200._lib')
    lib = L_.load()
    buf = torch.zeros(7 * 12 + 8, dtype=torch.int64, device='cuda')
    lib.txl_decode_fused_set_timestamps(buf.data_ptr())
    out = model.generate(input_ids=prompt, max_length=16 + 4, do_sample=True, top_k=8, eos_token_id=None, seed=77, use_cuda_graph=False, use_fused_step=True)
    torch.cuda.synchronize()
    lib.txl_decode_fused_set_timestamps(None)
    t = buf.cpu().tolist()
    # NOTE: the stamp after the embedding barrier is not emitted, so stage i is labelled with the name of stage i-1's successor:
    # read 'qkv' as attn, 'attn' as o, 'o' as ln1, 'ln1' as ff1, 'ff1' as ff2, 'ff2' as ln2, 'ln2' as the next layer's qkv.
    names = ['embed'] + [f'L{l}.{s}' for l in range(12) for s in ('qkv', 'attn', 'o', 'ln1', 'ff1', 'ff2', 'ln2')] + ['head', 'empty_sync']
    d = [(t[i + 1] - t[i]) / 1e3 for i in range(len(names))]
    agg = {}
    for n, v in zip(names, d):
        k = n.split('.')[-1]
        agg[k] = agg.get(k, 0) + v
    print('layer 5 stages:', {n: round(v, 1) for n, v in zip(names, d) if n.startswith('L5.')})
    print('fused step stage times (us, summed over layers):', {k: round(v, 1) for k, v in agg.items()}, 'total', round(sum(d), 1))
