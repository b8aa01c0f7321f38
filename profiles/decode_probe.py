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
    # per-stage time of the one-kernel step (CTA 0's %globaltimer after every stage), summed over the layers
    L_ = importlib.import_module('symbolic-music-generation_b200._lib')
    decode = importlib.import_module('symbolic-music-generation_b200.decode')
    lib = L_.load()
    nl = cfg.n_layer
    buf = torch.zeros(600, dtype=torch.int64, device='cuda')
    cluster = decode.cluster_supported(model, B)
    setter = lib.txl_decode_cluster_set_timestamps if cluster else lib.txl_decode_persist_set_timestamps
    setter(buf.data_ptr())
    out = model.generate(input_ids=prompt, max_length=16 + 4, do_sample=True, top_k=8, eos_token_id=None, seed=77, use_cuda_graph=False)
    torch.cuda.synchronize()
    setter(None)
    t = buf.cpu().tolist()
    stages = ('q+qt+bd', 'att', 'va+on', 'ln1', 'ff1+ff2', 'ln2') if cluster else ('qtbd', 'att', 'vaon', 'ln1', 'ff1', 'ff2', 'ln2')
    names = [f'L{l}.{s}' for l in range(nl) for s in stages] + ['head']
    d = [(t[i + 1] - t[i]) / 1e3 for i in range(len(names))]
    agg = {}
    for n, v in zip(names, d):
        k = n.split('.')[-1]
        agg[k] = agg.get(k, 0) + v
    print('engine:', 'cluster' if cluster else 'persistent', ' layer 5 stages (us):', {n: round(v, 2) for n, v in zip(names, d) if n.startswith('L5.')})
    if cluster:
        print('LN1 of the last layer on CTA 0 (ns): pull partials + slice statistics', t[131] - t[130], ' barrier', t[132] - t[131], ' pull statistics + normalise', t[133] - t[132], ' barrier', t[134] - t[133], ' all-gather', t[135] - t[134])
    print('one-kernel step, stage times of CTA 0 summed over layers (us):', {k: round(v, 1) for k, v in agg.items()}, 'total', round(sum(d), 1))
    if cluster:
        print('co-resident clusters of 8 CTAs:', lib.txl_decode_cluster_max_clusters(cfg.d_inner))
