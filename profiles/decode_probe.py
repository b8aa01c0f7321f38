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
    # per-stage time of the persistent step (CTA 0's %globaltimer after every grid barrier), summed over the layers
    L_ = importlib.import_module('symbolic-music-generation_b200._lib')
    lib = L_.load()
    nl = cfg.n_layer
    buf = torch.zeros(600, dtype=torch.int64, device='cuda')
    lib.txl_decode_persist_set_timestamps(buf.data_ptr())
    out = model.generate(input_ids=prompt, max_length=16 + 4, do_sample=True, top_k=8, eos_token_id=None, seed=77, use_cuda_graph=False)
    torch.cuda.synchronize()
    lib.txl_decode_persist_set_timestamps(None)
    t = buf.cpu().tolist()
    names = [f'L{l}.{s}' for l in range(nl) for s in ('qtbd', 'att', 'vaon', 'ln1', 'ff1', 'ff2', 'ln2')] + ['head']
    d = [(t[i + 1] - t[i]) / 1e3 for i in range(len(names))]
    agg = {}
    for n, v in zip(names, d):
        k = n.split('.')[-1]
        agg[k] = agg.get(k, 0) + v
    print('layer 5 stages (us):', {n: round(v, 2) for n, v in zip(names, d) if n.startswith('L5.')})
    print('attention item of CTA 0, last layer (cycles): before-wait / wait-full / phase A / consumer barrier / phase B =', t[120:125])
    st = [t[200 + 2 * c] for c in range(148)]; en = [t[201 + 2 * c] for c in range(148)]
    t0 = min(st)
    print('ATT last layer, per CTA (us after the first start): start min/max', 0, round((max(st) - t0) / 1e3, 2), ' end min/median/max', round((min(en) - t0) / 1e3, 2), round((sorted(en)[74] - t0) / 1e3, 2), round((max(en) - t0) / 1e3, 2), ' durations min/max', round(min(e - s_ for e, s_ in zip(en, st)) / 1e3, 2), round(max(e - s_ for e, s_ in zip(en, st)) / 1e3, 2))
    print('LN1 of the last layer on CTA 0 (ns): rows', t[131] - t[130], ' barrier', t[132] - t[131])
    print('persistent step, stage times summed over layers (us):', {k: round(v, 1) for k, v in agg.items()}, 'total', round(sum(d), 1))
