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
