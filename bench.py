#!/usr/bin/env python
"""bench.py — Transformer-XL training-step throughput on the BASELINE.json configuration.

A "step" = one full training pass of the hot path over one batch: forward (embedding, 12 layers of relative-position band
attention + position-wise FF, LM head + log-softmax + NLL) with CARRIED NON-ZERO mems, backward, bucketed gradient
all-reduce (N > 1), global-norm clip + AdamW.  Workload = BASELINE.json configs[1]:
  Transformer-XL music LM, 12 layers, d_model 512, 8 heads, d_inner 2048, seq 1024, mem_len 1024, vocab 1190, bf16,
  per-GPU batch 32 (weak scaling), dropout 0.1, synthetic uniform token ids (seed 77), random-init weights.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
`--impl reference` times the reference's CPU implementation of the same path (the oracle restatement of HF 4.25.1
TransfoXL — the reference's own dependency cannot be installed here) on the host cores.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = 'symbolic-music-generation_b200'

CFG2 = dict(model_size='small', vocab_size=1190, max_length=1024, mem_len=1024, cutoffs=[])      # reference train.py:521-527 override
FLOP_PER_TOKEN = {  # SURVEY §8d, band-aware, fwd+bwd with real (carried) mems
    'cfg2': 3.6856e8,
}


def flop_per_token(L, d, di, T, M, V, Kb):
    fwd_seq = L * (2 * T * d * d + 4 * (T + M) * d * d + 2 * T * d * d + 4 * T * d * di + 6 * T * Kb * d) + 2 * T * d * V
    bwd_seq = 2 * fwd_seq - L * 4 * M * d * d
    return (fwd_seq + bwd_seq) / T, fwd_seq / T


_REAL_STDOUT = None


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs (B200_PROFILING.md recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100', '-i', str(self.idx)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons),
                    power_w_max=max(pw) if pw else None, samples=len(sm))


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return dict(tf=p['bf16_tflops_sustained'], tf_burst=p['bf16_tflops'], hbm=p['hbm_gbs'], src='measured (MEASURED_PEAKS.json)')
    except Exception:
        return dict(tf=1400.0, tf_burst=1590.0, hbm=6650.0, src='fallback (B200_PROFILING.md)')


# ----------------------------------------------------------------------------- CPU baseline / reference arm
def cpu_reference_tokens_per_s(steps, warmup, budget_s=25.0, train=True):
    """The reference's CPU path (oracle restatement) on this box's host cores: cfg2 training step on a bounded sample
    (B=1 sequence of T=1024 with mem_len 1024 carried mems) — same model, same shapes per sequence as the GPU arm."""
    import torch
    from oracle.txl_ref import RefConfig, RefTransfoXLLMHeadModel
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(77)
    cfg = RefConfig.from_preset('small', vocab_size=1190, max_length=1024, mem_len=1024, cutoffs=[])
    model = RefTransfoXLLMHeadModel(cfg).train()
    opt = torch.optim.AdamW(model.parameters(), lr=3e-4, weight_decay=0.01)
    g = torch.Generator().manual_seed(77)
    B, T = 1, 1024
    ids = torch.randint(0, cfg.vocab_size, (B, T), generator=g)
    mems = [0.5 * torch.randn(cfg.mem_len, B, cfg.d_model) for _ in range(cfg.n_layer)]
    times = []
    t_start = time.time()
    for i in range(warmup + steps):
        t0 = time.time()
        out = model(input_ids=ids, mems=mems, labels=ids.clone())
        out.loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step(); opt.zero_grad(set_to_none=True)
        mems = [m.detach() for m in out.mems]
        dt = time.time() - t0
        if i >= warmup:
            times.append(dt)
        if time.time() - t_start > budget_s and len(times) >= 1:
            break
    ms = 1000 * sum(times) / len(times)
    return dict(value=B * T / (ms / 1000), ms_per_step=ms, cores=cores, steps=len(times),
                sample=f'cfg2 model (12L d512 T1024 mem1024 V1190) fp32 train step fwd+bwd+clip+AdamW, B=1 sequence with carried mems, '
                       f'{len(times)} timed steps, torch {torch.__version__} CPU, {cores} threads')


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    r = cpu_reference_tokens_per_s(max(1, min(args.steps, 3)), max(1, min(args.warmup, 1)), budget_s=60.0)
    line = {
        'impl': 'reference', 'metric': 'TXL train tokens/s', 'value': r['value'], 'unit': 'tokens/s', 'n_gpus': args.gpus, 'steps': r['steps'],
        'warmup': min(args.warmup, 1), 'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'cfg2: Transformer-XL 12L d512 H8 di2048 T1024 mem1024 V1190 training step (CPU reference arm, bounded sample B=1)',
                   'same_config_as_ours': False,
                   'bounded_sample': f"B=1 sequence x {r['steps']} timed steps on {r['cores']} host threads in fp32 (ours: B=32 per GPU, bf16); a CPU baseline, "
                                     'reported beside the GPU number, not a like-for-like speed-up'},
        'cpu_baseline': {'value': r['value'], 'unit': 'tokens/s', 'cores': r['cores'], 'kind': 'port', 'sample': r['sample']},
        'e2e': {'value': r['value'], 'unit': 'tokens/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'reference dependency transformers==4.25.1 is not installable here; this is the oracle restatement (kind=port)',
    }
    emit(line)


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module(PKG)
    ops = importlib.import_module(PKG + '.ops')
    L_ = importlib.import_module(PKG + '._lib')
    optim = importlib.import_module(PKG + '.optim')
    pdist = importlib.import_module(PKG + '.dist')

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    lib = L_.load()
    if args.decode_only:     # development aid: the cfg4 decode object alone (not a bench line the driver reads)
        cfg0 = pkg.MyTransfoXLConfig(compute_dtype=args.dtype, dropout=0.0, **dict(CFG2, mem_len=args.mem_len))
        line = decode_probe(torch, pkg, pdist, cfg0, args, dev, rank, world)
        if rank == 0:
            emit(line)
        if world > 1:
            dist.destroy_process_group()
        return

    B, T = args.batch, args.seq
    kw = dict(CFG2)
    kw.update(max_length=T, mem_len=args.mem_len)
    cfg = pkg.MyTransfoXLConfig(compute_dtype=args.dtype, dropout=args.dropout, **kw)
    torch.manual_seed(77)
    model = pkg.MyTransfoXLLMHeadModel(cfg).to(dev).train()
    model._ensure_engine()
    if world > 1:       # identical replicas: broadcast rank 0's flat parameters
        dist.broadcast(model._flat, 0)
        model.mark_params_dirty()
    opt = optim.FusedAdamW(model, lr=3e-4, weight_decay=0.01, max_grad_norm=1.0)
    bucketer = pdist.GradBucketer(model, bucket_mb=25.0) if world > 1 else None

    V = cfg.vocab_size
    g = torch.Generator().manual_seed(77 + rank)
    n_batches = 4
    host_ids = [torch.randint(0, V, (B, T), generator=g).pin_memory() for _ in range(n_batches)]
    host_lab = []
    for ids in host_ids:        # padded variant: 25 % of rows end in a pad tail (-100 labels)
        lab = ids.clone()
        for b in range(0, B, 4):
            tail = int(torch.randint(1, T // 4, (1,), generator=g))
            lab[b, T - tail:] = -100
        host_lab.append(lab.pin_memory())
    dev_ids = [x.to(dev) for x in host_ids]
    dev_lab = [x.to(dev) for x in host_lab]

    state = {'mems': None}

    def step_resident(i):
        out = model(input_ids=dev_ids[i % n_batches], mems=state['mems'], labels=dev_lab[i % n_batches])
        out.loss.backward()
        opt.step()
        opt.zero_grad()
        state['mems'] = out.mems           # carried, non-zero mems (SURVEY §8d zero-mems caveat)
        return out.loss

    def step_e2e(i):
        ids = host_ids[i % n_batches].to(dev, non_blocking=True)
        lab = host_lab[i % n_batches].to(dev, non_blocking=True)
        out = model(input_ids=ids, mems=state['mems'], labels=lab)
        out.loss.backward()
        opt.step()
        opt.zero_grad()
        state['mems'] = out.mems
        return float(out.loss.item())      # device->host read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        l0 = lib.txl_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for i in range(steps):
            last = fn(warmup + i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.txl_launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, launches, last

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_step, launches, last_loss = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _, last_e2e = timed(step_e2e, max(3, args.steps // 2), 3)

    tokens = B * T * world
    value = tokens / (ms_step / 1000)
    e2e_value = tokens / (ms_e2e / 1000)
    fpt, fpt_fwd = flop_per_token(cfg.n_layer, cfg.d_model, cfg.d_inner, T, args.mem_len, V, min(args.mem_len, T + args.mem_len))
    peaks = measured_peaks()
    step_tf = value / world * fpt / 1e12

    # ---- dominant kernel, timed alone with CUDA events on the launching stream
    dom = dominant_kernel_probe(torch, ops, model, cfg, B, T, args.mem_len, dev)

    line = {
        'metric': 'TXL train tokens/s', 'value': value, 'unit': 'tokens/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': args.dtype, 'data': 'synthetic',
        'config': {'workload': f'cfg2: Transformer-XL 12L d512 H8 dh64 di2048 T{T} mem{args.mem_len} V{V} bf16 training step '
                               f'(fwd+bwd+clip+AdamW, carried non-zero mems, dropout {args.dropout})',
                   'batch_per_gpu': B, 'global_batch': B * world, 'seq_len': T, 'mem_len': args.mem_len, 'parallelism': f'dp{world}',
                   'l2_policy': 'inputs larger than L2: ~5 GB of activations and 4 rotating batches per step; no explicit flush',
                   'flop_per_token': fpt, 'final_loss': float(last_loss.item()) if last_loss is not None else None},
        'e2e': {'value': e2e_value, 'unit': 'tokens/s', 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': 2 * B * T * 8, 'd2h_bytes_per_step': 4,
                'api': 'MyTransfoXLLMHeadModel.forward(input_ids, mems, labels) from pinned host ids/labels -> loss.backward() -> FusedAdamW.step() -> loss.item()'},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': dom,
        'roofline_step': {'bound': 'tensor', 'achieved': step_tf, 'peak': peaks['tf'], 'unit': 'TFLOP/s', 'frac': step_tf / peaks['tf'],
                          'note': f'whole step, algorithmic FLOP/token {fpt:.4e} (SURVEY §8d) / per-GPU tokens/s; peak = sustained bf16, {peaks["src"]}'},
    }
    if world > 1:
        line['dp_check'] = dp_check(torch, dist, model, bucketer, dev_ids[0], dev_lab[0], state['mems'], world)
    # free the cfg2 training state before the other workloads
    del opt, bucketer
    state['mems'] = None
    model._grad_hook = None
    model.zero_grad()
    model._gflat = None
    del model
    torch.cuda.empty_cache()
    if not args.no_cfg5:
        line['cfg5'] = cfg5_probe(torch, dist, pkg, optim, pdist, args, dev, rank, world)
    if not args.no_decode:
        line['decode'] = decode_probe(torch, pkg, pdist, cfg, args, dev, rank, world)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_tokens_per_s(1, 1, budget_s=25.0)
        line['cpu_baseline'] = {'value': r['value'], 'unit': 'tokens/s', 'cores': r['cores'], 'kind': 'port', 'sample': r['sample']}
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def dp_check(torch, dist, model, bucketer, ids, labels, mems, world):
    """the data-parallel step must equal the single-GPU step on the same global batch.  One extra (untimed) step, run twice on
    every rank with identical dropout seeds: (1) without the bucketer -> this rank's own gradients; all-gather two layers' slices and average
    them on the device = what one GPU holding the global batch would compute (per-replica `losses[losses != 0].mean()`, then the mean of
    replicas: DDP's arithmetic); (2) with the bucketed NCCL all-reduce fired from inside the backward.  Reports the largest
    absolute difference between (2) and the average of (1), and the scale of the gradients."""
    ranges = model.layer_param_ranges()
    picks = {'layer0': ranges[1], f'layer{len(ranges) - 2}': ranges[-1], 'embedding': ranges[0]}
    seed0 = model._step_seed

    def one_pass(hook):
        model._grad_hook = hook
        model._step_seed = seed0
        model.zero_grad()
        model(input_ids=ids, mems=mems, labels=labels).loss.backward()
        return model.flat_grads()
    g_local = one_pass(None)
    want = {}
    for k, (s, e) in picks.items():
        mine = g_local[s:e].clone()
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        want[k] = torch.stack(parts).mean(0)
    g_red = one_pass(bucketer)
    out = {'what': 'max |bucketed all-reduce gradient - mean over ranks of the un-reduced gradients| (fp32), same batch and dropout seed', 'world': world}
    worst = 0.0
    for k, (s, e) in picks.items():
        err = float((g_red[s:e] - want[k]).abs().max().item())
        out[k] = {'max_abs_err': err, 'max_abs_grad': float(want[k].abs().max().item())}
        worst = max(worst, err / max(out[k]['max_abs_grad'], 1e-30))
    out['max_rel_err'] = worst
    model.zero_grad()
    return out


def cfg5_probe(torch, dist, pkg, optim, pdist, args, dev, rank, world):
    """BASELINE configs[4]: longer-seq Transformer-XL (seq 2048, mem_len 2048, clamp_len 1024 = the `small` preset's), bf16 training step,
    16 sequences per GPU, data-parallel over the ranks with the same bucketed all-reduce; carried non-zero mems, dropout 0.1."""
    T = M = 2048
    B = args.cfg5_batch
    cfg = pkg.MyTransfoXLConfig(compute_dtype=args.dtype, dropout=args.dropout, **dict(CFG2, max_length=T, mem_len=M))
    torch.manual_seed(77)
    model = pkg.MyTransfoXLLMHeadModel(cfg).to(dev).train()
    model._ensure_engine()
    if world > 1:
        dist.broadcast(model._flat, 0)
        model.mark_params_dirty()
    opt = optim.FusedAdamW(model, lr=3e-4, weight_decay=0.01, max_grad_norm=1.0)
    bucketer = pdist.GradBucketer(model, bucket_mb=25.0) if world > 1 else None
    g = torch.Generator().manual_seed(770 + rank)
    ids = [torch.randint(0, cfg.vocab_size, (B, T), generator=g).to(dev) for _ in range(2)]
    state = {'mems': None}

    def step(i):
        out = model(input_ids=ids[i % 2], mems=state['mems'], labels=ids[i % 2])
        out.loss.backward()
        opt.step()
        opt.zero_grad()
        state['mems'] = out.mems
    steps, warm = max(3, args.steps // 2), 3
    for i in range(warm):
        step(i)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warm + i)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    fpt, _ = flop_per_token(cfg.n_layer, cfg.d_model, cfg.d_inner, T, M, cfg.vocab_size, M)
    peaks = measured_peaks()
    value = B * T * world / (ms / 1e3)
    tf = value / world * fpt / 1e12
    model._grad_hook = None
    del opt, bucketer, model
    state['mems'] = None
    torch.cuda.empty_cache()
    return {'metric': 'TXL train tokens/s', 'value': value, 'unit': 'tokens/s', 'ms_per_step': ms, 'steps': steps, 'warmup': warm, 'n_gpus': world,
            'config': {'workload': f'cfg5: Transformer-XL 12L d512 H8 dh64 di2048 T{T} mem{M} clamp_len {cfg.clamp_len} V{cfg.vocab_size} bf16 training step '
                                   f'(fwd+bwd+clip+AdamW, carried non-zero mems, dropout {args.dropout})', 'batch_per_gpu': B, 'global_batch': B * world,
                       'parallelism': f'dp{world}', 'flop_per_token': fpt},
            'roofline_step': {'bound': 'tensor', 'achieved': tf, 'peak': peaks['tf'], 'unit': 'TFLOP/s', 'frac': tf / peaks['tf'],
                              'note': 'algorithmic FLOP/token (SURVEY 8d, band-aware) x per-GPU tokens/s; peak = sustained bf16, ' + peaks['src']}}


def decode_probe(torch, pkg, pdist, cfg_train, args, dev, rank, world):
    """cfg4: batched autoregressive generation, `--decode-seqs` sequences in total sharded over the ranks (no collective), prompt 16
    tokens, `--decode-new` new tokens, top-k 8 sampling, mems cache of 1024, eos disabled for timing.  Timed through generate()
    (public API; CUDA events around the call, max over ranks); the prompt forward and cache build are inside the timed region."""
    import torch.distributed as dist
    lo, hi = pdist.shard_sequences(args.decode_seqs, rank, world)
    Bl = hi - lo
    cfg = pkg.MyTransfoXLConfig(compute_dtype=args.dtype, dropout=0.0, **dict(CFG2, mem_len=args.mem_len))
    torch.manual_seed(77)
    model = pkg.MyTransfoXLLMHeadModel(cfg).to(dev).eval()
    g = torch.Generator().manual_seed(77)
    prompt = torch.randint(1, cfg.vocab_size, (args.decode_seqs, 16), generator=g)[lo:hi].to(dev)
    kw = dict(do_sample=True, top_k=8, temperature=1.0, renormalize_logits=True, eos_token_id=None, seed=77, seq_offset=lo)    # seed: same tokens under any sharding
    model.generate(input_ids=prompt, max_length=16 + 8, **kw)          # warm-up (kernel attributes, allocator)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = model.generate(input_ids=prompt, max_length=16 + args.decode_new, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    assert out.shape == (Bl, 16 + args.decode_new)
    peaks = measured_peaks()
    dmod = importlib.import_module(PKG + '.decode')
    n_groups = dmod.sequence_groups(model, Bl)
    if dmod.cluster_supported(model, Bl):
        engine = ('cluster engine: one kernel per step, thread-block clusters (CTA = attention head) carry their sequences through all layers, DSMEM '
                  'hand-over, hidden-state ring with absorbed K/V projections; + the sampling tail kernel; CUDA-graph replay')
        ring_note = 'the ring that is read IS the hidden-state mems (1x the mems term); every cluster streams the weights from L2'
    elif dmod.persist_supported(model, Bl):
        engine = 'persistent cooperative kernel with grid barriers over the hidden-state ring; + the sampling tail kernel; CUDA-graph replay'
        ring_note = 'the ring that is read IS the hidden-state mems (1x the mems term)'
    else:
        engine = (f'launch chain: projected-K/V ring cache, CUDA-graph step with {n_groups} sequence group(s) as parallel branches, programmatic '
                  'dependent launch')
        ring_note = 'the K/V cache actually read is 2x the mems term'
    L, M, d = cfg.n_layer, cfg.mem_len, cfg.d_model
    n_params = sum(p.numel() for p in model.parameters())
    step_bytes = Bl * L * M * d * 2 + n_params * 2 + L * M * d * 2        # hidden-state mems + weights + R tables, bf16
    ms_step = ms / args.decode_new
    ach = step_bytes / (ms_step / 1e3) / 1e9
    return {'metric': 'TXL decode tokens/s', 'value': args.decode_seqs * args.decode_new / (ms / 1e3), 'unit': 'tokens/s', 'ms_per_token_step': ms_step,
            'config': {'workload': f'cfg4: {args.decode_seqs} sequences ({Bl} per GPU), prompt 16, {args.decode_new} new tokens, top_k 8, mem_len {M}, ' + engine},
            'roofline': {'bound': 'hbm', 'achieved': ach, 'peak': peaks['hbm'], 'unit': 'GB/s', 'frac': ach / peaks['hbm'], 'traffic': None,
                         'algorithmic_bytes_per_step': step_bytes,
                         'note': 'denominator bytes = hidden-state mems once + weights + R tables (SURVEY §8d); ' + ring_note}}


# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dQ pass (ncu --set full, cfg2 shape, 32 sequences)
DQ_TRAFFIC = 1408.2e6
DQ_TRAFFIC_SRC = ('dram__bytes_read.sum + dram__bytes_write.sum of one launch (809 MB read: saved P~ tiles + operands; 600 MB written: bf16 dS tiles), '
                  'profiles/r02_ncu_full_attention.txt')


def dominant_kernel_probe(torch, ops, model, cfg, B, T, M, dev):
    """Times the attention kernels alone (CUDA events on the launching stream, inputs of 32 sequences = 300 MB, far beyond L2).
    `roofline` describes the largest single kernel of the step — the dQ pass of the attention backward (25 % of the step in
    profiles/r01_launches_train_step_summary.txt); the forward kernel is reported next to it."""
    peaks = measured_peaks()
    d, H, dh = cfg.d_model, cfg.n_head, cfg.d_head
    dt = torch.bfloat16 if cfg.compute_dtype == 'bf16' else torch.float32
    torch.manual_seed(1)
    qkv = (0.5 * torch.randn(B * T, 3 * d, device=dev)).to(dt)
    kvm = (0.5 * torch.randn(B * M, 2 * d, device=dev)).to(dt)
    r = (0.5 * torch.randn(T + M, d, device=dev)).to(dt)
    rwb = 0.1 * torch.randn(d, device=dev)
    rrb = 0.1 * torch.randn(d, device=dev)
    band = ops.make_band(T, M, cfg.mem_len, cfg.clamp_len, cfg.same_length)
    dout = torch.randn(B * T, d, device=dev).to(dt)
    dqkv, dkvm = torch.empty_like(qkv), torch.empty_like(kvm)
    dr, drwb, drrb = torch.zeros(T + M, d, device=dev), torch.zeros(d, device=dev), torch.zeros(d, device=dev)

    def fwd():      # as the training step calls it: the forward also leaves its soft-max numerators (bf16 P~ tiles) for the backward
        return ops.relattn_fwd(qkv[:, :d], kvm[:, :d], kvm[:, d:], qkv[:, d:2 * d], qkv[:, 2 * d:], r, rwb, rrb, B, T, H, dh, band, save=True)
    out, lse, saved = fwd()

    def bwd():
        ops.relattn_bwd(qkv[:, :d], kvm[:, :d], kvm[:, d:], qkv[:, d:2 * d], qkv[:, 2 * d:], r, rwb, rrb, out, lse, dout, dqkv[:, :d], dkvm[:, :d],
                        dkvm[:, d:], dqkv[:, d:2 * d], dqkv[:, 2 * d:], dr, drwb, drrb, B, T, H, dh, band, saved=saved)

    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    ms_fwd = timeit(fwd)
    ms_bwd_all = timeit(bwd)
    lib = importlib.import_module(PKG + '._lib').load()
    lib.txl_relattn_bwd_probe(48, 0)      # dQ pass alone (prep kernel and lite passes skipped; operands in the workspace are stale but valid)
    try:
        ms_dq = timeit(bwd)
    finally:
        lib.txl_relattn_bwd_probe(0, 0)
    Kb = min(M, T + M)
    unit = 2.0 * T * Kb * d * B                      # one score-sized contraction over the live band, per launch
    tf = lambda units, ms: units * unit / (ms / 1e3) / 1e12
    dq = tf(3, ms_dq)                                # dP, dQw, dQr
    return {'kernel': 'relattn_bwd_dq_saved_kernel (dQ pass of the attention backward: P from the saved P~ tiles, dP, dS, dQw, dQr; writes bf16 dS tiles)',
            'bound': 'tensor', 'achieved': dq, 'peak': peaks['tf_burst'], 'unit': 'TFLOP/s', 'frac': dq / peaks['tf_burst'],
            'traffic': DQ_TRAFFIC, 'traffic_source': DQ_TRAFFIC_SRC,
            'ms_per_launch': ms_dq, 'algorithmic_flops_per_launch': 3 * unit, 'peak_source': 'burst bf16, ' + peaks['src'],
            'other_kernels': {
                'relattn_fwd_tc_kernel (AC+BD+rel_shift+band mask+softmax+PV)': {'ms_per_launch': ms_fwd, 'achieved': tf(3, ms_fwd), 'frac': tf(3, ms_fwd) / peaks['tf_burst'],
                                                                                  'algorithmic_flops_per_launch': 3 * unit, 'traffic': 787.6e6},
                'attention backward, all passes (prep + dQ + lite dK/dV + lite dR)': {'ms_per_call': ms_bwd_all, 'achieved': tf(6, ms_bwd_all),
                                                                                      'frac': tf(6, ms_bwd_all) / peaks['tf_burst'], 'algorithmic_flops_per_call': 6 * unit}}}


def main():
    # everything except the final JSON line goes to stderr (NCCL prints its version banner on stdout)
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--seq', type=int, default=1024)
    ap.add_argument('--mem-len', type=int, default=1024)
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'fp32'])
    ap.add_argument('--dropout', type=float, default=0.1)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-decode', action='store_true')
    ap.add_argument('--no-cfg5', action='store_true')
    ap.add_argument('--cfg5-batch', type=int, default=16)
    ap.add_argument('--decode-only', action='store_true')
    ap.add_argument('--decode-seqs', type=int, default=64)
    ap.add_argument('--decode-new', type=int, default=2048)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
