#!/bin/bash
# Round-2 hygiene pass, third part: the attention kernels after the per-row soft-max reference change (saving forward, prep kernel writing the
# scaled dO, dQ pass without the P tile store, dK/dV pass over the forward's tiles).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "(test_relattn_fwd_bwd and bf16 and (128-128-128-1024-1 or 192-64-64 or 64-192-192)) or (wide_score and 128-128-128 and 12.0)" > gpurun_out/r02c_sanitizer_memcheck.log 2>&1
tail -4 gpurun_out/r02c_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 0 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "test_relattn_fwd_bwd and bf16 and True and 128-128-128-1024-1" > gpurun_out/r02c_sanitizer_racecheck.log 2>&1
tail -6 gpurun_out/r02c_sanitizer_racecheck.log
