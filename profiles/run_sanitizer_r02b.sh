#!/bin/bash
# Round-2 hygiene pass, second part: the kernels added after run_sanitizer_r02.sh — the fused Linear + dropout + residual + LayerNorm GEMM
# epilogue (TMEM store / reload, staging buffers, named barriers) and the radix-selection sampler for large vocabularies.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "gemm_add_ln_fused or (sampler_large and 40000 and (8-1.0 or 0-0.9 or 50-0.5))" > gpurun_out/r02b_sanitizer_memcheck.log 2>&1
tail -4 gpurun_out/r02b_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 0 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "(gemm_add_ln_fused and (512-512-512 or 300-256-192)) or (sampler_large and 40000 and 50-0.5)" > gpurun_out/r02b_sanitizer_racecheck.log 2>&1
tail -6 gpurun_out/r02b_sanitizer_racecheck.log
