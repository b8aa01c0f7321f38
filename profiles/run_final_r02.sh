#!/bin/bash
# Round-2 final validation on one B200: full GPU test suite, default bench line, reference arm, ncu --set full of the fused LayerNorm GEMM.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_gpu_tests_final.log; tail -3 gpurun_out/r02_gpu_tests_final.log
timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; cut -c1-400 gpurun_out/r02_bench_final.json
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; cut -c1-300 gpurun_out/r02_bench_reference_arm.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6 > gpurun_out/r02_smoke.log; cat gpurun_out/r02_smoke.log
timeout 600 ncu --set full --clock-control none -k regex:tc_gemm_kernel -s 6 -c 2 -f -o gpurun_out/r02_gemm_ln python profiles/gemm_ncu_probe.py > gpurun_out/r02_ncu_gemm_ln.log 2>&1
ncu -i gpurun_out/r02_gemm_ln.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r02_gemm_ln_raw.csv
python profiles/summarize_ncu_full.py gpurun_out/r02_gemm_ln_raw.csv > gpurun_out/r02_ncu_full_gemm_ln.txt; cat gpurun_out/r02_ncu_full_gemm_ln.txt
