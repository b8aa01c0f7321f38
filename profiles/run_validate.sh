#!/bin/bash
# Round-end validation on one B200: full GPU test suite, the bench line (both arms), launch list + ncu --set full of the decode-step kernels.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_r01.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r01.log
tail -4 gpurun_out/gpu_tests_r01.log
timeout 600 python bench.py > gpurun_out/bench_r01_gen2.json 2> gpurun_out/bench_r01_gen2.err; tail -c 3000 gpurun_out/bench_r01_gen2.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_gen2_reference.json 2> gpurun_out/bench_r01_gen2_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/decode_launches_gen2.csv python profiles/decode_probe.py > gpurun_out/decode_probe_gen2.log 2>&1
python profiles/summarize_launches.py gpurun_out/decode_launches_gen2.csv > gpurun_out/decode_launches_gen2_summary.txt
