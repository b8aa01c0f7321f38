#!/bin/bash
# A/B of the decode step on one B200: GPU tests of the second-generation kernels, decode-only bench under the A/B switches, launch list.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "dec_linear or grouped" > gpurun_out/decode_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/decode_tests.log
tail -5 gpurun_out/decode_tests.log
run() { name=$1; shift; envs=$1; shift; env $envs timeout 300 python bench.py --decode-only "$@" > gpurun_out/decode_$name.json 2> gpurun_out/decode_$name.err; python -c "
import json,sys
d=json.load(open('gpurun_out/decode_$name.json')); print('$name', round(d['value']), 'tok/s', round(d['ms_per_token_step']*1e3,1), 'us/step', round(d['roofline']['frac'],4))"; }
run gen2_final A=1
run gen2_g4_s2 "TXL_DECODE_ATTN_SPLITS=2"
run gen2_final_b32 A=1 --decode-seqs 32
run gen2_final_b48 A=1 --decode-seqs 48
run gen2_b32_g1 TXL_DECODE_GROUPS=1 --decode-seqs 32
run gen2_final_b16 A=1 --decode-seqs 16
run gen2_b16_g2 TXL_DECODE_GROUPS=2 --decode-seqs 16
if [ -n "$WITH_NCU" ]; then
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/decode_launches_gen2.csv python profiles/decode_probe.py > gpurun_out/decode_probe_gen2.log 2>&1
python profiles/summarize_launches.py gpurun_out/decode_launches_gen2.csv > gpurun_out/decode_launches_gen2_summary.txt; head -18 gpurun_out/decode_launches_gen2_summary.txt
fi
