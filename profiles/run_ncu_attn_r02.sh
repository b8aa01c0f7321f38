#!/bin/bash
# ncu --set full of the round-2 attention kernels (third iteration of profiles/attn_probe.py: forward saving P~ with the per-row reference,
# prep, dQ pass over the saved tiles writing dS only, paired dK/dV over P~ and scaled dO, paired dR), cfg2 shape.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:relattn -s 10 -c 5 -f -o gpurun_out/r02_attn python profiles/attn_probe.py > gpurun_out/r02_ncu_attn.log 2>&1
ncu -i gpurun_out/r02_attn.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r02_attn_raw.csv
python profiles/summarize_ncu_full.py gpurun_out/r02_attn_raw.csv > gpurun_out/r02_ncu_full_attention.txt
cat gpurun_out/r02_ncu_full_attention.txt | grep -E "Kernel Name|gpu__time|dram__bytes|tensor_cycles|dram_throughput"
