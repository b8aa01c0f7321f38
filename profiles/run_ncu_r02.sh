#!/bin/bash
# Round-2 ncu evidence (one B200): launch list of a training step, launch list of one decode step per engine, ncu --set full of the cluster
# decode kernel (DRAM bytes per launch) and of the attention kernels.  Numbers printed by a run under ncu are never bench values.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# training step: skip the 3 warm-up steps (~360 launches each), take ~1.3 steps
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1100 -c 480 --csv --log-file gpurun_out/r02_launches_train_step.csv \
  python bench.py --steps 2 --warmup 3 --no-decode --no-cpu-baseline --no-cfg5 > gpurun_out/r02_ncu_train.log 2>&1
python profiles/summarize_launches.py gpurun_out/r02_launches_train_step.csv 40 > gpurun_out/r02_launches_train_step_summary.txt
# decode, 8 sequences (cluster engine) and 64 (launch chain): no CUDA graph so that ncu sees every kernel
for b in 8 64; do
  PROBE_B=$b PROBE_NEW=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_decode_b$b.csv \
    python profiles/decode_probe.py > gpurun_out/r02_ncu_decode_b$b.log 2>&1
  python profiles/summarize_launches.py gpurun_out/r02_launches_decode_b$b.csv 30 > gpurun_out/r02_launches_decode_b${b}_summary.txt
done
# full capture of the cluster decode kernel (3rd launch: steady state) and of the launch chain's attention kernel
PROBE_B=8 PROBE_NEW=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_cluster -s 2 -c 1 -f -o gpurun_out/r02_decode_cluster \
  python profiles/decode_probe.py > gpurun_out/r02_ncu_full_cluster.log 2>&1
ncu -i gpurun_out/r02_decode_cluster.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r02_decode_cluster_raw.csv
python profiles/summarize_ncu_full.py gpurun_out/r02_decode_cluster_raw.csv > gpurun_out/r02_ncu_full_decode_cluster.txt
cat gpurun_out/r02_ncu_full_decode_cluster.txt
