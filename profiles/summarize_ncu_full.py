"""Summarise `ncu -i X.ncu-rep --page raw --csv` (one row per captured launch): the metrics DESIGN.md / bench.py quote, plus the top stall reasons."""
import csv
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__block_size', 'launch__grid_size', 'launch__cluster_size', 'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f'{w:75s} {r[i][:100]} {units[i]}')
    st = sorted(((float(r[i]), hdr[i]) for i in range(len(hdr))
                 if 'smsp__average_warps_issue_stalled' in hdr[i] and hdr[i].endswith('_per_issue_active.ratio') and r[i]), reverse=True)
    print('  top warp stall reasons (per issue-active):', ', '.join(f"{n.split('stalled_')[1].split('_per')[0]}={v:.2f}" for v, n in st[:6]))
    print()
