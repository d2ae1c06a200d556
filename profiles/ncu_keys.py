"""Key metrics of one launch from `ncu -i X.ncu-rep --page raw --csv`.  python ncu_keys.py raw.csv"""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
for vals in rows[2:]:
    print('---', vals[hdr.index('Kernel Name')][:90] if 'Kernel Name' in hdr else '')
    for i, h in enumerate(hdr):
        if h in KEYS or ('issue_stalled' in h and 'per_issue_active' in h and float(vals[i] or 0) > 0.2):
            print(f'{h:90s} {vals[i]}')
