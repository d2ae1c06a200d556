"""Turn the captures of scripts/capture_evidence.sh (gpurun_out/ev_*) into the committed summaries under profiles/:
rNN_launches_bench_cfg2.csv / .summary.txt, rNN_ncu_full_top_cfg2.txt and traffic.json (read by bench.py).
python scripts/extract_evidence.py r02"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r02'
out = os.path.join(ROOT, 'profiles')
src = os.path.join(ROOT, 'gpurun_out')
shutil.copyfile(os.path.join(src, 'ev_launches.csv'), os.path.join(out, f'{tag}_launches_bench_cfg2.csv'))
summ = subprocess.run([sys.executable, os.path.join(out, 'summarize_launches.py'), os.path.join(out, f'{tag}_launches_bench_cfg2.csv')],
                      capture_output=True, text=True).stdout
open(os.path.join(out, f'{tag}_launches_bench_cfg2.summary.txt'), 'w').write(summ)
print(summ)
raw = subprocess.run(['ncu', '-i', os.path.join(src, 'ev_full.ncu-rep'), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
keep = ['dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__time_duration.sum',
        'launch__grid_size', 'launch__registers_per_thread', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active']
units = rows[1]
seen, text, traffic = set(), [], {}
key_of = {'round1_packed_kernel': 'detect.round1', 'round1_kernel': 'detect.round1', 'sample_planes_operands_kernel': 'sample',
          'sample_planes_kernel': 'sample', 'nn_top2_kernel': 'match.search',
          'prep_kernel': 'match.prep', 'sparse_kernel': 'detect.resolve'}
for r in rows[2:]:
    rec = dict(zip(hdr, r))
    name = rec.get('Kernel Name', '')
    short = next((k for k in key_of if k in name), None)
    if short is None or key_of[short] in seen:
        continue
    seen.add(key_of[short])
    text.append('---')
    text.append(f'  Kernel Name {name}')
    for k in keep:
        if k in rec:
            text.append(f'  {k} {rec[k]} {units[hdr.index(k)]}')
    def val(k):
        v, u = float(rec[k].replace(',', '')), units[hdr.index(k)]
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    rd, wr = val('dram__bytes_read.sum'), val('dram__bytes_write.sum')
    traffic[key_of[short]] = {'kernel': short, 'workload': 'cfg2-superpoint256-mha-480x640', 'maps': 128, 'dram_read_bytes': rd,
                              'dram_write_bytes': wr, 'dram_bytes_per_launch': rd + wr,
                              'sm__pipe_tensor_cycles_active_pct_of_peak_elapsed': float(rec['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']),
                              'smsp__inst_executed': float(rec['smsp__inst_executed.sum'].replace(',', '')),
                              'source': f'ncu --set full of bench.py --steps 2 --warmup 3 --cpu-pairs 0 --no-e2e --no-graph --no-others ({tag}_ncu_full_top_cfg2.txt)'}
open(os.path.join(out, f'{tag}_ncu_full_top_cfg2.txt'), 'w').write('\n'.join(text) + '\n')
json.dump(traffic, open(os.path.join(out, 'traffic.json'), 'w'), indent=1)
print('\n'.join(text))
