"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (B200_PROFILING.md).
Launches are normalised to ONE pipeline step with the known launch counts per step, because the capture
window of bench.py also covers its isolated per-stage timing loops."""
import collections
import csv
import io
import sys

PER_STEP = {'tau_kernel': 1, 'round1_kernel': 1, 'sparse_kernel': 1, 'nms_rounds_kernel': 1, 'select_kernel': 1,
            'warp_homography_kernel': 1, 'sample_planes_kernel': 1, 'sample_kernel': 1, 'prep_kernel': 1,
            'nn_top2_kernel': 1, 'resolve_kernel': 1, 'rescan_kernel': 1, 'gate_kernel': 1, 'pairs_kernel': 1,
            'rep_min_sorted_kernel': 2, 'rep_min_pruned_kernel': 2}


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO(''.join(lines))):
        name = row['Kernel Name'].split('(')[0].split('::')[-1].split('<')[0]
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
        agg.setdefault(name, []).append(v)
    per_step = {k: PER_STEP.get(k, 1) * sum(v) / len(v) for k, v in agg.items()}
    tot = sum(per_step.values())
    print(f'{"kernel":28s} {"captured":>8s} {"avg_us":>9s} {"x/step":>6s} {"us/step":>9s} {"share":>7s}')
    for k, v in agg.items():
        print(f'{k:28s} {len(v):8d} {sum(v) / len(v):9.1f} {PER_STEP.get(k, 1):6d} {per_step[k]:9.1f} {100 * per_step[k] / tot:6.1f}%')
    print(f'{"total per step":28s} {"":8s} {"":9s} {"":6s} {tot:9.1f}')


if __name__ == '__main__':
    main(sys.argv[1])
