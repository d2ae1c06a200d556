"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (B200_PROFILING.md)."""
import collections
import csv
import io
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(io.StringIO(''.join(lines))):
        name = row['Kernel Name'].split('(')[0][-48:]
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
        agg.setdefault(name, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f'{"kernel":50s} {"n":>4s} {"avg_us":>10s} {"total_us":>10s} {"share":>7s}')
    for k, v in agg.items():
        print(f'{k:50s} {len(v):4d} {sum(v) / len(v):10.1f} {sum(v):10.1f} {100 * sum(v) / tot:6.1f}%')


if __name__ == '__main__':
    main(sys.argv[1])
