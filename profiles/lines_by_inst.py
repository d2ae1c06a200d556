"""Executed warp instructions and stall samples per CUDA source line from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` (one launch).  python lines_by_inst.py file.csv [min_pct]"""
import csv
import os
import sys


def main(path, min_pct=0.4):
    rows = list(csv.reader(open(path)))
    cur = None
    hdr = None
    out = []
    for r in rows:
        if len(r) >= 2 and r[0] == 'File Path':
            cur = os.path.basename(r[1])
        elif len(r) > 5 and r[0] == 'Line No':
            hdr = r
        elif hdr and len(r) == len(hdr) and r[0].isdigit():
            out.append((cur, r))
    iex, isam = hdr.index('Instructions Executed'), hdr.index('# Samples')
    tot = sum(int(r[iex]) for _, r in out)
    tsam = sum(int(r[isam]) for _, r in out)
    print('warp instructions', tot, 'samples', tsam)
    for f, r in out:
        if int(r[iex]) > tot * min_pct / 100 or int(r[isam]) > tsam * min_pct / 100:
            print(f'{f[:20]:20s} {r[0]:>4s} {100 * int(r[iex]) / tot:6.2f}% inst {100 * int(r[isam]) / tsam:6.2f}% samples  {r[1].strip()[:100]}')


if __name__ == '__main__':
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.4)
