"""Top stalled SASS instructions from `ncu -i X.ncu-rep --page source --csv` (one launch)."""
import csv
import sys


def main(path, n=45):
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if 'Address' in r and '# Samples' in r)
    isrc, isam, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data = [r for r in rows if len(r) == len(hdr) and r[isam].isdigit()]
    tot = sum(int(r[isam]) for r in data)
    print('instructions', len(data), 'total samples', tot)
    top = sorted(range(len(data)), key=lambda k: -int(data[k][isam]))[:n]
    for k in sorted(top):
        r = data[k]
        st = sorted(((int(r[i]), h[6:]) for i, h in stall_cols), reverse=True)[:2]
        print(f'{k:5d} {r[isrc].strip()[:72]:72s} {r[isam]:>7s} {r[iex]:>9s} {st}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
