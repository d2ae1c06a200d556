"""SASS evidence for the Blackwell-specific paths: per kernel of csrc/libkb_b200.so the number of tcgen05 MMA
(UTCHMMA, .2CTA = cta_group::2), TMEM load (LDTM), TMA tensor load (UTMALDG), bulk async copy (UBLKCP), cp.async
(LDGSTS) and mbarrier (SYNCS) instructions.  python scripts/sass_evidence.py > profiles/r02_sass_evidence.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'keypoint_bench_b200', 'csrc', 'libkb_b200.so')
PATTERNS = ['UTCHMMA', 'UTCHMMA.2CTA', 'LDTM', 'UTMALDG', 'UBLKCP', 'LDGSTS', 'SYNCS', 'UTCBAR', 'FMNMX3', 'HMNMX2', 'VHMNMX', 'REDUX', 'ATOMS', 'DFMA']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    arch = re.findall(r'arch = (sm_\w+)', sass)
    print(f'cuobjdump -sass {os.path.relpath(LIB, ROOT)}: {len(sass.splitlines())} lines, arch {sorted(set(arch))}')
    print(f'{"kernel":70s} ' + ' '.join(f'{p:>12s}' for p in PATTERNS) + '  instructions')
    cur, counts, order = None, {}, []
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r'\(.*', '', cur.replace('(anonymous namespace)::', ''))
            counts[cur] = dict.fromkeys(PATTERNS, 0) | {'n': 0}
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)', line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]['n'] += 1
        for p in PATTERNS:
            if op == p or op.startswith(p + '.') or (p.count('.') and op.startswith(p)):
                counts[cur][p] += 1
    for k in order:
        c = counts[k]
        print(f'{k[:70]:70s} ' + ' '.join(f'{c[p]:12d}' for p in PATTERNS) + f'  {c["n"]}')


if __name__ == '__main__':
    sys.exit(main())
