#!/usr/bin/env python
"""Opcode histogram per kernel of libdecaf_b200.so (cuobjdump -sass), the evidence that the hot kernels use tcgen05 / TMEM /
TMA (UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store,
SYNCS = mbarrier, UTCATOMSWS = tcgen05.alloc) and where the others sit (HMMA = mma.sync, LDGSTS = cp.async, MUFU, ...).

    python tools/sass_summary.py > profiles/r02_sass_summary.txt        (no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'cvpr2025-decafnet_b200', 'decaf_b200', 'libdecaf_b200.so')
KEY = ['UTCHMMA', 'UTCBAR', 'LDTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UTCATOMSWS', 'SYNCS', 'HMMA', 'LDGSTS', 'LDG', 'STG', 'LDS', 'STS',
       'MUFU', 'FFMA2', 'FMUL2', 'FADD2', 'FFMA', 'SHFL', 'BAR', 'ATOMG', 'ATOMS', 'RED', 'ELECT', 'R2UR', 'LDL', 'STL']


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    if not os.path.exists(lib):
        for cand in (os.path.join(ROOT, 'cvpr2025-decafnet_b200', 'csrc', 'libdecaf_b200.so'), ):
            if os.path.exists(cand):
                lib = cand
    txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)', line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    dem = subprocess.run(['cu++filt'] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f'# cuobjdump -sass {os.path.relpath(lib, ROOT)}: {len(kernels)} kernels (sm_100a), instruction counts per kernel')
    print('# total | ' + ' '.join(KEY))
    rows = []
    for (mangled, cnt), name in zip(kernels.items(), dem if len(dem) == len(kernels) else list(kernels)):
        name = re.sub(r'^void ', '', name).replace('decaf::', '')
        depth = 0
        for i, ch in enumerate(name):            # cut the parameter list: the first '(' outside the template arguments
            if ch == '<':
                depth += 1
            elif ch == '>':
                depth -= 1
            elif ch == '(' and depth == 0:
                name = name[:i]
                break
        name = re.sub(r'\((?:int|bool)\)', '', name)
        rows.append((name, sum(cnt.values()), cnt))
    for name, total, cnt in sorted(rows, key=lambda r: r[0]):
        ks = ' '.join(f'{k}={cnt[k]}' for k in KEY if cnt.get(k))
        print(f'{name[:90]:<90} {total:6d} | {ks}')


if __name__ == '__main__':
    main()
