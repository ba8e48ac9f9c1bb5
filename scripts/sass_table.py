#!/usr/bin/env python
"""Static SASS table of every kernel in libsalsa_b200.so (cuobjdump -sass, no GPU needed): instruction count and the opcodes that
prove which hardware units a kernel drives -- UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), UTMALDG / UTMASTG
(TMA tensor loads / stores), UBLKCP (bulk copy), HMMA (mma.sync), DFMA/DADD/DMUL (float64), FFMA2 (packed fp32), REDG / RED (atomics).
Usage: python scripts/sass_table.py [path to .so] > profiles/<round>_sass_opcodes.md"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else 'salsa_b200/libsalsa_b200.so'
out = subprocess.run(['cuobjdump', '-sass', so], stdout=subprocess.PIPE, check=True).stdout.decode()
arch = sorted(set(re.findall(r'arch = (sm_\w+)', out)))
funcs = re.split(r'\n\s*Function : ', out)[1:]
COLS = ['UTCHMMA', 'LDTM', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'HMMA', 'DFMA', 'DADD', 'DMUL', 'FFMA2', 'FFMA', 'SHFL', 'LDS', 'STS', 'REDG', 'ATOMG', 'BAR']
print('Static SASS of `{}` (architectures: {}); one row per kernel, counts of static instructions.\n'.format(so, ', '.join(arch)))
print('| kernel | instr | ' + ' | '.join(COLS) + ' |')
print('|---|---|' + '---|' * len(COLS))
rows = []
for f in funcs:
    name = f.split('\n', 1)[0].strip()
    dem = subprocess.run(['c++filt', name], stdout=subprocess.PIPE).stdout.decode().strip()
    dem = re.sub(r'\(.*$', '', dem).replace('salsa::crnn::', 'crnn::').replace('salsa::', '').replace('void ', '')
    hist = collections.Counter()
    n = 0
    for m in re.finditer(r'/\*[0-9a-f]{4,5}\*/\s+(.*?);', f):
        t = re.sub(r'^@!?U?P\d+\s+', '', m.group(1).strip())
        op = t.split()[0].split('.')[0]
        hist[op] += 1
        n += 1
    rows.append((dem, n, hist))
for dem, n, hist in sorted(rows):
    print('| `{}` | {} | '.format(dem, n) + ' | '.join(str(hist.get(c, 0)) if hist.get(c, 0) else '' for c in COLS) + ' |')
