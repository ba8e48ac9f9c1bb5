#!/usr/bin/env python
"""Static SASS of one kernel of libsalsa_b200.so: opcode histogram of an address range (default: the innermost loop that ends
with the last backward branch).  Usage: sass_loop.py <substring of the mangled kernel name> [lo hi]"""
import collections
import re
import subprocess
import sys

so = 'salsa_b200/libsalsa_b200.so'
pat = sys.argv[1]
out = subprocess.run(['cuobjdump', '-sass', so], stdout=subprocess.PIPE, check=True).stdout.decode()
funcs = re.split(r'\n\s*Function : ', out)
sel = [f for f in funcs if f.split('\n', 1)[0].find(pat) >= 0]
assert len(sel) == 1, [f.split('\n', 1)[0] for f in sel]
body = sel[0]
print(body.split('\n', 1)[0])
ins = []
for m in re.finditer(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', body):
    ins.append((int(m.group(1), 16), m.group(2).strip()))
print('instructions', len(ins))
if len(sys.argv) >= 4:
    lo, hi = int(sys.argv[2], 16), int(sys.argv[3], 16)
else:
    back = [(a, t) for a, t in ins if re.search(r'\bBRA\b', t) and int(re.search(r'0x([0-9a-f]+)', t).group(1), 16) < a]
    a, t = back[-2] if len(back) > 1 and 'BRA 0x' in back[-1][1] and back[-1][0] == int(re.search(r'0x([0-9a-f]+)', back[-1][1]).group(1), 16) else back[-1]
    lo, hi = int(re.search(r'0x([0-9a-f]+)', t).group(1), 16), a
print('range %x..%x' % (lo, hi))
hist = collections.Counter()
for a, t in ins:
    if lo <= a <= hi:
        t = re.sub(r'^@!?U?P\d\s+', '', t)
        op = t.split()[0]
        hist[op.split('.')[0] if not op.startswith(('LDS', 'STS', 'LDG', 'STG', 'SHFL', 'F2F', 'MUFU')) else op] += 1
tot = sum(hist.values())
print('in range', tot)
for k, v in hist.most_common():
    print('%5d %s' % (v, k))
