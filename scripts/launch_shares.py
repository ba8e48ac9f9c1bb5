#!/usr/bin/env python
"""Kernel shares from an ncu launch list (--metrics gpu__time_duration.sum --csv): python scripts/launch_shares.py launches.csv"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
head, rows = rows[0], rows[1:]
ki, vi = head.index('Kernel Name'), head.index('Metric Value')
tot, cnt = collections.OrderedDict(), collections.Counter()
for r in rows:
    name = re.sub(r'\(.*$', '', r[ki]).replace('void ', '').strip()
    tot[name] = tot.get(name, 0.0) + float(r[vi]) / 1e6
    cnt[name] += 1
s = sum(tot.values())
print('| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|')
for k, v in tot.items():
    print('| {} | {} | {:.3f} | {:.1f} | {:.1f}% |'.format(k, cnt[k], v, v / cnt[k] * 1e3, 100 * v / s))
