#!/usr/bin/env python
"""Per-source-line summary of an ncu report (run in the build container, no GPU needed).

    python scripts/ncu_lines.py gpurun_out/prof.ncu-rep salsa_fused_kernelIdLi4 [--top 40] [--by func]

Joins `ncu --page source --csv` (SASS view: executed instructions and stall samples per instruction)
with `nvdisasm -gi` line information of the in-tree shared library, and aggregates by the OUTERMOST
source line in the kernel body (default), by innermost line (--by inner) or by innermost file
(--by file).
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def disassemble(so_path, kernel_substr):
    tmp = tempfile.mkdtemp(prefix='cub_')
    subprocess.run(['cuobjdump', '-xelf', 'all', so_path], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    lines = []
    for cubin in sorted(f for f in os.listdir(tmp) if 'sm_100' in f):
        txt = subprocess.run(['nvdisasm', '-gi', '-c', os.path.join(tmp, cubin)], check=True, stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL).stdout.decode()
        lines += txt.split('\n')
    out = {}          # offset -> (inner (file, line), outer (file, line), sass)
    in_kernel = False
    pending = []
    for ln in lines:
        if ln.startswith('//-----') and '.text.' in ln:
            in_kernel = kernel_substr in ln
            pending = []
            continue
        if not in_kernel:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
        if m:
            pending.append(m.groups())
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m:
            off = int(m.group(1), 16)
            if pending:
                inner = (os.path.basename(pending[0][0]), int(pending[0][1]))
                last = pending[-1]
                outer = (os.path.basename(last[0]), int(last[1]))
                cur = (inner, outer)
                pending = []
            out[off] = (cur[0], cur[1], m.group(2).strip())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('kernel', help='substring of the mangled kernel name, e.g. salsa_fused_kernelIdLi4')
    ap.add_argument('--so', default=os.path.join(ROOT, 'salsa_b200', 'libsalsa_b200.so'))
    ap.add_argument('--top', type=int, default=40)
    ap.add_argument('--by', default='outer', choices=['outer', 'inner', 'file', 'op'])
    ap.add_argument('--outer-line', type=int, default=None, help='only instructions whose outermost line is this')
    args = ap.parse_args()
    dis = disassemble(args.so, args.kernel)
    raw = subprocess.run(['ncu', '-i', args.report, '--page', 'source', '--csv'], check=True, stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL).stdout.decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
    hdr = rows[hdr_i]
    col = {n: i for i, n in enumerate(hdr)}
    base = None
    agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    tot_inst = tot_samp = tot_thr = 0
    stall_cols = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr):
            continue
        addr = int(r[0], 16)
        if base is None:
            base = addr
        off = addr - base
        inst, thr, samp = int(r[col['Instructions Executed']]), int(r[col['Thread Instructions Executed']]), int(r[col['# Samples']])
        info = dis.get(off)
        if args.outer_line is not None and (info is None or info[1][1] != args.outer_line):
            continue
        if info is None:
            key = ('?', 0)
        elif args.by == 'outer':
            key = info[1]
        elif args.by == 'inner':
            key = info[0]
        elif args.by == 'file':
            key = (info[0][0], 0)
        else:
            key = (info[2].split()[0] if not info[2].startswith('@') else info[2].split()[1], 0)
        a = agg[key]
        a[0] += inst
        a[1] += thr
        a[2] += samp
        for n in stall_cols:
            v = int(r[col[n]])
            if v:
                a[3][n] += v
        tot_inst += inst
        tot_thr += thr
        tot_samp += samp
    print('total warp instructions {:,}  thread instr {:,} (avg {:.1f} threads)  samples {:,}'.format(
        tot_inst, tot_thr, tot_thr / max(1, tot_inst), tot_samp))
    print('{:34s} {:>8s} {:>8s} {:>6s}  top stalls'.format('where', 'inst %', 'samp %', 'thr'))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:args.top]:
        st = ', '.join('{} {:.0f}%'.format(n.replace('stall_', ''), 100.0 * v / max(1, a[2])) for n, v in a[3].most_common(3))
        print('{:34s} {:8.2f} {:8.2f} {:6.1f}  {}'.format('{}:{}'.format(*key), 100.0 * a[0] / tot_inst, 100.0 * a[2] / tot_samp,
                                                        a[1] / max(1, a[0]), st))


if __name__ == '__main__':
    sys.exit(main())
