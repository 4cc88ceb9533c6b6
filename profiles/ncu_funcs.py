"""Aggregate an ncu source page by source-line RANGES (functions) given on the command line or built in.

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-id ::regex:NAME:1 > src.csv
    python profiles/ncu_funcs.py src.csv
Prints warp-instructions, share, lane efficiency (thread-instructions / 32 / warp-instructions) and stall samples.
"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
cur = hdr = None
byline = collections.defaultdict(lambda: [0, 0, 0])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split('/')[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0].strip().isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        try:
            inst = int(d.get("Instructions Executed", "0") or 0)
            tinst = int(d.get("Thread Instructions Executed", "0") or 0)
            samp = int(d.get("# Samples", "0") or 0)
        except ValueError:
            continue
        b = byline[(cur, int(r[0]))]
        b[0] += inst; b[1] += tinst; b[2] += samp

# function ranges from the sources themselves: a line that starts a PKB_* / template function at column 0
import os
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'pokerl_b200', 'csrc')
ranges = []
for f in ('pkb_game.cuh', 'pkb_philox.cuh', 'pkb_eval.cuh', 'pkb_common.cuh', 'pkb_kernels.cu'):
    lines = open(os.path.join(root, f)).read().split('\n')
    starts = []
    for i, l in enumerate(lines, 1):
        m = re.match(r'^(PKB_\w+|__device__|__global__|static|template).*?(\w+)\s*\(', l)
        if l.startswith(('PKB_', '__device__', '__global__')) or (l.startswith('    PKB_HDM') ):
            m2 = re.search(r'(\w+)\s*\(', l.replace('__launch_bounds__(', 'LB('))
            name = m2.group(1) if m2 else '?'
            if name in ('LB',):
                m3 = re.search(r'\)\s*(\w+)\s*\(', l)
                name = m3.group(1) if m3 else name
            starts.append((i, name))
    for j, (i, name) in enumerate(starts):
        end = starts[j + 1][0] - 1 if j + 1 < len(starts) else len(lines)
        ranges.append((f, i, end, name))
tot = sum(b[0] for b in byline.values()); tots = sum(b[2] for b in byline.values())
acc = collections.defaultdict(lambda: [0, 0, 0])
for (f, l), b in byline.items():
    name = f
    for (rf, a, e, n) in ranges:
        if rf == f and a <= l <= e:
            name = f"{f}:{n}"
            break
    for k in range(3):
        acc[name][k] += b[k]
print(f"total warp-instructions {tot}, samples {tots}")
for name, b in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    if b[0] == 0 and b[2] == 0:
        continue
    eff = b[1] / 32 / b[0] if b[0] else 0
    print(f"{name:45s} {b[0]:>10} {100*b[0]/tot:5.1f}%  lanes {eff*32:5.1f}  samples {b[2]:>6} {100*b[2]/max(tots,1):5.1f}%")
if len(sys.argv) > 2:  # per-line listing of one file:function range
    f, a, e = sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    for (ff, l), b in sorted(byline.items()):
        if ff == f and a <= l <= e and b[0]:
            print(f"{ff}:{l:4d} {b[0]:>9} lanes {b[1]/b[0]:5.1f} samp {b[2]}")
