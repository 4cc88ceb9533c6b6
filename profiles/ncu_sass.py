"""Per-SASS-instruction view of an ncu source page (--page source --csv --print-source cuda,sass):
sorts instructions by address, prints cumulative samples/instructions per address bucket, and the
hottest instructions, so that code regions (e.g. warp roles) can be told apart.

    python profiles/ncu_sass.py src.csv [bucket_count]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
cur = None
ins = {}
for r in rows:
    if not r:
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] in ("File Path", "Function Name"):
        continue
    if hdr is None:
        continue
    if r[0].strip().isdigit():
        cur = (int(r[0]), r[1].strip()[:70])
        continue
    addr = r[2].strip() if len(r) > 2 else ""
    if addr.startswith("0x"):
        d = dict(zip(hdr[4:], r[4:]))
        try:
            samp = int(d.get("# Samples", "0") or 0)
            inst = int(d.get("Instructions Executed", "0") or 0)
        except ValueError:
            continue
        a = int(addr, 16)
        if a not in ins:
            ins[a] = [r[3].strip(), samp, inst, cur]
addrs = sorted(ins)
base = addrs[0]
tot_s = sum(v[1] for v in ins.values())
tot_i = sum(v[2] for v in ins.values())
print(f"{len(addrs)} SASS instructions, {(addrs[-1]-base)//1024} KB of code, samples {tot_s}, warp-instructions {tot_i}")
per = (len(addrs) + nb - 1) // nb
for k in range(0, len(addrs), per):
    chunk = addrs[k:k + per]
    s = sum(ins[a][1] for a in chunk)
    i = sum(ins[a][2] for a in chunk)
    top = max(chunk, key=lambda a: ins[a][1])
    print(f"+{chunk[0]-base:6d}..+{chunk[-1]-base:6d}  samples {s:5d} ({100*s/tot_s:4.1f}%)  inst {i:9d} ({100*i/tot_i:4.1f}%)  hot: {ins[top][0][:40]:40s} {ins[top][1]:4d} @ {ins[top][3]}")
