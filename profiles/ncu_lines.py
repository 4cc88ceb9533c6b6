"""Aggregate an ncu source page (--page source --csv --print-source cuda,sass) by CUDA source line.

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-id ::regex:NAME:1 > src.csv
    python profiles/ncu_lines.py src.csv [top]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0].strip().isdigit():
        d = dict(zip(hdr[4:], r[4:]))
        try:
            inst = int(d.get("Instructions Executed", "0") or 0)
            samp = int(d.get("# Samples", "0") or 0)
        except ValueError:
            continue
        stalls = {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
        out.append((inst, samp, cur_file, int(r[0]), r[1].strip()[:80], stalls))
ti, ts = sum(o[0] for o in out), sum(o[1] for o in out)
print(f"total warp-instructions {ti}, total samples {ts}")
print("== top lines by instructions executed")
for o in sorted(out, key=lambda x: -x[0])[:top]:
    print(f"{o[0]:>10} {100*o[0]/ti:5.1f}%  samp {o[1]:>6}  {o[2]}:{o[3]}  {o[4]}")
print("== top lines by stall samples")
for o in sorted(out, key=lambda x: -x[1])[:top]:
    big = sorted(o[5].items(), key=lambda kv: -kv[1])[:3]
    print(f"{o[1]:>8} {100*o[1]/ts:5.1f}%  inst {o[0]:>9}  {o[2]}:{o[3]}  {o[4][:60]}  {big}")
agg = {}
for o in out:
    for k, v in o[5].items():
        agg[k] = agg.get(k, 0) + v
print("== stall reasons overall:", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
