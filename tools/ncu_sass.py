#!/usr/bin/env python3
"""Top SASS instructions by stall samples. usage: ncu_sass.py report.ncu-rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
hdr = rows[hi]
end = len(rows)
for i in range(hi + 1, len(rows)):
    if rows[i] and rows[i][0] == "Kernel Name": end = i; break
blk = [r for r in rows[hi + 1:end] if len(r) >= len(hdr)]
ci, cs, ct = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Avg. Threads Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[cs] or 0) for r in blk)
agg = {}
for i, h in stall_cols:
    agg[h] = sum(int(r[i] or 0) for r in blk)
print("stall totals:", ", ".join(f"{k[6:]} {100*v/max(1,sum(agg.values())):.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
opc = {}
for r in blk:
    op = r[1].split()[0] if r[1].split() else "?"
    if op.startswith("@"): op = r[1].split()[1]
    op = op.split(".")[0]
    o = opc.setdefault(op, [0, 0]); o[0] += int(r[ci] or 0); o[1] += int(r[cs] or 0)
ti = sum(v[0] for v in opc.values())
print("opcode mix (warp inst):", ", ".join(f"{k} {100*v[0]/ti:.1f}%" for k, v in sorted(opc.items(), key=lambda kv: -kv[1][0])[:18]))
print("opcode stall samples:", ", ".join(f"{k} {100*v[1]/max(1,tot):.1f}%" for k, v in sorted(opc.items(), key=lambda kv: -kv[1][1])[:14]))
for r in sorted(blk, key=lambda r: -int(r[cs] or 0))[:top]:
    st = sorted(((int(r[i] or 0), h[6:]) for i, h in stall_cols), reverse=True)[:3]
    print(f"{100*int(r[cs] or 0)/max(1,tot):5.2f}% smp inst={int(r[ci]):>9d} thr={r[ct]:>4s} {r[1].strip()[:60]:60s} {st}")
