#!/usr/bin/env python3
"""Aggregate an ncu report's per-line metrics (needs -lineinfo + --import-source on).
usage: ncu_hot.py report.ncu-rep [top] [function-substring]"""
import csv, subprocess, sys, io, os
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40; want = sys.argv[3] if len(sys.argv) > 3 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = cur_fn = ""; hdr = None; data = []; seen_fn = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = os.path.basename(r[1]); continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        if cur_fn not in seen_fn: seen_fn.append(cur_fn)
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0].isdigit():
        if want and want not in cur_fn: continue
        if len(seen_fn) > 1 and not want and cur_fn != seen_fn[0]: continue
        i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        try: data.append((int(r[i_i]), int(r[i_s] or 0), int(r[i_t]), cur_file, int(r[0]), r[1].strip()[:105]))
        except ValueError: pass
ti = sum(d[0] for d in data); ts = sum(d[1] for d in data); tt = sum(d[2] for d in data)
print("functions:", [f[:70] for f in seen_fn])
print(f"total warp-inst {ti}  samples {ts}  avg active threads/inst {tt / max(1, ti):.1f}")
for d in sorted(data, reverse=True)[:top]:
    print(f"{100 * d[0] / ti:5.1f}% inst {100 * d[1] / max(1, ts):5.1f}% smp thr={d[2] / max(1, d[0]):4.1f} {d[3]}:{d[4]:<4d}| {d[5]}")
