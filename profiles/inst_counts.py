#!/usr/bin/env python
"""Per-SASS-instruction execution counts and stall samples of an `ncu --set full --import-source on` capture (first kernel).
    python profiles/inst_counts.py x.ncu-rep > x_inst.txt"""
import csv, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
hdr, rows, name = None, [], None
for r in csv.reader(txt.splitlines()):
    if r and r[0] == "Kernel Name":
        if name is not None:
            break
        name = r[1]; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr is not None and len(r) > 10:
        rows.append(r)
ix = {n: i for i, n in enumerate(hdr)}
col_exec = next(n for n in hdr if n.startswith("# Instructions Executed") or n == "Instructions Executed")
col_thr = next((n for n in hdr if "Thread Instructions Executed" in n), None)
print("# kernel:", name)
print("# columns: index, warp-level instructions executed, thread-level, stall samples, SASS")
tot = sum(int(r[ix[col_exec]]) for r in rows)
print("# total warp instructions:", tot)
for i, r in enumerate(rows):
    print(f"{i:5d} {int(r[ix[col_exec]]):9d} {int(r[ix[col_thr]]) if col_thr else 0:10d} {int(r[ix['# Samples']]):5d}  {r[ix['Source']].strip()[:110]}")
