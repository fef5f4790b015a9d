#!/usr/bin/env python
"""Where the warp-stall samples of an `ncu --set full --import-source on` capture sit: top instructions per kernel.
    python profiles/stalls.py gpurun_out/x.ncu-rep [topN]"""
import csv, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
ks, cur = [], None
for r in csv.reader(txt.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; ks.append(cur); continue
    if r and r[0] == "Address":
        cur["hdr"] = r; continue
    if cur is not None and len(r) > 10:
        cur["rows"].append(r)
seen = set()
for k in ks:
    if k["name"] in seen:
        continue
    seen.add(k["name"])
    ix = {n: i for i, n in enumerate(k["hdr"])}
    rows = k["rows"]
    tot = sum(int(r[ix["# Samples"]]) for r in rows)
    print(f"== {k['name']}  ({len(rows)} SASS instructions, {tot} samples)")
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i][ix["# Samples"]]))[:top]
    for i in sorted(order):
        r = rows[i]
        st = {n[6:]: int(r[ix[n]]) for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n and int(r[ix[n]]) > 0}
        main = sorted(st.items(), key=lambda x: -x[1])[:2]
        prev = rows[i - 1][ix["Source"]].strip()[:60] if i else ""
        print(f"  #{i:4d} {100 * int(r[ix['# Samples']]) / tot:5.1f}%  {r[ix['Source']].strip()[:70]:70s} {main}   <- prev: {prev}")
