#!/usr/bin/env python
"""Hot spots of an ncu source-page CSV (ncu -i rep --page source --csv): samples per SASS instruction.
usage: ncu_source_hot.py file.csv [top=40]   -- also prints cumulative samples per 64-instruction window"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = rows[2:]
base = int(body[0][ia], 16)
tot = sum(int(r[isamp]) for r in body)
print("total samples", tot, "instructions", len(body))
recs = []
for k, r in enumerate(body):
    st = {hdr[i][6:]: int(r[i]) for i in stall_cols if r[i] not in ("", "0")}
    recs.append((int(r[isamp]), int(r[ia], 16) - base, r[isrc].strip(), int(r[iex]), st))
print("--- windows of 64 instructions: offset, samples, share, executed(first)")
for k in range(0, len(recs), 64):
    s = sum(x[0] for x in recs[k:k + 64])
    if s * 200 >= tot:
        print(f"  {recs[k][1]:#7x} {s:7d} {100.0 * s / tot:5.1f}%  exec {recs[k][3]}")
print("--- top instructions")
for s, off, src, ex, st in sorted(recs, reverse=True)[:top]:
    main = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"  {off:#7x} {s:6d} {100.0 * s / tot:4.1f}% exec {ex:8d}  {src[:60]:60s} {main}")
