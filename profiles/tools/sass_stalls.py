#!/usr/bin/env python3
"""Top SASS instructions by stall samples of an `ncu --page source --csv --print-source cuda,sass` dump, each with the
source line it belongs to.  usage: sass_stalls.py dump.csv [min_samples]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
mins = int(sys.argv[2]) if len(sys.argv) > 2 else 8
hdr = None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        H = len(r)
        break
iS = hdr.index("# Samples")
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
cur = None
fpath = None
out = []
tot = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        continue
    if len(r) >= H and r[0].strip().isdigit():
        cur = f"{fpath}:{r[0]}"
    elif len(r) >= H and r[0] == "" and r[2].startswith("0x"):
        off = len(r) - H
        try:
            s = int(r[iS + off])
        except ValueError:
            continue
        tot += s
        if s >= mins:
            st = " ".join(f"{h[6:]}:{r[i + off]}" for i, h in stall if r[i + off] not in ("0", ""))
            out.append((s, r[3].strip()[:64], cur, st))
print("total samples", tot)
for s, ins, src, st in sorted(out, reverse=True):
    print(f"{s:4d} {100 * s / max(tot, 1):5.1f}%  {ins:64s} {src:24s} {st}")
