#!/usr/bin/env python3
"""Aggregate `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per source line: executed warp-instructions,
stall samples and the dominant stall reasons.  usage: src_regions.py dump.csv [top_n] [--bins N]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 60
    bins = int(sys.argv[sys.argv.index("--bins") + 1]) if "--bins" in sys.argv else 1
    rows = list(csv.reader(open(path)))
    fpath, hdr = None, None
    agg = {}
    stall_tot = collections.Counter()
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            H = len(hdr)
            iN, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
            stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
            continue
        if hdr and r and r[0].strip().isdigit() and len(r) >= H:
            off = len(r) - H   # commas inside the source text shift the columns
            try:
                n, s = int(r[iN + off]), int(r[iS + off])
            except ValueError:
                continue
            key = (fpath, int(r[0]) // bins * bins)
            a = agg.setdefault(key, [0, 0, collections.Counter(), ""])
            a[0] += n
            a[1] += s
            for i, h in stall_cols:
                try:
                    v = int(r[i + off])
                except ValueError:
                    v = 0
                if v:
                    a[2][h[6:]] += v
                    stall_tot[h[6:]] += v
            if not a[3]:
                a[3] = ",".join(r[1:2 + off]).strip()[:70]
    tn = sum(a[0] for a in agg.values()) or 1
    ts = sum(a[1] for a in agg.values()) or 1
    print(f"total warp-instructions {tn}, samples {ts}")
    print("stalls overall:", ", ".join(f"{k} {100 * v / ts:.1f}%" for k, v in stall_tot.most_common(10)))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        st = " ".join(f"{k}:{v}" for k, v in a[2].most_common(3))
        print(f"{a[0]:9d} {100 * a[0] / tn:5.1f}%  smp {a[1]:5d} {100 * a[1] / ts:5.1f}%  {key[0]}:{key[1]}  [{st}]  {a[3]}")


if __name__ == "__main__":
    main()
