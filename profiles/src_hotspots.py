"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` dump per CUDA source line.

usage: python profiles/src_hotspots.py dump.csv [top_n]
"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    out = []
    fpath, hdr = None, None
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fpath = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            H = len(hdr)
            iN = hdr.index("Instructions Executed") - H
            iS = hdr.index("# Samples") - H
            continue
        if hdr and r and r[0].strip().isdigit() and len(r) >= H:
            try:
                n, s = int(r[iN]), int(r[iS])
            except ValueError:
                continue
            out.append((n, s, fpath, int(r[0]), ",".join(r[1:len(r) - H + 2]).strip()[:100]))
    tot = sum(a[0] for a in out) or 1
    tots = sum(a[1] for a in out) or 1
    print(f"total warp-instructions {tot}, stall samples {tots}")
    for n, s, f, ln, src in sorted(out, reverse=True)[:top]:
        print(f"{n:10d} {100 * n / tot:5.1f}%  samples {100 * s / tots:5.1f}%  {f}:{ln}: {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
