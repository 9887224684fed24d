"""Aggregate an ncu `--page source --csv --print-source cuda,sass` dump by kernel phase (samples + instructions).

usage: python profiles/phase_breakdown.py dump.csv path/to/t2d_kernels.cu path/to/t2d_math.cuh
Phases are found from the `// ----... name` section comments of the kernel source and the function headers.
"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    out, fpath, hdr = [], None, None
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
                out.append((fpath, int(r[0]), int(r[iN]), int(r[iS])))
            except ValueError:
                pass
    return out


def marks(src_path, pats):
    ms = []
    for i, line in enumerate(open(src_path).read().split("\n")):
        t = line.strip()
        for p in pats:
            m = re.match(p, t)
            if m:
                ms.append((i + 1, m.group(1)[:48]))
                break
    return ms


def region(ms, ln):
    cur = "(top)"
    for a, name in ms:
        if a <= ln:
            cur = name
        else:
            break
    return cur


def main(dump, kern, math):
    out = load(dump)
    km = marks(kern, [r"// -{20,} (.*)", r"__device__ __\w+__ \w[\w ]*?(\w+)\(", r"__global__.*?(t2d_\w+)\("])
    mm = marks(math, [r"T2D_HD \w[\w ]*?(\w+)\("])
    agg_n, agg_s = collections.Counter(), collections.Counter()
    for f, ln, n, s in out:
        if f == kern.split("/")[-1]:
            key = "K: " + region(km, ln)
        elif f == math.split("/")[-1]:
            key = "M: " + region(mm, ln)
        else:
            key = "lib: " + str(f)
        agg_n[key] += n
        agg_s[key] += s
    tn, ts = sum(agg_n.values()) or 1, sum(agg_s.values()) or 1
    print(f"{'phase':55s} {'instr':>9s} {'%':>6s} {'samples':>8s} {'%':>6s}")
    for k, s in agg_s.most_common(30):
        print(f"{k:55s} {agg_n[k]:9d} {100 * agg_n[k] / tn:6.1f} {s:8d} {100 * s / ts:6.1f}")
    print(f"{'total':55s} {tn:9d} {'':6s} {ts:8d}")


if __name__ == "__main__":
    main(*sys.argv[1:4])
