#!/usr/bin/env python3
"""SASS opcode evidence per kernel of the built library (no GPU needed): instruction counts and the mnemonics that prove
the Blackwell paths (UBLKCP = TMA bulk copy, SYNCS = mbarrier, FFMA2 / FADD2 / FMUL2 = packed fp32, FMNMX3 = 3-input
min/max, DFMA = fp64).   usage: sass_summary.py [lib.so] > profiles/rNN_sass_summary.txt"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "tactics2d_b200/libt2d_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FMNMX3", "FFMA", "DFMA", "MUFU", "F2I", "FRND", "LDS", "STS", "LDG", "STG", "LDL", "STL",
        "ATOMS", "BAR", "CALL"]
name = None
n = collections.Counter()
c = collections.defaultdict(collections.Counter)
for line in txt.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        n[name] += 1
        c[name][m.group(2)] += 1
print(f"# {so}: cuobjdump -sass, per kernel (callees included): instructions | " + " ".join(KEYS))
for f in sorted(n):
    d = subprocess.run(["c++filt", f], capture_output=True, text=True).stdout.strip() or f
    print(f"{d[:78]:78s} {n[f]:6d} | " + " ".join(f"{k}={c[f][k]}" for k in KEYS if c[f][k]))
