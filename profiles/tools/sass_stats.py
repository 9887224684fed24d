#!/usr/bin/env python3
"""Static SASS statistics of one kernel of a built .so (no GPU needed).

usage: sass_stats.py lib.so [kernel-substring] [--dump out.txt]
Prints, for the kernel's main body and for each out-of-line callee: instruction count, local-memory (STL/LDL) count,
an opcode histogram of the main body and the instruction count per source region (needs -lineinfo).
"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    so = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "t2d_step_kernelILi4ELb1"
    dump = sys.argv[sys.argv.index("--dump") + 1] if "--dump" in sys.argv else None
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    lines = txt.split("\n")
    start = end = None
    for i, l in enumerate(lines):
        if start is None and l.startswith("_Z") and pat in l and l.rstrip().endswith(":"):
            start = i
        elif start is not None and l.startswith("//---------------------"):
            end = i
            break
    assert start is not None, "kernel not found"
    cur, fn = None, "main"
    per_fn, local_fn, ops, per_line = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    out = []
    for l in lines[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            c = (m.group(1).split("/")[-1], int(m.group(2)))
            if c != cur:
                out.append("## %s:%d" % c)
                cur = c
            continue
        if "$" in l and l.rstrip().endswith(":"):
            fn = l.strip().split("$")[-1].rstrip(":")[:48]
            out.append(l)
        elif l.startswith(".L_"):
            out.append(l)
        m3 = re.match(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", l)
        if m3:
            op = m3.group(3).split(".")[0]
            per_fn[fn] += 1
            if op in ("STL", "LDL"):
                local_fn[fn] += 1
            if fn == "main":
                ops[op] += 1
                per_line[cur] += 1
            out.append("  %s  %s%s%s" % (m3.group(1), m3.group(2) or "", m3.group(3), m3.group(4)))
    if dump:
        open(dump, "w").write("\n".join(out))
    print("kernel:", lines[start].rstrip(":"))
    for k, v in per_fn.most_common():
        print(f"  {v:6d} instr  {local_fn[k]:4d} STL/LDL  {k}")
    print("main opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common(28)))
    reg = collections.Counter()
    for (f, ln), v in per_line.items():
        reg[(f, ln // 20 * 20)] += v
    print("main by source (20-line bins):")
    for (f, ln), v in sorted(reg.items(), key=lambda kv: -kv[1])[:40]:
        print(f"  {v:5d}  {f}:{ln}")


if __name__ == "__main__":
    main()
