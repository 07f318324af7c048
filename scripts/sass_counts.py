#!/usr/bin/env python
"""Static SASS counts of the hot kernels of the built library -> profiles/r2_sass_counts.md (run here, no GPU needed).

    python scripts/sass_counts.py
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sourmash_b200", "libsourmash_b200.so")
KERNELS = [("hash_kmers_kernel<21>", r"hash_kmers_kernelILi21ELb0"), ("hash_kmers_kernel<31>", r"hash_kmers_kernelILi31ELb0"),
           ("hash_kmers_kernel<51>", r"hash_kmers_kernelILi51ELb0"), ("hash_kmers_fused_kernel (21+31+51)", r"hash_kmers_fused_kernel"),
           ("join_stripe_kernel<u16, upper, 2 CTAs per SM>", r"join_stripe_kernelItLb1ELi2E"),
           ("join_stripe_kernel<u16, upper, 1 CTA per SM>", r"join_stripe_kernelItLb1ELi1E"),
           ("one_vs_many_range_major_kernel", r"one_vs_many_range_major_kernel")]
OPS = ("IMAD", "LOP3", "SHF", "IADD3", "LDG", "LDS", "ATOMS")


def symbols():
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and name:
            regs[name] = int(m.group(1))
    return regs


def sass(sym):
    out = subprocess.run(["cuobjdump", "-sass", "-fun", sym, LIB], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    return ins


def main():
    regs = symbols()
    rows = []
    for label, pat in KERNELS:
        syms = [s for s in regs if re.search(pat, s)]
        if not syms:
            rows.append((label, None))
            continue
        sym = sorted(syms, key=len)[0]
        ins = sass(sym)
        addr = [a for a, _ in ins]
        ops = collections.Counter(re.sub(r"^@!?U?P\w+\s+", "", t).split()[0].split(".")[0] for _, t in ins)
        loop = 0
        for a, t in ins:
            m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                loop = max(loop, sum(1 for x in addr if int(m.group(1), 16) <= x <= a))
        rows.append((label, (len(ins), loop, [ops.get(o, 0) for o in OPS], regs[sym])))
    path = os.path.join(ROOT, "profiles", "r2_sass_counts.md")
    with open(path) as fh:
        head = fh.read().split("| kernel |")[0]
    with open(path, "w") as fh:
        fh.write(head)
        fh.write("| kernel | SASS instructions | largest loop body | " + " | ".join(OPS) + " | registers |\n")
        fh.write("|---|---:|---:|" + "---:|" * (len(OPS) + 1) + "\n")
        for label, r in rows:
            if r:
                fh.write("| `%s` | %d | %d | %s | %d |\n" % (label, r[0], r[1], " | ".join(map(str, r[2])), r[3]))
    print(open(path).read())


if __name__ == "__main__":
    main()
