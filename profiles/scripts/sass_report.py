#!/usr/bin/env python
"""Per-kernel SASS opcode histogram and the ptxas register / spill / shared-memory table of the product library.
usage: python profiles/scripts/sass_report.py [out_dir]      (no GPU needed: cuobjdump reads the sm_100a cubin)"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
LIB = os.path.join(REPO, "cosmo-cutout_b200", "lib", "libcosmo_cutout_b200.so")
LOG = os.path.join(REPO, "cosmo-cutout_b200", "build", "ptxas.log")
out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "profiles", "r02")


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n


sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
hist, order, cur = {}, [], None
arch = re.search(r"arch = (sm_\w+)", sass)
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1))
        hist[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
with open(os.path.join(out_dir, "sass_opcodes.txt"), "w") as f:
    f.write(f"# cuobjdump -sass {os.path.relpath(LIB, REPO)}  ({arch.group(1) if arch else '?'}); static instruction counts per kernel\n")
    f.write("# bulk-copy engine: UBLKCP (cp.async.bulk), async copies: LDGSTS (cp.async); tensor pipe: none (no contraction on this path)\n")
    for k in order:
        c = hist[k]
        tot = sum(c.values())
        short = re.sub(r"\(.*", "", k)
        f.write(f"\n== {short}: {tot} instructions\n")
        f.write("   " + "  ".join(f"{op}:{n}" for op, n in c.most_common(28)) + "\n")
        flags = [op for op in ("UBLKCP", "LDGSTS", "SYNCS", "ATOMG", "ATOMS", "RED", "STL", "LDL", "DFMA", "DMUL", "MUFU", "VOTE", "MATCH", "HMMA", "UTCHMMA")
                 if c.get(op)]
        f.write("   notable: " + ", ".join(f"{op} x{c[op]}" for op in flags) + "\n")

rows, cur = [], None
for line in open(LOG):
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = dict(name=re.sub(r"\(.*", "", demangle(m.group(1))), regs="?", spill="0/0", smem="0", stack="0")
        rows.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and cur.get("_seen_stack") is None:
        cur["stack"], cur["spill"] = m.group(1), f"{m.group(2)}/{m.group(3)}"
        cur["_seen_stack"] = True
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = m.group(1)
        s = re.search(r"(\d+) bytes smem", line)
        cur["smem"] = s.group(1) if s else "0"
with open(os.path.join(out_dir, "ptxas_table.txt"), "w") as f:
    f.write("# nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xptxas -v  (cosmo-cutout_b200/build/ptxas.log)\n")
    f.write(f"{'kernel':70s} {'regs':>5s} {'spill st/ld B':>14s} {'static smem B':>14s} {'stack B':>8s}\n")
    for r in rows:
        f.write(f"{r['name'][:70]:70s} {r['regs']:>5s} {r['spill']:>14s} {r['smem']:>14s} {r['stack']:>8s}\n")
print("wrote", os.path.join(out_dir, "sass_opcodes.txt"), "and ptxas_table.txt;", len(order), "kernels")
