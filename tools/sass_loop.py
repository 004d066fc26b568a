#!/usr/bin/env python
"""Opcode histogram of the innermost-but-one loop (the per-row loop) of one
kernel of libplb.so -- static, no GPU:  python tools/sass_loop.py '<2, 2, 3>'"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "pylabolt_b200", "lib", "libplb.so")
want = sys.argv[1] if len(sys.argv) > 1 else "<2, 2, 3>"
if len(sys.argv) > 2:
    LIB = sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
name, ins = None, []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        full = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout
        name = full if ("k_bulk_fused" + want) in full else None
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and name:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
# widest backward branch that is not the work-item loop = the row loop
loops = []
for a, t in ins:
    m = re.search(r"BRA\S*\s+(?:\S+,\s+)?0x([0-9a-f]+)$", t)
    if m and int(m.group(1), 16) < a and ".ANY" not in t:
        loops.append((a - int(m.group(1), 16), int(m.group(1), 16), a))
loops.sort(reverse=True)
span, lo, hi = loops[1] if len(loops) > 1 and loops[0][0] > 1.2 * loops[1][0] else loops[0]
body = [t for a, t in ins if lo <= a <= hi]
c = collections.Counter(re.sub(r"^@!?U?P\d\s+", "", t).split()[0].split(".")[0] for t in body)
fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DADD", "DMUL"))
print(f"k_bulk_fused{want}: row loop {lo:#x}..{hi:#x}, {len(body)} instructions, {fp64} fp64")
print(", ".join(f"{k} {v}" for k, v in c.most_common(40)))
