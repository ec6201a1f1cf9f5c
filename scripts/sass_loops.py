#!/usr/bin/env python3
"""Instruction mix of the innermost loops of one kernel in an object file (cuobjdump -sass): for every backward
branch, the opcodes between its target and itself.  Usage: sass_loops.py <object> <mangled-kernel-substring>"""
import re
import subprocess
import sys
from collections import Counter

obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, body = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        body[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        body[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in body.items():
    if pat not in name:
        continue
    print(name[:60], len(ins), "instructions")
    for i, (addr, text) in enumerate(ins):
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", text)
        if not m or "BRA" not in text:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= addr:
            continue
        loop = [t for a, t in ins if tgt <= a <= addr]
        if len(loop) < 40:
            continue
        c = Counter()
        for t in loop:
            w = t.split()
            op = w[1] if w[0].startswith("@") else w[0]
            c[op.split(".")[0]] += 1
        fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD"))
        print(f"  loop {tgt:#x}..{addr:#x}: {len(loop)} instr, {fp64} FP64, {len(loop) - fp64} other ->", dict(c.most_common()))
