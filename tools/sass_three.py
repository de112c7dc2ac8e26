#!/usr/bin/env python
"""List the FP64 instructions of a kernel's hot loop that read three distinct vector registers not served by
.reuse (3 cycles instead of 2 on the FP64 pipe, tools/ubench/fp64_operands.cu).
Usage: sass_three.py <lib.so> <kernel-substring>"""
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for b in re.split(r"\n\s*Function : ", txt):
    name = b.split("\n", 1)[0]
    if pat not in name:
        continue
    ins = []
    for line in b.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    for addr, text in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", text)
        if m and int(m.group(1), 16) < addr:
            tgt = int(m.group(1), 16)
            body = [t for a, t in ins if tgt <= a <= addr]
            nd = sum(1 for t in body if re.match(r"D(FMA|MUL|ADD)", t))
            if best is None or nd > best[0]:
                best = (nd, body)
    prev_reuse, prev_text = {}, ""
    for t in best[1]:
        m = re.match(r"(DFMA|DMUL|DADD)\s+(R\d+),\s*(.*)", t)
        if not m:
            continue
        regs, keep = [], {}
        for slot, o in enumerate(x.strip() for x in m.group(3).split(",")):
            r = re.match(r"[-|]*\|?(R\d+)(\.reuse)?", o)
            if r and not o.lstrip("-|").startswith("RZ"):
                if prev_reuse.get(slot) != r.group(1):
                    regs.append(r.group(1))
                if r.group(2):
                    keep[slot] = r.group(1)
        if len(set(regs)) >= 3:
            print(f"  prev: {prev_text:42s} 3-read: {t}")
        prev_reuse, prev_text = keep, t
