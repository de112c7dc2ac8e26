#!/usr/bin/env python
"""List the FP64 instructions of a kernel's hot loops that read THREE distinct vector registers not served by the
operand-reuse cache (the ones that cost extra FP64-pipe cycles, tools/ubench/fp64_banks3.cu), in program order.
Usage: sass_three.py <lib.so> <kernel-substring> [--all]   (--all prints every FP64 instruction, 3-reads marked)"""
import re
import subprocess
import sys

lib, pat = sys.argv[1], sys.argv[2]
show_all = "--all" in sys.argv
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
    print(name[:110])
    for addr, text in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", text)
        if not (m and int(m.group(1), 16) < addr):
            continue
        tgt = int(m.group(1), 16)
        body = [(a, t) for a, t in ins if tgt <= a <= addr]
        if sum(1 for _, t in body if re.match(r"D(FMA|MUL|ADD)", t)) < 50:
            continue
        print(f" loop 0x{tgt:x}..0x{addr:x}")
        prev = {}
        for a, t in body:
            m2 = re.match(r"(DFMA|DMUL|DADD)\s+(R\d+),\s*(.*)", t)
            if not m2:
                if show_all:
                    print(f"      {a:05x}  {t}")
                continue
            ops = [o.strip() for o in m2.group(3).split(",")]
            regs, keep = [], {}
            for slot, o in enumerate(ops):
                r = re.match(r"[-|]*\|?(R\d+)(\.reuse)?", o)
                if r and not o.lstrip("-|").startswith("RZ"):
                    if prev.get(slot) != r.group(1):
                        regs.append(r.group(1))
                    if r.group(2):
                        keep[slot] = r.group(1)
            prev = keep
            three = len(set(regs)) >= 3
            if three or show_all:
                print(f"  {'3R' if three else '  '}  {a:05x}  {t}")
