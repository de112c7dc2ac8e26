#!/usr/bin/env python
"""Print the hottest loop (largest backward branch span containing DFMA) of a kernel's SASS,
with an opcode histogram.  Usage: tools/sass_loop.py <lib.so> <kernel-name-substring> [--full]"""
import re
import subprocess
import sys
from collections import Counter

lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
for b in blocks:
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
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr:
                body = [t for a, t in ins if tgt <= a <= addr]
                nd = sum(1 for t in body if re.match(r"(@!?U?P\d\s+)?D(FMA|MUL|ADD)", t))
                if best is None or nd > best[0]:
                    best = (nd, tgt, addr, body)
    print("==", name[:120])
    if best:
        nd, tgt, addr, body = best
        hist = Counter(re.sub(r"^@!?U?P\d\s+", "", t).split()[0].split(".")[0] for t in body)
        print(f"loop 0x{tgt:x}..0x{addr:x}: {len(body)} instructions, {nd} FP64")
        print("  ".join(f"{k}:{v}" for k, v in hist.most_common()))
        if "--full" in sys.argv:
            for t in body:
                print("   ", t)
