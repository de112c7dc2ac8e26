#!/usr/bin/env python
"""FP64 instruction mix and estimated FP64-pipe cycles of a kernel's hot loops on sm_100a, read from its SASS, with the
measured rule  cycles(FP64 instr) = max(2, #distinct 64-bit vector-register operands not served by .reuse)
(tools/ubench/fp64_operands.cu, fp64_reuse.cu; at the pair kernel's occupancy an unserved third read costs ~2, not 1,
extra cycles: tools/ubench/fp64_banks3.cu).

    sass_cost.py <lib.so> <kernel-substring> [pairs-per-iteration]

bench.py imports `default_kernel_loops` to put the instruction counts of the LOADED library into the roofline block
(so they cannot go stale), and tests/test_sass_counts.py pins them.
"""
import re
import subprocess
import sys

# fresnel_pairs_kernel<DPT=4, MODE=FAITHFUL, TILE=256, STAGES=3, TBL=4096, MINB=2, FORM=15, SPI=2, THREADS=256>
DEFAULT_KERNEL = "fresnel_pairs_kernelILi4ELi0ELi256ELi3ELi4096ELi2ELi15ELi2ELi256E"
DEFAULT_PAIRS = 8  # 4 detector points x 2 sources per loop iteration
# the REFERENCED kernel: MODE=2, FORM=tan|polar|shortcos|wfold|e2|rowt = 95
REFERENCED_KERNEL = "fresnel_pairs_kernelILi4ELi2ELi256ELi3ELi4096ELi2ELi95ELi2ELi256E"


def _functions(lib):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    for b in re.split(r"\n\s*Function : ", txt):
        name = b.split("\n", 1)[0]
        ins = []
        for line in b.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        yield name, ins


def hot_loops(ins, min_fp64=50):
    """Innermost backward-branch loops with at least `min_fp64` FP64 instructions: list of (start, body)."""
    loops = []
    for addr, text in ins:
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", text)
        if not (m and int(m.group(1), 16) < addr):
            continue
        tgt = int(m.group(1), 16)
        body = [t for a, t in ins if tgt <= a <= addr]
        nd = sum(1 for t in body if re.match(r"D(FMA|MUL|ADD)", t))
        inner = True
        for a2, t2 in ins:
            if tgt <= a2 < addr and "BRA" in t2:
                mm = re.search(r"0x([0-9a-f]+)", t2)
                if mm and tgt <= int(mm.group(1), 16) < a2:
                    inner = False
        if nd >= min_fp64 and inner:
            loops.append((tgt, body))
    return loops


def loop_cost(body, pairs):
    cyc = three = other = 0
    ops_n = {"DFMA": 0, "DMUL": 0, "DADD": 0}
    prev_reuse = {}  # slot -> register kept by the previous FP64 instruction of this warp
    for t in body:
        m = re.match(r"(DFMA|DMUL|DADD)\s+(R\d+),\s*(.*)", t)
        if not m:
            other += 1
            continue
        ops_n[m.group(1)] += 1
        regs, keep = [], {}
        for slot, o in enumerate(o.strip() for o in m.group(3).split(",")):
            r = re.match(r"[-|]*\|?(R\d+)(\.reuse)?", o)
            if r and not o.lstrip("-|").startswith("RZ"):
                if prev_reuse.get(slot) != r.group(1):
                    regs.append(r.group(1))
                if r.group(2):
                    keep[slot] = r.group(1)
        prev_reuse = keep
        d = len(set(regs))
        three += d >= 3
        cyc += max(2, d)
    n = sum(ops_n.values())
    return {"fp64_instr_per_pair": n / pairs, "dfma_per_pair": ops_n["DFMA"] / pairs, "dmul_per_pair": ops_n["DMUL"] / pairs,
            "dadd_per_pair": ops_n["DADD"] / pairs, "exec_flop_per_pair": (2 * ops_n["DFMA"] + ops_n["DMUL"] + ops_n["DADD"]) / pairs,
            "three_read_per_pair": three / pairs, "model_cycles_per_pair": cyc / pairs, "other_instr_per_pair": other / pairs}


def kernel_loops(lib, kernel, pairs=DEFAULT_PAIRS):
    """{'planar_row': {...}, 'general': {...}} for the pair kernel whose mangled name contains `kernel`."""
    for name, ins in _functions(lib):
        if kernel in name:
            costs = sorted((loop_cost(body, pairs) for _, body in hot_loops(ins)), key=lambda c: c["fp64_instr_per_pair"])
            if len(costs) != 2:
                raise RuntimeError(f"expected the planar-row and the general loop, found {len(costs)} hot loops")
            return {"kernel": kernel, "planar_row": costs[0], "general": costs[1]}
    raise RuntimeError(f"pair kernel {kernel} not found in {lib}")


def default_kernel_loops(lib):
    """The default FAITHFUL pair kernel of `lib` (what bench.py times)."""
    return kernel_loops(lib, DEFAULT_KERNEL)


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    for name, ins in _functions(lib):
        if pat not in name:
            continue
        print(name[:100])
        for tgt, body in hot_loops(ins):
            c = loop_cost(body, pairs)
            print(f"  loop at 0x{tgt:x}: FP64 instrs {c['fp64_instr_per_pair'] * pairs:.0f} ({c['fp64_instr_per_pair']:.2f}/pair), "
                  f"3-read instrs {c['three_read_per_pair'] * pairs:.0f}, FP64 pipe cycles {c['model_cycles_per_pair'] * pairs:.0f} "
                  f"({c['model_cycles_per_pair']:.1f}/pair), other instrs {c['other_instr_per_pair'] * pairs:.0f} "
                  f"({c['other_instr_per_pair']:.1f}/pair); DFMA:DMUL:DADD = {c['dfma_per_pair']:.2f}:{c['dmul_per_pair']:.2f}:{c['dadd_per_pair']:.2f}")


if __name__ == "__main__":
    main()
