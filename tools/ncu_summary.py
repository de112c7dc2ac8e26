#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/*.md: one block per profiled launch with
the metrics the rooflines are built from.  Usage: ncu_summary.py <report.ncu-rep> <out.md> [title]"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
title = sys.argv[3] if len(sys.argv) > 3 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
]
STALL = "smsp__average_warps_issue_stalled_"
with open(out, "w") as fh:
    fh.write(f"# {title}\n\nSource: `{rep}` (ncu --set full --clock-control none), summarised by tools/ncu_summary.py.\n"
             "Per-launch numbers under ncu are cold-cache and serialised: use them for shares and counters, not as bench values.\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        fh.write(f"\n## {d.get('Kernel Name', '?')}\n\n| metric | value | unit |\n|---|---|---|\n")
        for k in KEYS:
            if k in d and d[k] != "":
                fh.write(f"| {k} | {d[k]} | {u[k]} |\n")
        stalls = sorted(((float(d[h]), h[len(STALL):].replace('_per_issue_active.ratio', '')) for h in hdr
                         if h.startswith(STALL) and d[h] not in ("", "n/a")), reverse=True)
        try:  # executed FP64 flop/s from the SASS op counters (2*dfma + dmul + dadd)
            f = lambda n: float(d[f"smsp__sass_thread_inst_executed_op_{n}_pred_on.sum.per_cycle_elapsed"].replace(",", ""))
            hz = float(d["sm__cycles_elapsed.avg.per_second"].replace(",", "")) * {"Ghz": 1e9, "Mhz": 1e6, "hz": 1.0}.get(
                u["sm__cycles_elapsed.avg.per_second"], 1e9)
            per_cycle = 2 * f("dfma") + f("dmul") + f("dadd")
            lanes = f("dfma") + f("dmul") + f("dadd")
            fh.write(f"\nexecuted FP64 work from the op counters: {per_cycle:.1f} flop/cycle = {per_cycle * hz / 1e12:.2f} TFLOP/s "
                     f"(2*dfma + dmul + dadd); {lanes:.1f} FP64 thread-instructions/cycle of the GPU's "
                     f"148 SM x 64 lanes = {100 * lanes / (148 * 64):.1f} % of the FP64 lanes\n")
        except (KeyError, ValueError):
            pass
        fh.write("\nwarp stall reasons (warps per issue-active cycle): " +
                 ", ".join(f"{n} {v:.2f}" for v, n in stalls[:7]) + "\n")
print("wrote", out)
