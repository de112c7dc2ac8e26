#!/usr/bin/env python
"""Headline benchmark: Fresnel terms/s of the Huygens-Fresnel pair sum on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path over one synthetic batch, inputs resident in HBM:
  mirror chain trace of 1000x1000 rays (fused chain kernel) -> calc_dS -> exp(-ik OPL)
  -> Fresnel pair sum of the 1e6 last-mirror points onto the focal grid.
Every N runs BASELINE config C4 as written -- the configuration the metric "Fresnel terms/s at 1/2/4/8 B200" is quoted
on: AKB four-mirror trace, 1e6 rays x a FIXED 2048x2048 detector = 4.2e12 terms per step, STRONG scaling.  Every
rank passes the full grid to the reference-shaped multi-GPU call forward_propagation_cupy_batch_multi_gpu
(-> fresnel_sum_sharded -> akb_fresnel_sum_sharded for N > 1), which computes the rank's array_split block and
all-gathers the blocks in place over NCCL.  Before timing, every rank checks the (gathered) field against the CPU
oracle on 256 random detector points, its bit-equality across ranks, and an uneven case (M not divisible by N); the
outcome is the `parity` block of the JSON line.  At N = 1 the line also carries C3 (KB two-mirror trace, 1e6 rays ->
512x512 grid, the round-1 headline) and the rooflines of the other kernels as secondary blocks.

Prints ONE JSON line (rank 0).  `value` = terms of the whole job / max-over-ranks device time.
`e2e` = the same stage through the reference-facing call with HOST (NumPy) buffers: H2D + kernels (+ all-gather)
+ D2H inside the timed region.  `--impl reference` times the CPU restatement of the reference's numba path (oracle/,
all host threads) on a bounded detector subset of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RAYS = 1000           # ray grid side -> 1e6 source points
WAVELENGTH = 13.5e-9  # CPU0402:243
ALG_FLOP_PER_TERM = 23.0   # SURVEY.md 8(d) convention


def workload(n_gpus):
    """(geometry tag, mirrors, focal grid side) of the configuration this run measures: C4 at every N."""
    return ("c4", 4, 2048)


def workload_config(n_gpus):
    tag, K, G = workload(n_gpus)
    if n_gpus == 1:
        name = ("C4: AKB four-mirror trace of 1e6 rays + Fresnel sum onto a fixed 2048x2048 detector, 1 GPU "
                "(forward_propagation_cupy_batch_multi_gpu on one device)")
    else:
        name = (f"C4: AKB four-mirror trace of 1e6 rays + Fresnel sum onto a fixed 2048x2048 detector, array_split over "
                f"{n_gpus} GPUs through forward_propagation_cupy_batch_multi_gpu (akb_fresnel_sum_sharded: block kernel + "
                f"in-place NCCL all-gather), strong scaling")
    return {"workload": name, "rays": RAYS * RAYS, "mirrors": K, "detector_points": G * G,
            "terms_per_step": float(RAYS * RAYS) * G * G, "wavelength_m": WAVELENGTH,
            "phase_mode": "faithful", "l2": "flushed between timed steps (256 MiB write)"}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------ reference arm

def cpu_workload(n_gpus, n_rays=RAYS):
    """The same source set and focal grid, built WITHOUT the CUDA library: oracle chain trace + oracle calc_dS."""
    import oracle
    tag, K, G = workload(n_gpus)
    geo = np.load(os.path.join(ROOT, "akbraytracing_b200", "data", "geometry.npz"))
    tan_h, tan_v = geo[f"{tag}__tan_h"], geo[f"{tag}__tan_v"]
    n = n_rays
    raw = np.vstack([np.ones(n * n), np.tile(tan_h, n), np.repeat(tan_v, n)])
    ray = oracle.normalize_vector(raw)
    src = np.repeat(geo[f"{tag}__source_point"][:, None], n * n, axis=1)
    tr = oracle.trace_chain(list(geo[f"{tag}__coeffs"]), [bool(b) for b in geo[f"{tag}__negative"]], geo[f"{tag}__plane"], ray, src)
    last = tr["points"][-1]
    k = 2.0 * np.pi / WAVELENGTH
    opl = tr["dist"][0]
    for d in tr["dist"][1:]:
        opl = opl + d
    u = np.exp(-1j * (k * opl))
    ds = oracle.calc_dS(last, n, n).ravel()
    det = tr["det"]
    yc, zc = (det[1].min() + det[1].max()) / 2, (det[2].min() + det[2].max()) / 2
    yy, zz = np.meshgrid(np.linspace(yc - 1e-6, yc + 1e-6, G), np.linspace(zc - 1e-6, zc + 1e-6, G))
    return dict(det_x=np.full(G * G, det[0].mean()), det_y=yy.ravel(), det_z=zz.ravel(),
                src_x=np.ascontiguousarray(last[0]), src_y=np.ascontiguousarray(last[1]),
                src_z=np.ascontiguousarray(last[2]), u=u, ds=ds, k=k)


def host_threads():
    """Every host core this process may run on.  torch.distributed.run exports OMP_NUM_THREADS=1 to its
    workers; the CPU arm must not inherit that (it would time ONE thread), so the count is passed explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def time_cpu_sample(w, n_det, threads, rng_seed=0):
    import oracle
    sel = np.sort(np.random.default_rng(rng_seed).choice(w["det_x"].shape[0], n_det, replace=False))
    t0 = time.perf_counter()
    oracle.fresnel_sum(w["det_x"][sel], w["det_y"][sel], w["det_z"][sel], w["src_x"], w["src_y"], w["src_z"],
                       w["u"], w["k"], w["ds"], nthreads=threads)
    dt = time.perf_counter() - t0
    return n_det * w["src_x"].shape[0] / dt, dt


def cpu_baseline(target_s=12.0):
    """Oracle port (C + OpenMP, all host threads) on a bounded detector subset of the C4 stage."""
    import oracle
    oracle.build()
    threads = host_threads()
    w = cpu_workload(1)
    total = w["det_x"].shape[0]
    time_cpu_sample(w, max(threads, 16), threads)            # warms the thread pool
    rate, _ = time_cpu_sample(w, 4 * max(threads, 16), threads)  # calibration
    n_det = int(min(total, max(threads, rate * target_s / w["src_x"].shape[0])))
    n_det = max(threads, (n_det // threads) * threads)
    rate, dt = time_cpu_sample(w, n_det, threads)
    return {"value": rate, "unit": "terms/s", "cores": threads, "kind": "port",
            "sample": f"{n_det} random detector points of the C4 2048x2048 grid x all 1e6 source points "
                      f"({n_det * w['src_x'].shape[0]:.3g} terms, {dt:.1f} s), oracle/akb_oracle.c with OpenMP"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    threads = host_threads()
    w = cpu_workload(args.gpus)
    n_src, total = w["src_x"].shape[0], w["det_x"].shape[0]
    tag, K, G = workload(args.gpus)
    time_cpu_sample(w, max(threads, 16), threads)
    rate, _ = time_cpu_sample(w, 4 * max(threads, 16), threads)
    step_s = float(os.environ.get("AKB_BENCH_REF_STEP_S", "4.0"))  # CPU seconds per step (tests shorten it)
    n_det = int(max(threads, rate * step_s / n_src))
    n_det = min(total, max(threads, (n_det // threads) * threads))
    for i in range(args.warmup):
        time_cpu_sample(w, n_det, threads, rng_seed=100 + i)
    t = 0.0
    for i in range(args.steps):
        _, dt = time_cpu_sample(w, n_det, threads, rng_seed=i)
        t += dt
    value = args.steps * n_det * n_src / t
    sample = (f"per step: {n_det} random detector points of the {tag.upper()} {G}x{G} grid x all {n_src} source points "
              f"({n_det * n_src:.3g} terms), CPU restatement of CPU0402:71-124 (oracle/akb_oracle.c, OpenMP)")
    print(json.dumps({
        "impl": "reference", "metric": "fresnel_terms_per_s", "value": value, "unit": "terms/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "terms/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "terms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------ our arm

def sass_costs():
    """FP64 instruction mix per pair of the loops the default pair kernel runs, read from the loaded library's SASS
    (tools/sass_cost.py) so that the roofline's instruction counts cannot go stale; None without cuobjdump."""
    try:
        from tools import sass_cost
        from akbraytracing_b200 import _lib
        return sass_cost.default_kernel_loops(_lib.LIB_PATH)
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


def ncu_traffic(key):
    """dram bytes per launch of the pair kernel on workload `key` ('c4_n1': the N = 1 bench launch) from this round's
    committed ncu capture (profiles/r02_ncu_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")) as fh:
            return json.load(fh).get(key)
    except (OSError, ValueError):
        return None


def run_ours(args):
    import ctypes
    import torch
    import torch.distributed as dist
    from akbraytracing_b200 import build as akb_build
    akb_build.ensure_built()  # no-op when the in-tree library is current; ranks serialise on a file lock
    import akbraytracing_b200 as akb
    from akbraytracing_b200 import handoff, raytrace, workloads, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: library banners printed while the communicator comes up
        # (NCCL_DEBUG=VERSION on some boxes) are sent to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    L = _lib.load()
    tag, K, GRID = workload(world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident inputs: ray bundle, full focal grid (every rank holds it, like the reference's x, y, z)
    coeffs, neg, plane, ray, src = workloads.chain_inputs(tag, RAYS, dev)
    k = 2.0 * np.pi / WAVELENGTH
    tr0 = raytrace.trace_chain(coeffs, neg, plane, ray, src)
    gx, gy, gz = workloads.focal_grid(tr0["det"], GRID)
    G_total = GRID * GRID
    begin, count = _lib.shard_range(G_total, world, rank)
    sl = slice(begin, begin + count)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def sources():
        tr = raytrace.trace_chain(coeffs, neg, plane, ray, src, check=False)
        last = tr["points"][-1]
        opl = tr["dist"][0]
        for q in range(1, K):
            opl = opl + tr["dist"][q]
        u = handoff.opl_to_field(opl, k)
        ds = handoff.calc_dS(last, RAYS, RAYS).reshape(-1)
        return last, u, ds

    one_device = [local]  # at N = 1 the multi-GPU call is told to stay on this process's device

    def multi(*a):
        """The reference-shaped multi-GPU call: full arrays in, full field out (on every rank)."""
        if world > 1:
            return akb.forward_propagation_cupy_batch_multi_gpu(*a)
        return akb.forward_propagation_cupy_batch_multi_gpu(*a, devices=one_device)

    def step():
        last, u, ds = sources()
        return multi(gx, gy, gz, last[0], last[1], last[2], u, k, ds)

    # ---- parity before timing (every rank)
    parity = field_parity(akb, torch, dist, step(), sources, (gx, gy, gz), k, world, rank, dev, multi)

    L.akb_fresnel_timing(1)
    for _ in range(args.warmup):
        step()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    pair_ms = []
    _lib.launch_count(reset=True)
    with ClockSampler(local) as clocks:
        barrier()
        for s in range(args.steps):
            flush.fill_(s & 0xFF)  # evict L2 between timed steps (not timed)
            ev[s][0].record()
            out = step()
            ev[s][1].record()
            p, t = ctypes.c_double(), ctypes.c_double()
            sp, bx, ps = ctypes.c_int(), ctypes.c_int64(), ctypes.c_int()
            _lib.check(L.akb_fresnel_last_timing(ctypes.byref(p), ctypes.byref(t), ctypes.byref(sp), ctypes.byref(bx),
                                                 ctypes.byref(ps)), "akb_fresnel_last_timing")
            pair_ms.append(p.value)
        barrier()
    launches = _lib.launch_count()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    pair_mean = torch.tensor([float(np.mean(pair_ms))], dtype=torch.float64, device=dev)
    nlaunch = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(pair_mean, op=dist.ReduceOp.MAX)
        dist.all_reduce(nlaunch, op=dist.ReduceOp.SUM)
    total_ms, pair_mean = float(total_ms.item()), float(pair_mean.item())
    terms_step = float(RAYS * RAYS) * G_total
    value = terms_step * args.steps / (total_ms * 1e-3)
    plan = {"source_splits": sp.value, "detector_blocks": bx.value, "resident_blocks_per_sm": ps.value}

    # ---- end to end through the reference-facing call with HOST buffers (H2D + kernels (+ gather) + D2H timed)
    last, u, ds = sources()
    dev_args = [gx, gy, gz, last[0].contiguous(), last[1].contiguous(), last[2].contiguous(), u, ds]

    def host_copy(t, pin):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=pin)
        h.copy_(t)
        return h.numpy()
    api = multi

    def e2e_run(pin, steps):
        host = [host_copy(t, pin) for t in dev_args]
        res = api(*host[:7], k, host[7])  # warm-up
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            res = api(*host[:7], k, host[7])
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        # bytes that cross PCIe on ALL ranks: every rank uploads the whole source set and its own block of the
        # detector coordinates (fresnel_sum_sharded slices host arrays before the upload)
        moved = world * sum(a.nbytes for a in host[3:]) + sum(a.nbytes for a in host[:3])
        return terms_step * steps / float(dt.item()), res, moved
    e2e_steps = max(1, min(args.steps, 10))
    e2e_value, host_out, h2d = e2e_run(True, e2e_steps)
    e2e_pageable, _, _ = e2e_run(False, max(1, min(args.steps, 5)))
    d2h = G_total * 16
    same = float((torch.as_tensor(host_out).to(dev) - out).abs().max())

    result = None
    if rank == 0:
        peaks, peak_src = measured_peaks()
        tf = ctypes.c_double()
        _lib.check(L.akb_fp64_peak_probe(4096, ctypes.byref(tf), None), "akb_fp64_peak_probe")
        fp64_peak = tf.value
        costs = sass_costs()
        row = (costs or {}).get("planar_row") or {}
        instr_per_term, exec_flop = row.get("fp64_instr_per_pair"), row.get("exec_flop_per_pair")
        pair_terms = float(RAYS * RAYS) * count
        rate = pair_terms / (pair_mean * 1e-3)
        achieved = rate * ALG_FLOP_PER_TERM / 1e12
        traffic = ncu_traffic("c4_n1") if world == 1 else None
        roofline = {
            "kernel": "fresnel_pairs_kernel<faithful> [%s], planar-row loop" % L.akb_fresnel_variant_name().decode(),
            "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
            "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
            "traffic_source": traffic["source"] if traffic else "no ncu capture of this launch committed for this round",
            "peak_source": "on-box DFMA microbenchmark (akb_fp64_peak_probe, measured in this run); "
                           "MEASURED_PEAKS.json has no FP64 figure; nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2",
            "peak_nominal": 37.2, "frac_of_nominal": achieved / 37.2,
            "algorithmic_flop_per_term": ALG_FLOP_PER_TERM,
            "sass": costs,  # instruction mix per pair read from the loaded .so (planar-row and general loop)
            "executed_flop_per_term": exec_flop, "fp64_instr_per_term": instr_per_term,
            "achieved_exec": rate * exec_flop / 1e12 if exec_flop else None,
            "frac_exec": rate * exec_flop / 1e12 / fp64_peak if exec_flop else None,
            # issue-slot view of the same pipe: a DFMA-only stream reaches `peak` with one FP64 instruction per
            # 2 cycles; this kernel issues fp64_instr_per_term of them per term
            "frac_pipe_issue": rate * instr_per_term * 2.0 / (fp64_peak * 1e12) if instr_per_term else None,
            "kernel_ms": pair_mean, "kernel_share_of_step": pair_mean * args.steps / total_ms,
            "terms_per_s_kernel": rate, "plan": plan,
        }
        result = {
            "metric": "fresnel_terms_per_s", "value": value, "unit": "terms/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world), "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "terms/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h * world,
                    "api": ("forward_propagation_cupy_batch_multi_gpu with NumPy buffers (one device: akb_fresnel_sum_host, "
                            "H2D + kernels + D2H)" if world == 1 else
                            "forward_propagation_cupy_batch_multi_gpu with NumPy buffers on every rank "
                            "(H2D, akb_fresnel_sum_sharded incl. the NCCL all-gather, D2H of the full field)"),
                    "host_buffers": "pinned", "timer": "host wall clock around the synchronous call, max over ranks",
                    "steps": e2e_steps, "max_abs_diff_vs_device_path": same,
                    "pageable": {"value": e2e_pageable, "unit": "terms/s",
                                 "note": "the same call on ordinary (pageable) NumPy arrays, what a NumPy caller holds"}},
            "gpu_launches": int(nlaunch.item()), "roofline": roofline,
        }
        result["parity"] = parity
    if world == 1 and rank == 0:
        # ---- secondary blocks (N = 1 only): other loops / kernels of the two paths, each with its own roofline
        ref_field = out
        result["c3"] = bench_c3(akb, handoff, raytrace, workloads, torch, dev, L, k)
        result["roofline_m2m"] = bench_m2m(akb, handoff, raytrace, workloads, torch, dev, L, fp64_peak, costs, k)
        g_out, g_rate, g_ms = gpu0402_restatement(torch, (gx, gy, gz), (last[0].contiguous(), last[1].contiguous(),
                                                  last[2].contiguous()), u, ds, k)
        result["gpu_baseline"] = {
            "value": g_rate, "unit": "terms/s", "n_gpus": 1,
            "kind": "restatement of forward_propagation_cupy_batch (GPU0402:64-136) in torch on the same B200; cupy is not installed",
            "sample": f"{g_out.shape[0]} detector points x {last.shape[1]} sources in batches of 128 ({g_ms:.1f} ms, best of 3 passes)",
            "rel_l2_vs_fused_kernel": float(torch.linalg.vector_norm(g_out - ref_field[:g_out.shape[0]]) /
                                            torch.linalg.vector_norm(ref_field[:g_out.shape[0]]))}
        del g_out
        torch.cuda.empty_cache()
        phase_modes = {}
        for mode_name, mode_id in (("exact", akb.PHASE_EXACT), ("referenced", akb.PHASE_REFERENCED)):
            fm = akb.fresnel_sum(gx, gy, gz, last[0], last[1], last[2], u, k, ds, mode=mode_id)  # one pass: no warm-up needed at 7 s
            p_ms = ctypes.c_double()
            L.akb_fresnel_last_timing(ctypes.byref(p_ms), None, None, None, None)
            phase_modes[mode_name] = {
                "terms_per_s": terms_step / (p_ms.value * 1e-3),
                "rel_l2_vs_faithful": float(torch.linalg.vector_norm(fm - ref_field) / torch.linalg.vector_norm(ref_field)),
                "peak_pixel_same_as_faithful": int(fm.abs().argmax()) == int(ref_field.abs().argmax())}
            del fm
        result["phase_modes"] = phase_modes
        result["roofline_ray"] = bench_ray_c2(akb, workloads, torch, dev, peaks, peak_src)
        result["roofline_chain"] = bench_chain(workloads, torch, L, _lib, peaks, peak_src)
        result["small_call_us"] = bench_small_call(akb, workloads)
    if world > 1:
        dist.barrier()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            result["cpu_baseline"] = cpu_baseline()
        print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


def field_parity(akb, torch, dist, field, sources, grid, k, world, rank, dev, multi):
    """Checks of the timed call on every rank: (1) the (gathered) C4 field against the CPU oracle on 256 random
    detector points (different points on every rank, all 1e6 sources), peak pixel of the sample included;
    (2) bit-equality of the gathered field across ranks; (3) an uneven case, M = 4096*world + 3 detector points
    against the first 20000 sources, in full against the oracle.  Returns the `parity` block (max over ranks)."""
    import oracle
    oracle.build()
    gx, gy, gz = grid
    last, u, ds = sources()
    threads = max(1, (os.cpu_count() or 8) // world)
    rng = np.random.default_rng(1000 + rank)
    sel = np.sort(rng.choice(field.shape[0], 256, replace=False))
    st = torch.as_tensor(sel, device=dev)
    h = [t.cpu().numpy() for t in (gx[st], gy[st], gz[st], last[0], last[1], last[2], u, ds)]
    ref = oracle.fresnel_sum(*h[:7], k, h[7], nthreads=threads)
    got = field[st].cpu().numpy()
    rel = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    peak_same = int(np.argmax(np.abs(got))) == int(np.argmax(np.abs(ref)))
    # (2) every rank holds the same bits as rank 0
    ranks_equal = True
    if world > 1:
        mine = field.clone()
        dist.broadcast(field.view(torch.float64) if field.is_cuda else field, src=0)
        ranks_equal = bool(torch.equal(mine, field))
        del mine
    # (3) uneven shards through the same call
    M = 4096 * world + 3
    n_src = 20000
    got_u = multi(gx[:M], gy[:M], gz[:M], last[0][:n_src], last[1][:n_src], last[2][:n_src], u[:n_src], k, ds[:n_src])
    hu = [t.cpu().numpy() for t in (gx[:M], gy[:M], gz[:M], last[0][:n_src], last[1][:n_src], last[2][:n_src], u[:n_src], ds[:n_src])]
    ref_u = oracle.fresnel_sum(*hu[:7], k, hu[7], nthreads=threads)
    rel_u = float(np.linalg.norm(got_u.cpu().numpy() - ref_u) / np.linalg.norm(ref_u))
    agg = torch.tensor([rel, rel_u, 0.0 if peak_same else 1.0, 0.0 if ranks_equal else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
    a = agg.cpu().tolist()
    return {"rel_l2": a[0], "peak_same": a[2] == 0.0, "ranks_equal": a[3] == 0.0 if world > 1 else None, "uneven_rel_l2": a[1],
            "uneven_case": f"M = {M} detector points over {world} ranks x {n_src} sources, checked in full",
            "sample": f"256 random detector points per rank x 1e6 sources vs oracle/akb_oracle.c ({threads} host threads per rank); "
                      f"max over {world} ranks", "gate": 1e-6}


def bench_c3(akb, handoff, raytrace, workloads, torch, dev, L, k, G=512):
    """BASELINE config C3 (KB two-mirror trace of 1e6 rays -> 512x512 focal grid, 2.6e11 terms): the round-1 headline,
    one warm-up + best of 3 of the whole stage (trace, calc_dS, exp(-ik OPL), pair sum), oracle parity on 64 points."""
    import ctypes
    import oracle
    coeffs, neg, plane, ray, src = workloads.chain_inputs("c3", RAYS, dev)
    tr0 = raytrace.trace_chain(coeffs, neg, plane, ray, src)
    gx, gy, gz = workloads.focal_grid(tr0["det"], G)

    def stage():
        tr = raytrace.trace_chain(coeffs, neg, plane, ray, src, check=False)
        last = tr["points"][-1]
        u = handoff.opl_to_field(tr["dist"][0] + tr["dist"][1], k)
        ds = handoff.calc_dS(last, RAYS, RAYS).reshape(-1)
        return akb.fresnel_sum(gx, gy, gz, last[0], last[1], last[2], u, k, ds), (last, u, ds)
    best, best_pair = None, None
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        field, (last, u, ds) = stage()
        e1.record()
        torch.cuda.synchronize()
        p = ctypes.c_double()
        L.akb_fresnel_last_timing(ctypes.byref(p), None, None, None, None)
        if rep:
            best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
            best_pair = p.value if best_pair is None else min(best_pair, p.value)
    terms = float(RAYS * RAYS) * G * G
    sel = np.sort(np.random.default_rng(3).choice(G * G, 64, replace=False))
    st = torch.as_tensor(sel, device=dev)
    h = [t.cpu().numpy() for t in (gx[st], gy[st], gz[st], last[0], last[1], last[2], u, ds)]
    ref = oracle.fresnel_sum(*h[:7], k, h[7])
    got = field[st].cpu().numpy()
    return {"workload": "C3: KB two-mirror trace of 1e6 rays + Fresnel sum onto a 512x512 focal grid, 1 GPU",
            "terms_per_step": terms, "ms_per_step": best, "terms_per_s": terms / (best * 1e-3),
            "pair_kernel_ms": best_pair, "terms_per_s_kernel": terms / (best_pair * 1e-3),
            "rel_l2_vs_oracle": float(np.linalg.norm(got - ref) / np.linalg.norm(ref)),
            "parity_sample": "64 random detector points x 1e6 sources vs oracle/akb_oracle.c",
            "timing": "CUDA events around the whole stage, best of 3 after a warm-up"}


def bench_m2m(akb, handoff, raytrace, workloads, torch, dev, L, fp64_peak, costs, k):
    """The mirror-to-mirror stage of the reference's chain (CPU0402:283-327: N x N terms, three of them per AKB run):
    detector set = a mirror's traced point cloud (irregular: the pair kernel's GENERAL loop), here the first 262144
    points of the AKB chain's 2nd mirror against the 1e6 points of its 1st mirror."""
    import ctypes
    import oracle
    coeffs, neg, plane, ray, src = workloads.chain_inputs("c4", RAYS, dev)
    tr = raytrace.trace_chain(coeffs, neg, plane, ray, src)
    back, front = tr["points"][0], tr["points"][1][:, :512 * 512].contiguous()
    u = handoff.opl_to_field(tr["dist"][0], k)
    ds = handoff.calc_dS(back, RAYS, RAYS).reshape(-1)
    terms = float(back.shape[1]) * front.shape[1]
    out = {}
    fields = {}
    for name, mode in (("faithful", akb.PHASE_FAITHFUL), ("exact", akb.PHASE_EXACT), ("referenced", akb.PHASE_REFERENCED)):
        best = None
        for rep in range(3):
            f = akb.fresnel_sum(front[0], front[1], front[2], back[0], back[1], back[2], u, k, ds, mode=mode)
            p, t = ctypes.c_double(), ctypes.c_double()
            L.akb_fresnel_last_timing(ctypes.byref(p), ctypes.byref(t), None, None, None)
            if rep:
                best = p.value if best is None else min(best, p.value)
        fields[name] = f
        out[name] = {"kernel_ms": best, "terms_per_s": terms / (best * 1e-3)}
    sel = np.sort(np.random.default_rng(7).choice(front.shape[1], 64, replace=False))
    st = torch.as_tensor(sel, device=dev)
    h = [t.cpu().numpy() for t in (front[0][st], front[1][st], front[2][st], back[0], back[1], back[2], u, ds)]
    ref = oracle.fresnel_sum(*h[:7], k, h[7])
    for name in fields:
        got = fields[name][st].cpu().numpy()
        out[name]["rel_l2_vs_oracle"] = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    from akbraytracing_b200.stagechain import auto_phase_mode
    auto = auto_phase_mode(k, front.cpu().numpy(), back.cpu().numpy())
    gen = (costs or {}).get("general") or {}
    rate = out["faithful"]["terms_per_s"]
    achieved = rate * ALG_FLOP_PER_TERM / 1e12
    from akbraytracing_b200 import _lib as _l
    names = {_l.PHASE_FAITHFUL: "faithful", _l.PHASE_EXACT: "exact", _l.PHASE_REFERENCED: "referenced"}
    return {"kernel": "fresnel_pairs_kernel<faithful>, general loop (irregular detector set)",
            "workload": f"AKB mirror 1 (1e6 points) -> first {front.shape[1]} points of mirror 2, {terms:.3g} terms per launch",
            "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
            "traffic": None, "fp64_instr_per_term": gen.get("fp64_instr_per_pair"),
            "executed_flop_per_term": gen.get("exec_flop_per_pair"),
            "frac_exec": rate * gen["exec_flop_per_pair"] / 1e12 / fp64_peak if gen.get("exec_flop_per_pair") else None,
            "modes": out, "auto_phase_mode": names[auto],
            "parity_sample": "64 random detector points x 1e6 sources vs oracle/akb_oracle.c",
            "timing": "CUDA events inside the library around the pair kernel, best of 2 after a warm-up"}


def gpu0402_restatement(torch, det, src, u, ds, k, batch=128, batches=4):
    """The reference's CuPy path (GPU0402:64-136: materialised (batch x N_back) temporaries of
    dist, amplitude, phase, exp, then a ZGEMV), restated op for op in torch because cupy is not
    installed.  A BASELINE timed next to the fused kernel, never part of the product path.  The
    reference sizes its batch from free memory / 4 (GPU0402:93-97); here the batch is bounded to
    2 GiB complex128 temporaries so the sample stays small -- per-term cost does not depend on it."""
    dx, dy, dz = det
    sx, sy, sz = src
    w = u * ds                                                     # GPU0402:67
    out = torch.empty(batch * batches, dtype=torch.complex128, device=dx.device)

    def one(i):
        xs, ys, zs = dx[i:i + batch], dy[i:i + batch], dz[i:i + batch]
        dist = torch.sqrt((xs[:, None] - sx[None, :]) ** 2 + (ys[:, None] - sy[None, :]) ** 2 +
                          (zs[:, None] - sz[None, :]) ** 2)       # GPU0402:112-116
        amplitude = 1.0 / dist
        phase = -k * dist
        factor = amplitude * torch.exp(1j * phase)
        out[i:i + batch] = torch.mv(factor, w)                     # cp.dot(u_back_u, factor.T)

    one(0)  # warm-up (allocator, kernels)
    ms = None
    for _ in range(3):  # best of 3: the caching allocator settles after the first pass
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for b in range(batches):
            one(b * batch)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) if ms is None else min(ms, e0.elapsed_time(e1))
    return out, batch * batches * sx.shape[0] / (ms * 1e-3), ms


def bench_ray_c2(akb, workloads, torch, dev, peaks, peak_src, n=3163, reps=5):
    co, ray, src = workloads.c2_rays(n, dev)
    N = ray.shape[1]
    times = {}
    for want_normal, bytes_per_ray in ((True, 120.0), (False, 96.0)):
        best = None
        for r in range(reps + 2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            akb.intersect_reflect(co, ray, src, want_normal=want_normal, check=False)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if r >= 2:
                best = ms if best is None else min(best, ms)
        times[want_normal] = (best, bytes_per_ray)
    ms, bpr = times[True]
    gbs = N * bpr / (ms * 1e-3) / 1e9
    ms2, bpr2 = times[False]
    # the reference's NumPy path (ER3D:18-71 via ell.calc_reflect, ER3D:241-245) restated op for op in
    # oracle/numpy_port.py, on the first 1e6 rays of the same bundle, one host thread (NumPy elementwise
    # code is single-threaded in the reference too)
    from oracle import numpy_port
    n_cpu = min(N, 1_000_000)
    rh, sh = ray[:, :n_cpu].cpu().numpy(), src[:, :n_cpu].cpu().numpy()
    coh = np.asarray(co, dtype=np.float64)
    t0 = time.perf_counter()
    ph = numpy_port.mirr_ray_intersection(coh, rh, sh)
    nh = numpy_port.norm_vector(coh, ph)
    numpy_port.reflect_ray(rh, nh)
    cpu_s = time.perf_counter() - t0
    cpu = {"value": n_cpu / cpu_s, "unit": "rays/s", "cores": 1, "kind": "port",
           "sample": f"first {n_cpu} rays of the C2 bundle, oracle/numpy_port.py (NumPy restatement of ER3D:18-71), {cpu_s:.2f} s"}
    return {"kernel": "intersect_reflect_strided_kernel<2, true> (ell.calc_reflect: points + normal + reflect)",
            "workload": f"C2: {N} rays, single elliptical mirror", "cpu_baseline": cpu,
            "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
            "traffic": None, "peak_source": peak_src, "bytes_per_ray": bpr, "rays_per_s": N / (ms * 1e-3),
            "kernel_ms": ms, "without_normal": {"kernel": "intersect_reflect_strided_kernel<2, false>", "bytes_per_ray": bpr2,
                                                "kernel_ms": ms2, "achieved": N * bpr2 / (ms2 * 1e-3) / 1e9,
                                                "rays_per_s": N / (ms2 * 1e-3)},
            "timing": "best of 5 after 2 warm-ups, CUDA events; arrays of 240 MB each exceed L2"}


def bench_chain(workloads, torch, L, _lib, peaks, peak_src, n=3163, reps=5):
    """The fused K-mirror chain at 1e7 rays through the C-ABI on preallocated device buffers: K = 2 (KB geometry) and
    K = 4 (AKB geometry).  Outputs: hit points + last direction + detector point + the K segment lengths (what
    trace_chain returns by default and the Fresnel step consumes: SURVEY 8d, 48 + 24 K + 24 + 24 + 8 K bytes per ray), and
    the same with the summed optical path instead of the segments (+ 8 bytes per ray)."""
    import ctypes
    out = {}
    for tag in ("c3", "c4"):
        coeffs, neg, plane, ray, src = workloads.chain_inputs(tag, n, "cuda")
        K, N = len(neg), ray.shape[1]
        co = np.ascontiguousarray(np.asarray(coeffs, dtype=np.float64)); ng = np.ascontiguousarray(np.asarray(neg, dtype=np.int32))
        pl = np.ascontiguousarray(np.asarray(plane, dtype=np.float64))
        e = lambda *s: torch.empty(*s, dtype=torch.float64, device="cuda")  # noqa: E731
        pts, last, det, dist, opl = e(K, 3, N), e(3, N), e(3, N), e(K, N), e(N)
        flags = torch.empty(4, dtype=torch.int32, device="cuda")
        p = _lib.dev_ptr
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for form, d_ptr, o_ptr, extra in (("segments", p(dist), None, 8 * K), ("opl", None, p(opl), 8)):
            best = None
            for r in range(reps + 2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = L.akb_trace_chain(_lib.host_ptr(co), _lib.host_ptr(ng), K, _lib.host_ptr(pl), p(ray), p(src), N, p(pts), None,
                                       None, p(last), p(det), d_ptr, o_ptr, 0, p(flags), st)
                e1.record()
                torch.cuda.synchronize()
                _lib.check(rc, "akb_trace_chain")
                if r >= 2:
                    best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
            bpr = 48 + 24 * K + 24 + 24 + extra
            gbs = N * bpr / (best * 1e-3) / 1e9
            hits = K * N / (best * 1e-3)
            out[f"K{K}_{form}"] = {
                "workload": f"{'KB' if K == 2 else 'AKB'} chain, {N} rays, {K} mirrors + plane + "
                            f"{'segment lengths' if form == 'segments' else 'optical path'}",
                "bytes_per_ray": bpr, "kernel_ms": best, "achieved": gbs, "frac": gbs / peaks["hbm_gbs"],
                "rays_per_s": N / (best * 1e-3), "mirror_hits_per_s": hits, "misses": int(flags[0]),
                # second bound: ~206 executed FP64 instructions per mirror hit (ncu op counters of the K = 2 launch,
                # profiles/r02d_ncu_full_chain_kernel.md) against 148 SM x 64 FP64 lanes x 1.965 GHz
                "fp64_pipe_frac_est": hits * 206.0 / (148 * 64 * 1.965e9)}
        del pts, last, det, dist, opl, ray, src
        torch.cuda.empty_cache()
    k2 = out["K2_segments"]
    return {"kernel": "trace_chain_kernel<1, 4> (one ray per thread, 4 resident blocks/SM, streaming loads/stores)", "bound": "hbm",
            "achieved": k2["achieved"], "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": k2["frac"], "traffic": None,
            "peak_source": peak_src, "cases": out,
            "note": "headline = K2_segments.  K = 4 does four mirrors' worth of FP64 work (~180 FP64 instructions per mirror "
                    "hit, bit-exact reference operation order) for 200-224 B per ray: it runs into the FP64 pipe and its "
                    "latencies before HBM (ncu: profiles/r02d_ncu_full_chain_kernel.md)",
            "timing": "best of 5 after 2 warm-ups, CUDA events, C-ABI on preallocated buffers; arrays exceed L2"}


def bench_small_call(akb, workloads, reps=30):
    """BASELINE config C1 (1e4 sources -> 64x64 grid) through the host-buffer call: latency of one call."""
    c = workloads.c1_patch()
    args = (c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], c["k"], c["ds"])
    for _ in range(3):
        akb.forward_propagation_numpy_batch(*args)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        akb.forward_propagation_numpy_batch(*args)
        ts.append(time.perf_counter() - t0)
    med = float(np.median(ts))
    return {"median": med * 1e6, "min": float(min(ts)) * 1e6, "terms_per_s": len(c["x"]) * len(c["sx"]) / med,
            "workload": "C1: 4096 detector points x 1e4 sources, forward_propagation_numpy_batch with NumPy arrays "
                        "(one staged H2D copy, pack + pair + reduce kernels, one D2H copy)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
