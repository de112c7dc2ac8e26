#!/usr/bin/env python
"""Headline benchmark: Fresnel terms/s of the Huygens-Fresnel pair sum on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path over one synthetic batch, inputs resident in HBM:
  KB two-mirror trace of 1000x1000 rays (fused chain kernel) -> calc_dS -> exp(-ik OPL)
  -> Fresnel pair sum of the 1e6 last-mirror points onto the focal grid.
N = 1: BASELINE config C3 (1e6 rays -> 512x512 grid, 2.6e11 terms per step).
N > 1: weak scaling -- every rank holds a 512x512 block of a 512 x (512 N) grid (N = 8 is half of
C4's 2048x2048 detector), the source set is replicated, detector blocks are array_split
contiguous blocks, and the step ends with the NCCL all-gather of the field.

Prints ONE JSON line (rank 0).  `value` = terms of all ranks / max-over-ranks device time.
`e2e` = the same stage through the reference-facing host-buffer call
(forward_propagation_numpy_batch -> akb_fresnel_sum_host: H2D + kernels + D2H inside).
`--impl reference` times the CPU restatement of the reference's numba path (oracle/, all host
threads) on a bounded detector subset of the same workload.
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

GRID = 512            # focal grid side per rank
RAYS = 1000           # ray grid side -> 1e6 source points
WAVELENGTH = 13.5e-9  # CPU0402:243
ALG_FLOP_PER_TERM = 23.0   # SURVEY.md 8(d) convention
# Counted in the SASS of the loop this workload runs (faithful mode, planar-row blocks: the focal grid is a
# plane x = const whose rows align with the 4 points of a thread; tools/sass_cost.py), per pair:
# 14 DFMA (x2) + 8.25 DMUL + 5.25 DADD.  (General loop, irregular detector sets: 14 + 10 + 7 = 31 instr, 45 flop.)
EXEC_FLOP_PER_TERM = 41.5
FP64_INSTR_PER_TERM = 27.5  # each occupies the FP64 pipe of an SM sub-partition for >= 2 cycles


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

def cpu_workload(n_rays=RAYS):
    """The same C3 source set, built WITHOUT the CUDA library: oracle chain trace + oracle calc_dS."""
    import oracle
    geo = np.load(os.path.join(ROOT, "akbraytracing_b200", "data", "geometry.npz"))
    tan_h, tan_v = geo["c3__tan_h"], geo["c3__tan_v"]
    n = n_rays
    raw = np.vstack([np.ones(n * n), np.tile(tan_h, n), np.repeat(tan_v, n)])
    ray = oracle.normalize_vector(raw)
    src = np.repeat(geo["c3__source_point"][:, None], n * n, axis=1)
    tr = oracle.trace_chain(list(geo["c3__coeffs"]), [bool(b) for b in geo["c3__negative"]], geo["c3__plane"], ray, src)
    last = tr["points"][-1]
    k = 2.0 * np.pi / WAVELENGTH
    opl = tr["dist"][0] + tr["dist"][1]
    u = np.exp(-1j * (k * opl))
    ds = oracle.calc_dS(last, n, n).ravel()
    det = tr["det"]
    yc, zc = (det[1].min() + det[1].max()) / 2, (det[2].min() + det[2].max()) / 2
    yy, zz = np.meshgrid(np.linspace(yc - 1e-6, yc + 1e-6, GRID), np.linspace(zc - 1e-6, zc + 1e-6, GRID))
    return dict(det_x=np.full(GRID * GRID, det[0].mean()), det_y=yy.ravel(), det_z=zz.ravel(),
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
    """Oracle port (C + OpenMP, all host threads) on a bounded detector subset of the C3 stage."""
    import oracle
    oracle.build()
    threads = host_threads()
    w = cpu_workload()
    time_cpu_sample(w, max(threads, 16), threads)            # warms the thread pool
    rate, _ = time_cpu_sample(w, 4 * max(threads, 16), threads)  # calibration
    n_det = int(min(GRID * GRID, max(threads, rate * target_s / w["src_x"].shape[0])))
    n_det = max(threads, (n_det // threads) * threads)
    rate, dt = time_cpu_sample(w, n_det, threads)
    return {"value": rate, "unit": "terms/s", "cores": threads, "kind": "port",
            "sample": f"{n_det} random detector points of the C3 grid x all 1e6 source points "
                      f"({n_det * w['src_x'].shape[0]:.3g} terms, {dt:.1f} s), oracle/akb_oracle.c with OpenMP"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.build()
    threads = host_threads()
    w = cpu_workload()
    n_src = w["src_x"].shape[0]
    time_cpu_sample(w, max(threads, 16), threads)
    rate, _ = time_cpu_sample(w, 4 * max(threads, 16), threads)
    step_s = float(os.environ.get("AKB_BENCH_REF_STEP_S", "4.0"))  # CPU seconds per step (tests shorten it)
    n_det = int(max(threads, rate * step_s / n_src))
    n_det = min(GRID * GRID, max(threads, (n_det // threads) * threads))
    for i in range(args.warmup):
        time_cpu_sample(w, n_det, threads, rng_seed=100 + i)
    t = 0.0
    for i in range(args.steps):
        _, dt = time_cpu_sample(w, n_det, threads, rng_seed=i)
        t += dt
    value = args.steps * n_det * n_src / t
    sample = (f"per step: {n_det} random detector points of the C3 512x512 grid x all {n_src} source points "
              f"({n_det * n_src:.3g} terms), CPU restatement of CPU0402:71-124 (oracle/akb_oracle.c, OpenMP)")
    print(json.dumps({
        "impl": "reference", "metric": "fresnel_terms_per_s", "value": value, "unit": "terms/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": "terms/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "terms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(n_gpus):
    name = ("C3: KB two-mirror trace of 1e6 rays + Fresnel sum onto a 512x512 focal grid, 1 GPU" if n_gpus == 1 else
            f"C3 weak-scaled: 1e6 source points x 512x{512 * n_gpus} focal grid, one 512x512 array_split block per "
            f"GPU, NCCL all-gather (N=8 is half of C4's 2048x2048 detector)")
    return {"workload": name, "rays": RAYS * RAYS, "mirrors": 2, "detector_points": GRID * GRID * n_gpus,
            "terms_per_step": float(RAYS * RAYS) * GRID * GRID * n_gpus, "wavelength_m": WAVELENGTH,
            "phase_mode": "faithful", "l2": "flushed between timed steps (256 MiB write)"}


# ------------------------------------------------------------------------------------ our arm

def run_ours(args):
    import torch
    import torch.distributed as dist
    from akbraytracing_b200 import build as akb_build
    akb_build.ensure_built()  # no-op when the in-tree library is current; ranks serialise on a file lock
    import akbraytracing_b200 as akb
    from akbraytracing_b200 import handoff, raytrace, workloads, _lib
    import ctypes

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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident inputs
    coeffs, neg, plane, ray, src = workloads.chain_inputs("c3", RAYS, dev)
    k = 2.0 * np.pi / WAVELENGTH
    tr0 = raytrace.trace_chain(coeffs, neg, plane, ray, src)
    G_total = GRID * GRID * world
    # global detector grid: 512 x (512*world); rank r owns the r-th array_split block
    det = tr0["det"]
    yc = float((det[1].min() + det[1].max()) / 2); zc = float((det[2].min() + det[2].max()) / 2)
    yg = torch.linspace(yc - 1e-6, yc + 1e-6, GRID, dtype=torch.float64, device=dev)
    zg = torch.linspace(zc - 1e-6 * world, zc + 1e-6 * world, GRID * world, dtype=torch.float64, device=dev)
    zz, yy = torch.meshgrid(zg, yg, indexing="ij")
    gx = torch.full((G_total,), float(det[0].mean()), dtype=torch.float64, device=dev)
    gy, gz = yy.reshape(-1).contiguous(), zz.reshape(-1).contiguous()
    begin, count = _lib.shard_range(G_total, world, rank)
    sl = slice(begin, begin + count)
    dx, dy, dz = gx[sl].contiguous(), gy[sl].contiguous(), gz[sl].contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    gathered = torch.empty(G_total, dtype=torch.complex128, device=dev) if world > 1 else None

    def step():
        tr = raytrace.trace_chain(coeffs, neg, plane, ray, src, check=False)
        last = tr["points"][-1]
        u = handoff.opl_to_field(tr["dist"][0] + tr["dist"][1], k)
        ds = handoff.calc_dS(last, RAYS, RAYS).reshape(-1)
        field = akb.fresnel_sum(dx, dy, dz, last[0], last[1], last[2], u, k, ds)
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(torch.float64), field.view(torch.float64))
            return gathered
        return field

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
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(pair_mean, op=dist.ReduceOp.MAX)
    total_ms, pair_mean = float(total_ms.item()), float(pair_mean.item())
    terms_step = float(RAYS * RAYS) * G_total
    value = terms_step * args.steps / (total_ms * 1e-3)
    plan = {"source_splits": sp.value, "detector_blocks": bx.value, "resident_blocks_per_sm": ps.value}

    # ---- end to end through the host-buffer call (H2D + kernels + D2H inside the timed region)
    tr = raytrace.trace_chain(coeffs, neg, plane, ray, src)
    last = tr["points"][-1]
    u = handoff.opl_to_field(tr["dist"][0] + tr["dist"][1], k)
    ds = handoff.calc_dS(last, RAYS, RAYS).reshape(-1)

    def pinned(t):
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        return h.numpy()
    host = [pinned(t) for t in (dx, dy, dz, last[0].contiguous(), last[1].contiguous(), last[2].contiguous(), u, ds)]
    h2d = sum(a.nbytes for a in host)
    d2h = count * 16
    e2e_steps = max(1, min(args.steps, 3))
    akb.forward_propagation_numpy_batch(*host[:7], k, host[7])  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_out = akb.forward_propagation_numpy_batch(*host[:7], k, host[7])
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = terms_step * e2e_steps / float(e2e_s.item())
    same = float((torch.as_tensor(host_out).to(dev) - out[sl] if world > 1 else
                  torch.as_tensor(host_out).to(dev) - out).abs().max())

    result = None
    if rank == 0:
        peaks, peak_src = measured_peaks()
        tf = ctypes.c_double()
        _lib.check(L.akb_fp64_peak_probe(4096, ctypes.byref(tf), None), "akb_fp64_peak_probe")
        fp64_peak = tf.value
        pair_terms = float(RAYS * RAYS) * count
        achieved = pair_terms * ALG_FLOP_PER_TERM / (pair_mean * 1e-3) / 1e12
        roofline = {
            "kernel": "fresnel_pairs_kernel<faithful> [%s]" % L.akb_fresnel_variant_name().decode(), "bound": "fp64",
            "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
            # dram__bytes_read.sum + dram__bytes_write.sum of THIS launch (1e6 sources x 512x512 detectors,
            # 15 source splits) from ncu --set full: profiles/r01h_ncu_full_bench_launch.md (57.7 MB + 24.7 MB).
            # Algorithmic bytes: 48 MB packed sources + 6.3 MB detector xyz + 63 MB partial sums (mostly L2-resident
            # until the reduction kernel consumes them).
            "traffic": 82.3e6 if world == 1 else None, "traffic_unit": "bytes per launch (ncu, round 1)",
            "peak_source": "on-box DFMA microbenchmark (akb_fp64_peak_probe, measured in this run); "
                           "MEASURED_PEAKS.json has no FP64 figure; nominal 148 SM x 64 lanes x 2 x 1.965 GHz = 37.2",
            "algorithmic_flop_per_term": ALG_FLOP_PER_TERM,
            "executed_flop_per_term": EXEC_FLOP_PER_TERM,
            "achieved_exec": achieved * EXEC_FLOP_PER_TERM / ALG_FLOP_PER_TERM,
            "frac_exec": achieved * EXEC_FLOP_PER_TERM / ALG_FLOP_PER_TERM / fp64_peak,
            # issue-slot view of the same pipe: a DFMA-only stream reaches `peak` with one FP64
            # instruction per 2 cycles; this kernel issues FP64_INSTR_PER_TERM of them per term
            "fp64_instr_per_term": FP64_INSTR_PER_TERM,
            "frac_pipe_issue": (pair_terms / (pair_mean * 1e-3)) * FP64_INSTR_PER_TERM * 2.0 / (fp64_peak * 1e12),
            "kernel_ms": pair_mean, "kernel_share_of_step": pair_mean * args.steps / total_ms,
            "terms_per_s_kernel": pair_terms / (pair_mean * 1e-3), "plan": plan,
        }
        # the reference's GPU path (CuPy, restated in torch) on the first detector points of this rank
        g_out, g_rate, g_ms = gpu0402_restatement(torch, (dx, dy, dz), (last[0].contiguous(), last[1].contiguous(),
                                                  last[2].contiguous()), u, ds, k)
        g_ref = (out[sl] if world > 1 else out)[:g_out.shape[0]]
        gpu_baseline = {"value": g_rate, "unit": "terms/s", "n_gpus": 1,
                        "kind": "restatement of forward_propagation_cupy_batch (GPU0402:64-136) in torch on the same "
                                "B200; cupy is not installed",
                        "sample": f"{g_out.shape[0]} detector points x {last.shape[1]} sources in batches of 128 "
                                  f"({g_ms:.1f} ms, best of 3 passes)",
                        "rel_l2_vs_fused_kernel": float(torch.linalg.vector_norm(g_out - g_ref) /
                                                        torch.linalg.vector_norm(g_ref))}
        del g_out
        torch.cuda.empty_cache()
        # the two non-default phase modes on this rank's block of the same stage (one warm-up, one timed pass each)
        ref_field = out[sl] if world > 1 else out
        phase_modes = {}
        for mode_name, mode_id in (("exact", akb.PHASE_EXACT), ("referenced", akb.PHASE_REFERENCED)):
            for rep_i in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fm = akb.fresnel_sum(dx, dy, dz, last[0], last[1], last[2], u, k, ds, mode=mode_id)
                e1.record()
                torch.cuda.synchronize()
            phase_modes[mode_name] = {
                "terms_per_s": float(RAYS * RAYS) * count / (e0.elapsed_time(e1) * 1e-3),
                "rel_l2_vs_faithful": float(torch.linalg.vector_norm(fm - ref_field) / torch.linalg.vector_norm(ref_field))}
        del fm
        # secondary line: the HBM-bound ray kernel at config C2 (1e7 rays, one mirror)
        ray_roof = bench_ray_c2(akb, workloads, torch, dev, peaks, peak_src)
        result = {
            "metric": "fresnel_terms_per_s", "value": value, "unit": "terms/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world), "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "terms/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "forward_propagation_numpy_batch (akb_fresnel_sum_host), pinned host buffers",
                    "timer": "host wall clock around the synchronous call, max over ranks",
                    "steps": e2e_steps, "max_abs_diff_vs_device_path": same},
            "gpu_launches": int(launches) * world, "roofline": roofline, "roofline_ray": ray_roof,
            "gpu_baseline": gpu_baseline, "phase_modes": phase_modes,
        }
    if world > 1:
        dist.barrier()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            result["cpu_baseline"] = cpu_baseline()
        print(json.dumps(result))
    if world > 1:
        dist.destroy_process_group()


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
    return {"kernel": "intersect_reflect_kernel<2, normal>", "workload": f"C2: {N} rays, single elliptical mirror",
            "cpu_baseline": cpu,
            "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
            "traffic": None, "peak_source": peak_src, "bytes_per_ray": bpr, "rays_per_s": N / (ms * 1e-3),
            "kernel_ms": ms, "without_normal": {"bytes_per_ray": bpr2, "kernel_ms": ms2,
                                                "achieved": N * bpr2 / (ms2 * 1e-3) / 1e9,
                                                "rays_per_s": N / (ms2 * 1e-3)},
            "timing": "best of 5 after 2 warm-ups, CUDA events; arrays of 240 MB each exceed L2"}


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
