#!/usr/bin/env python
"""Run BASELINE.json's five configs (C1..C5) on the GPU box with a parity spot-check each and print one
JSON object per config (also written to gpurun_out/configs.json on rank 0).

    python tools/run_configs.py [--configs c1,c2,c3,c4,c5] [--planes 32] [--c4-grid 2048] [--c5-grid 1024]
    python -m torch.distributed.run --nproc-per-node N ... tools/run_configs.py --configs c4,c5

C1..C3 run on one GPU (rank 0).  C4/C5 shard the detector points over all ranks (array_split blocks +
NCCL all-gather).  Parity: C1 against the full oracle; C2/C3/C4/C5 on a random subset of rays /
detector points (the CPU oracle needs minutes for the full grids).  Timings: CUDA events, best of 3.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import akbraytracing_b200 as akb  # noqa: E402
import oracle  # noqa: E402
from akbraytracing_b200 import workloads  # noqa: E402


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def timed(fn, reps=3):
    best, out = None, None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best, out


def field_subset_parity(w, full, n_check, seed=5):
    rng = np.random.default_rng(seed)
    sel = np.sort(rng.choice(full.shape[0], n_check, replace=False))
    sel_t = torch.as_tensor(sel, device=full.device)
    h = {k: v.cpu().numpy() for k, v in w.items() if k in ("src_x", "src_y", "src_z", "u", "ds")}
    det = [w[k][sel_t].cpu().numpy() for k in ("det_x", "det_y", "det_z")]
    ref = oracle.fresnel_sum(det[0], det[1], det[2], h["src_x"], h["src_y"], h["src_z"], h["u"], w["k"], h["ds"])
    got = full[sel_t].cpu().numpy()
    return rel_l2(got, ref), int(np.argmax(np.abs(got))) == int(np.argmax(np.abs(ref)))


def run_c1():
    c = workloads.c1_patch()
    ref = oracle.fresnel_sum(c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], c["k"], c["ds"])
    args = (c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], c["k"], c["ds"])
    akb.forward_propagation_numpy_batch(*args)
    t0 = time.perf_counter()
    for _ in range(5):
        got = akb.forward_propagation_numpy_batch(*args)
    dt = (time.perf_counter() - t0) / 5
    terms = len(c["x"]) * len(c["sx"])
    return {"config": "C1: 1e4 rays -> 64x64 grid, host-buffer call (NumPy in/out)", "terms": terms,
            "ms_e2e": dt * 1e3, "terms_per_s_e2e": terms / dt, "rel_l2_vs_oracle_full": rel_l2(got, ref),
            "peak_pixel_same": int(np.argmax(np.abs(got))) == int(np.argmax(np.abs(ref)))}


def run_c2(dev):
    co, ray, src = workloads.c2_rays(3163, dev)
    N = ray.shape[1]
    ms, out = timed(lambda: akb.intersect_reflect(co, ray, src, check=False), reps=5)
    p, n, r = out
    sel = torch.arange(0, N, 7919, device=dev)
    rh, sh = ray[:, sel].cpu().numpy(), src[:, sel].cpu().numpy()
    po = oracle.mirr_ray_intersection(co, rh, sh)
    no = oracle.norm_vector(co, po)
    ro = oracle.reflect_ray(rh, no)
    exact = bool(np.array_equal(p[:, sel].cpu().numpy(), po) and np.array_equal(n[:, sel].cpu().numpy(), no)
                 and np.array_equal(r[:, sel].cpu().numpy(), ro))
    return {"config": "C2: single elliptical mirror, intersect+normal+reflect", "rays": N, "ms": ms,
            "rays_per_s": N / ms * 1e3, "GB_per_s_120B_per_ray": N * 120 / ms / 1e6,
            "subset_bit_identical_to_oracle": exact, "subset": int(sel.numel())}


def run_field_config(tag, name, n_rays, G, dev, world, rank, planes=None, n_check=256, psf=False):
    w = workloads.traced_field_inputs(tag, n_rays, G, device=dev, planes=planes)
    M = w["det_x"].shape[0]
    args = (w["det_x"], w["det_y"], w["det_z"], w["src_x"], w["src_y"], w["src_z"], w["u"], w["k"], w["ds"])

    G2 = G * G

    def step():
        if planes is not None:  # C5 through the product call: planes x pixels flattened, sharded over the ranks
            return akb.fresnel_sum_planes(w["det_y"][:G2], w["det_z"][:G2], w["det_x"][::G2].contiguous(), w["src_x"], w["src_y"],
                                          w["src_z"], w["u"], w["k"], w["ds"]).reshape(-1)
        return akb.forward_propagation_cupy_batch_multi_gpu(*args)  # C4: the reference-shaped multi-GPU call
    step()  # warm-up
    ms, full = timed(step, reps=2 if M * n_rays * n_rays > 1e13 else 3)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    terms = float(M) * n_rays * n_rays
    res = {"config": name, "terms": terms, "n_gpus": world, "ms": ms, "terms_per_s": terms / ms * 1e3}
    if rank == 0:
        err, same_peak = field_subset_parity(w, full, n_check)
        res.update({"rel_l2_vs_oracle_subset": err, "subset": n_check, "subset_peak_pixel_same": same_peak})
    if psf:  # C5: PSF of every plane from the assembled field, planes array_split over the ranks
        npl = len(planes)
        stack = full.reshape(npl, G * G)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = akb.psf_stack(stack, (G, G), workloads.WAVELENGTH_EUV, 2e-6 / (G - 1), 0.1, pad_factor=2)
        torch.cuda.synchronize()
        tp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        ok = torch.tensor([1.0 if bool(((out["I"].amax(dim=(-2, -1)) - 1.0).abs() < 1e-12).all()) else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if rank == 0:
            res["psf_ms_all_planes"] = float(tp.item()) * 1e3
            res["psf_planes_per_rank"] = len(out["planes"])
            res["psf_pad_factor"] = 2
            res["psf_peaks_all_one"] = bool(ok.item() == 1.0)
    return res


def run_trace(tag, name, dev, n=1000, n_cpu=200_000):
    """The ray part of C3 / C4 alone: K mirrors + detector plane + segment lengths for n*n rays in ONE kernel
    (BIG:11039-11054 / BIG:2881-2905), next to the reference's NumPy call sequence restated in
    oracle/numpy_port.py on the first n_cpu rays (one host thread), with a bit-for-bit check on that subset."""
    from oracle import numpy_port
    from akbraytracing_b200 import raytrace
    coeffs, neg, plane, ray, src = workloads.chain_inputs(tag, n, dev)
    K = len(neg)
    raytrace.trace_chain(coeffs, neg, plane, ray, src, check=False)
    ms, tr = timed(lambda: raytrace.trace_chain(coeffs, neg, plane, ray, src, check=False), reps=5)
    N = ray.shape[1]
    rh, sh = ray[:, :n_cpu].cpu().numpy(), src[:, :n_cpu].cpu().numpy()
    t0 = time.perf_counter()
    cur_r, cur_s, pts = rh, sh, []
    for co, ng in zip(coeffs, neg):
        p = numpy_port.mirr_ray_intersection(np.asarray(co), cur_r, cur_s, ng)
        nv = numpy_port.norm_vector(np.asarray(co), p)
        cur_r, cur_s = numpy_port.reflect_ray(cur_r, nv), p
        pts.append(p)
    det = numpy_port.plane_ray_intersection(np.asarray(plane), cur_r, cur_s)
    cpu_s = time.perf_counter() - t0
    same = bool(np.array_equal(tr["det"][:, :n_cpu].cpu().numpy(), det) and
                all(np.array_equal(tr["points"][i][:, :n_cpu].cpu().numpy(), pts[i]) for i in range(K)))
    bytes_per_ray = 48 + 24 * K + 24 + 24 + 8 * K  # in, K hit points, final direction, detector point, K lengths
    return {"config": name, "rays": N, "mirrors": K, "ms": ms, "rays_per_s": N / ms * 1e3,
            "GB_per_s": N * bytes_per_ray / ms / 1e6, "bytes_per_ray": bytes_per_ray,
            "cpu_numpy_rays_per_s_1_thread": n_cpu / cpu_s, "cpu_sample_rays": n_cpu,
            "subset_bit_identical_to_numpy_restatement": same}


def run_mirror_to_mirror(dev, n=700, n_check=192):
    """The stage that dominates the reference's real workflow (CPU0402:283-301): the field on mirror 1
    propagated to the hit points of mirror 2 -- an irregular detector set (no plane, no rows), so the
    pair kernel's general loop runs."""
    from akbraytracing_b200 import handoff, raytrace
    coeffs, neg, plane, ray, src = workloads.chain_inputs("c3", n, dev)
    tr = raytrace.trace_chain(coeffs, neg, plane, ray, src, want_dist=True)
    m1, m2 = tr["points"][0], tr["points"][1]
    k = 2.0 * np.pi / workloads.WAVELENGTH_EUV
    w = dict(det_x=m2[0].contiguous(), det_y=m2[1].contiguous(), det_z=m2[2].contiguous(), src_x=m1[0].contiguous(),
             src_y=m1[1].contiguous(), src_z=m1[2].contiguous(), u=handoff.opl_to_field(tr["dist"][0], k),
             ds=handoff.calc_dS(m1, n, n).reshape(-1), k=k)
    args = (w["det_x"], w["det_y"], w["det_z"], w["src_x"], w["src_y"], w["src_z"], w["u"], w["k"], w["ds"])
    akb.fresnel_sum(*args)
    ms, full = timed(lambda: akb.fresnel_sum(*args), reps=3)
    terms = float(n * n) ** 2
    err, same_peak = field_subset_parity(w, full, n_check)
    return {"config": f"M2M: KB mirror 1 -> mirror 2 stage, {n * n} x {n * n} points (irregular detector set)",
            "terms": terms, "n_gpus": 1, "ms": ms, "terms_per_s": terms / ms * 1e3, "rel_l2_vs_oracle_subset": err,
            "subset": n_check, "subset_peak_pixel_same": same_peak}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c2,c3,c4")
    ap.add_argument("--planes", type=int, default=32)
    ap.add_argument("--c4-grid", type=int, default=2048)
    ap.add_argument("--c5-grid", type=int, default=1024)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)  # library banners (NCCL_DEBUG=VERSION) go to stderr, stdout keeps the JSON lines
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            os.dup2(saved, 1)
            os.close(saved)
    results = []
    for c in args.configs.split(","):
        c = c.strip().lower()
        if c == "c1" and rank == 0:
            results.append(run_c1())
        elif c == "c2" and rank == 0:
            results.append(run_c2(dev))
        elif c == "c3" and rank == 0:
            results.append(run_field_config("c3", "C3: KB two-mirror trace 1e6 rays -> 512x512 grid", 1000, 512, dev, 1, 0))
        elif c == "c3t" and rank == 0:
            results.append(run_trace("c3", "C3 trace: KB two mirrors + plane, 1e6 rays, one fused kernel", dev))
        elif c == "c4t" and rank == 0:
            results.append(run_trace("c4", "C4 trace: AKB four mirrors + plane, 1e6 rays, one fused kernel", dev))
        elif c == "m2m" and rank == 0:
            results.append(run_mirror_to_mirror(dev))
        elif c == "c4":
            G = args.c4_grid
            results.append(run_field_config("c4", f"C4: AKB four-mirror trace 1e6 rays x {G}x{G} detector", 1000, G, dev,
                                            world, rank))
        elif c == "c5":
            G = args.c5_grid
            planes = np.linspace(-1e-3, 1e-3, args.planes)
            results.append(run_field_config("c4", f"C5: through-focus stack {args.planes} planes x {G}x{G} + PSF per plane",
                                            1000, G, dev, world, rank, planes=planes, psf=True))
        if world > 1:
            dist.barrier()
    if rank == 0:
        for r in results:
            print(json.dumps(r))
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/configs_n{world}.json", "w") as fh:
            json.dump(results, fh, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
