#!/usr/bin/env python
"""Time the HBM-bound ray kernels: the fused single-mirror kernel at config C2 (3163^2 ~ 1e7 rays) and the fused
K-mirror chain at 1e7 rays (K = 2: KB geometry, K = 4: AKB geometry), best of 5 after 2 warm-ups, CUDA events on the
launching stream, device pointers, outputs preallocated (the C-ABI is timed, not the Python mirror).
Usage: ray_bench.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import akbraytracing_b200 as akb  # noqa: E402
from akbraytracing_b200 import workloads, _lib  # noqa: E402

co, ray, src = workloads.c2_rays(3163, "cuda")
N = ray.shape[1]
for want_normal, bpr in ((True, 120.0), (False, 96.0)):
    best = 1e9
    for r in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        akb.intersect_reflect(co, ray, src, want_normal=want_normal, check=False)
        e1.record()
        torch.cuda.synchronize()
        if r >= 2:
            best = min(best, e0.elapsed_time(e1))
    print(f"single mirror, normal={want_normal}: {best:.4f} ms  {N * bpr / best / 1e6:.0f} GB/s  {N / best / 1e6:.1f} Grays/s")
del ray, src
torch.cuda.empty_cache()


def chain(tag, n, outputs):
    """outputs: subset of {'last', 'det', 'dist', 'opl'}; points are always written."""
    coeffs, neg, plane, ray, src = workloads.chain_inputs(tag, n, "cuda")
    K, N = len(neg), ray.shape[1]
    L = _lib.load()
    co = np.ascontiguousarray(np.asarray(coeffs, dtype=np.float64)); ng = np.ascontiguousarray(np.asarray(neg, dtype=np.int32))
    pl = np.ascontiguousarray(np.asarray(plane, dtype=np.float64))
    e = lambda *s: torch.empty(*s, dtype=torch.float64, device="cuda")  # noqa: E731
    pts = e(K, 3, N)
    last = e(3, N) if "last" in outputs else None
    det = e(3, N) if "det" in outputs else None
    dist = e(K, N) if "dist" in outputs else None
    opl = e(N) if "opl" in outputs else None
    flags = torch.empty(4, dtype=torch.int32, device="cuda")
    p = lambda t: _lib.dev_ptr(t) if t is not None else None  # noqa: E731
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    best = 1e9
    for r in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = L.akb_trace_chain(_lib.host_ptr(co), _lib.host_ptr(ng), K, _lib.host_ptr(pl), p(ray), p(src), N, p(pts), None, None,
                               p(last), p(det), p(dist), p(opl), 0, p(flags), st)
        e1.record()
        torch.cuda.synchronize()
        _lib.check(rc, "akb_trace_chain")
        if r >= 2:
            best = min(best, e0.elapsed_time(e1))
    bpr = 48 + 24 * K + (24 if last is not None else 0) + (24 if det is not None else 0) + (8 * K if dist is not None else 0) \
        + (8 if opl is not None else 0)
    print(f"chain {tag} K={K} N={N} outputs points+{sorted(outputs)}: {best:.4f} ms  {bpr} B/ray  {N * bpr / best / 1e6:.0f} GB/s  "
          f"{N / best / 1e6:.1f} Grays/s  miss={int(flags[0])}")


for tag in ("c3", "c4"):
    chain(tag, 3163, {"last", "det", "opl"})
    chain(tag, 3163, {"last", "det", "dist"})
    chain(tag, 1000, {"last", "det", "opl"})
