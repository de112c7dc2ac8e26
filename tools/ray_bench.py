#!/usr/bin/env python
"""Time the fused single-mirror kernel at config C2 (3163^2 rays).  Usage: ray_bench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import akbraytracing_b200 as akb  # noqa: E402
from akbraytracing_b200 import workloads  # noqa: E402

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
    print(f"normal={want_normal}: {best:.4f} ms  {N * bpr / best / 1e6:.0f} GB/s  {N / best / 1e6:.1f} Grays/s "
          f"scalar={os.environ.get('AKB_RAY_SCALAR', '0')}")
coeffs, neg, plane, ray4, src4 = workloads.chain_inputs("c4", 1000, "cuda")
for r in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    akb.trace_chain(coeffs, neg, plane, ray4, src4, check=False)
    e1.record()
    torch.cuda.synchronize()
print(f"chain c4 (4 mirrors + plane + dist, 1e6 rays): {e0.elapsed_time(e1):.4f} ms (includes torch.empty of outputs)")
