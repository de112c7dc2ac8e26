#!/usr/bin/env python
"""Through-focus stack on one GPU: akb_fresnel_sum_planes against the same stack evaluated as a plane-major flat detector
set by akb_fresnel_sum, per phase mode.  In a product build the two are the same kernel; in an AKB_AB_VARIANTS build
with AKB_PLANES_KERNEL=1 the first one runs the rejected "one pixel on four planes per thread" kernel
(profiles/r02_variants_ab.md section 7).  Usage: planes_bench.py [G] [P]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import akbraytracing_b200 as akb  # noqa: E402
from akbraytracing_b200 import workloads, _lib  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 512
P = int(sys.argv[2]) if len(sys.argv) > 2 else 4
L = _lib.load()
w = workloads.traced_field_inputs("c4", 1000, G, device="cuda")
x0 = float(w["det_x"][0])
planes = torch.as_tensor(x0 + np.linspace(-1e-3, 1e-3, P), device="cuda")
y, z = w["det_y"], w["det_z"]
src = (w["src_x"], w["src_y"], w["src_z"], w["u"], w["k"], w["ds"])
gx, gy, gz = planes.repeat_interleave(G * G), y.repeat(P), z.repeat(P)
terms = 1e6 * G * G * P
L.akb_fresnel_timing(1)


def pair_ms():
    p = ctypes.c_double()
    _lib.check(L.akb_fresnel_last_timing(ctypes.byref(p), None, None, None, None), "timing")
    return p.value


for mode, name in ((0, "faithful"), (1, "exact"), (2, "referenced")):
    best_p = best_f = 1e30
    for it in range(3):
        stack = akb.fresnel_sum_planes(y, z, planes, *src, mode=mode)
        best_p = min(best_p, pair_ms()) if it else best_p
        flat = akb.fresnel_sum(gx, gy, gz, *src, mode=mode)
        best_f = min(best_f, pair_ms()) if it else best_f
    dev = float(torch.linalg.vector_norm(stack.reshape(-1) - flat) / torch.linalg.vector_norm(flat))
    print(f"{name}: {P} planes x {G}x{G} x 1e6 sources: akb_fresnel_sum_planes {best_p:.1f} ms = {terms / best_p / 1e6:.1f} Gterms/s; "
          f"plane-major flat set {best_f:.1f} ms = {terms / best_f / 1e6:.1f} Gterms/s; rel-L2 between them {dev:.1e}")
