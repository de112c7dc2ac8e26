#!/usr/bin/env python
"""ncu driver for the A/B of profiles/r02_variants_ab.md section 7: one through-focus stack through
akb_fresnel_sum_planes and the same stack as a flat detector set (AKB_AB_VARIANTS build + AKB_PLANES_KERNEL=1 to reach
the four-plane kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import akbraytracing_b200 as akb
from akbraytracing_b200 import workloads
G, P = 256, 4
w = workloads.traced_field_inputs("c4", 1000, G, device="cuda")
x0 = float(w["det_x"][0])
planes = torch.as_tensor(x0 + np.linspace(-1e-3, 1e-3, P), device="cuda")
y, z = w["det_y"], w["det_z"]
src = (w["src_x"], w["src_y"], w["src_z"], w["u"], w["k"], w["ds"])
gx, gy, gz = planes.repeat_interleave(G * G), y.repeat(P), z.repeat(P)
a = akb.fresnel_sum_planes(y, z, planes, *src)
b = akb.fresnel_sum(gx, gy, gz, *src)
torch.cuda.synchronize()
print("ok", float((a.reshape(-1) - b).abs().max()))
