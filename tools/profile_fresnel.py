#!/usr/bin/env python
"""Small driver for ncu: ONE Fresnel stage (1e6 traced source points -> G x G grid) and one C2-sized
ray launch.  Usage: python tools/profile_fresnel.py [G] [mode] [c3|c4]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import akbraytracing_b200 as akb  # noqa: E402
from akbraytracing_b200 import workloads  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tag = sys.argv[3] if len(sys.argv) > 3 else "c3"
w = workloads.traced_field_inputs(tag, 1000, G, device="cuda")
for _ in range(2):
    out = akb.fresnel_sum(w["det_x"], w["det_y"], w["det_z"], w["src_x"], w["src_y"], w["src_z"], w["u"], w["k"],
                          w["ds"], mode=mode)
torch.cuda.synchronize()
co, ray, src = workloads.c2_rays(3163, "cuda")
for _ in range(2):
    akb.intersect_reflect(co, ray, src, check=False)
torch.cuda.synchronize()
print("ok", complex(out[0]))
