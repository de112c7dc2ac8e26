#!/usr/bin/env python
"""Small driver for ncu: ONE mirror-to-mirror Fresnel stage (AKB mirror 1, 1e6 points -> the first 512 x 512 points of
mirror 2: the pair kernel's GENERAL loop, bench.py's `roofline_m2m` workload).  Usage: python tools/profile_m2m.py [mode]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import akbraytracing_b200 as akb  # noqa: E402
from akbraytracing_b200 import handoff, raytrace, workloads  # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
RAYS = 1000
dev = torch.device("cuda", 0)
coeffs, neg, plane, ray, src = workloads.chain_inputs("c4", RAYS, dev)
tr = raytrace.trace_chain(coeffs, neg, plane, ray, src)
back, front = tr["points"][0], tr["points"][1][:, :512 * 512].contiguous()
k = 2.0 * np.pi / workloads.WAVELENGTH_EUV
u = handoff.opl_to_field(tr["dist"][0], k)
ds = handoff.calc_dS(back, RAYS, RAYS).reshape(-1)
for _ in range(2):
    out = akb.fresnel_sum(front[0], front[1], front[2], back[0], back[1], back[2], u, k, ds, mode=mode)
torch.cuda.synchronize()
print("ok", complex(out[0]))
