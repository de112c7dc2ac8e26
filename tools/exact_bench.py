#!/usr/bin/env python
"""EXACT vs FAITHFUL on the two loops of the pair kernel: the C3 focal-grid stage (planar-row loop) and the
mirror-to-mirror stage of bench.py's `roofline_m2m` (general loop).  Prints kernel ms (library CUDA events), terms/s and the
rel-L2 distance of the EXACT field from the FAITHFUL one.  Usage: python tools/exact_bench.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import akbraytracing_b200 as akb  # noqa: E402
from akbraytracing_b200 import _lib, handoff, raytrace, workloads  # noqa: E402

L = _lib.load()
L.akb_fresnel_timing(1)
dev = torch.device("cuda", 0)
RAYS = 1000


def timed(args, mode):
    best = 1e30
    for it in range(4):
        out = akb.fresnel_sum(*args, mode=mode)
        p, t = ctypes.c_double(), ctypes.c_double()
        L.akb_fresnel_last_timing(ctypes.byref(p), ctypes.byref(t), None, None, None)
        if it:
            best = min(best, p.value)
    return out, best


w = workloads.traced_field_inputs("c3", RAYS, 512, device="cuda")
cases = {"C3 focal grid 512x512 (planar-row loop)": (w["det_x"], w["det_y"], w["det_z"], w["src_x"], w["src_y"], w["src_z"], w["u"], w["k"], w["ds"])}
coeffs, neg, plane, ray, src = workloads.chain_inputs("c4", RAYS, dev)
tr = raytrace.trace_chain(coeffs, neg, plane, ray, src)
back, front = tr["points"][0], tr["points"][1][:, :512 * 512].contiguous()
k = 2.0 * np.pi / workloads.WAVELENGTH_EUV
cases["AKB mirror 1 -> mirror 2, 1e6 x 262144 (general loop)"] = (
    front[0], front[1], front[2], back[0], back[1], back[2], handoff.opl_to_field(tr["dist"][0], k), k,
    handoff.calc_dS(back, RAYS, RAYS).reshape(-1))
for name, args in cases.items():
    terms = float(args[0].shape[0]) * args[3].shape[0]
    ref, ms0 = timed(args, akb.PHASE_FAITHFUL)
    got, ms1 = timed(args, akb.PHASE_EXACT)
    rel = float(torch.linalg.vector_norm(got - ref) / torch.linalg.vector_norm(ref))
    print(f"{name}: faithful {ms0:.2f} ms = {terms / ms0 / 1e6:.1f} Gterms/s; exact {ms1:.2f} ms = {terms / ms1 / 1e6:.1f} Gterms/s; "
          f"exact vs faithful rel-L2 {rel:.2e}")
