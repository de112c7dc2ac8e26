#!/usr/bin/env python
"""Time one Fresnel stage (1e6 traced source points -> G x G grid) for the pair-kernel variant chosen with
AKB_FRESNEL_VARIANT, and check it against the first variant's field.  Usage: variant_bench.py [G] [mode] [general]
("general": the same detector points in shuffled order, which takes the general loop like a mirror-to-mirror stage)"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import akbraytracing_b200 as akb  # noqa: E402
from akbraytracing_b200 import workloads, _lib  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L = _lib.load()
general = len(sys.argv) > 3 and sys.argv[3] == "general"
w = workloads.traced_field_inputs("c3", 1000, G, device="cuda")
if general:
    perm = torch.randperm(w["det_x"].shape[0], device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    for key in ("det_x", "det_y", "det_z"):
        w[key] = w[key][perm].contiguous()
args = (w["det_x"], w["det_y"], w["det_z"], w["src_x"], w["src_y"], w["src_z"], w["u"], w["k"], w["ds"])
L.akb_fresnel_timing(1)
best = 1e30
for it in range(4):
    out = akb.fresnel_sum(*args, mode=mode)
    p, t = ctypes.c_double(), ctypes.c_double()
    sp, bx, ps = ctypes.c_int(), ctypes.c_int64(), ctypes.c_int()
    _lib.check(L.akb_fresnel_last_timing(ctypes.byref(p), ctypes.byref(t), ctypes.byref(sp), ctypes.byref(bx), ctypes.byref(ps)), "t")
    if it:
        best = min(best, p.value)
terms = 1e6 * G * G
ref_path = f"/tmp/variant_ref_{G}_{mode}_{int(general)}.pt"
o = out.cpu()
if os.path.exists(ref_path):
    ref = torch.load(ref_path)
    err = float((o - ref).abs().pow(2).sum().sqrt() / ref.abs().pow(2).sum().sqrt())
else:
    torch.save(o, ref_path)
    err = 0.0
print(f"variant {os.environ.get('AKB_FRESNEL_VARIANT', '0')} [{L.akb_fresnel_variant_name().decode()}] mode {mode}{' general loop' if general else ''}: "
      f"{best:.2f} ms  {terms / best / 1e6:.1f} Gterms/s  splits {sp.value} blocks/SM {ps.value}  rel-L2 vs first {err:.2e}")
