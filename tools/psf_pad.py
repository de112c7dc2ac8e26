#!/usr/bin/env python
"""What the reference's pad_factor = 16 (psf_calc, AKB_raytrace_20250312.py:1200) costs for a 1024 x 1024 plane on one B200:
time and peak memory of throughfocus.psf_stack for P planes at pad_factor 2 and 16.  Usage: psf_pad.py [planes]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import akbraytracing_b200 as akb  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = 1024
g = torch.Generator(device="cuda").manual_seed(0)
stack = torch.polar(torch.rand(P, n * n, dtype=torch.float64, device="cuda", generator=g),
                    6.28 * torch.rand(P, n * n, dtype=torch.float64, device="cuda", generator=g))
for pad in (2, 16):
    torch.cuda.reset_peak_memory_stats()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = akb.psf_stack(stack, (n, n), 13.5e-9, 2e-9, 0.1, pad_factor=pad)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"pad_factor {pad}: {P} planes of {n}x{n} -> PSF {tuple(out['I'].shape)}: {dt * 1e3:.0f} ms "
          f"({dt / P * 1e3:.0f} ms/plane), peak device memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, "
          f"peaks all 1: {bool(((out['I'].amax(dim=(-2, -1)) - 1).abs() < 1e-12).all())}")
    del out
    torch.cuda.empty_cache()
