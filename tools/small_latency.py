#!/usr/bin/env python
"""Where the time of a C1-sized call goes (4096 detector points x 1e4 sources): host-buffer call, device-resident call,
pair kernel alone.  Usage: small_latency.py"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import akbraytracing_b200 as akb  # noqa: E402
from akbraytracing_b200 import workloads, _lib  # noqa: E402

L = _lib.load()
c = workloads.c1_patch()
host = (c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], c["k"], c["ds"])
dev = tuple(torch.as_tensor(a, device="cuda") if isinstance(a, np.ndarray) else a for a in host)
for name, args in (("host buffers (NumPy in/out)", host), ("device tensors", dev)):
    for _ in range(5):
        out = akb.fresnel_sum(*args)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        out = akb.fresnel_sum(*args)
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) / 50 * 1e6:.1f} us per call")
L.akb_fresnel_timing(1)
akb.fresnel_sum(*dev)
p, t = ctypes.c_double(), ctypes.c_double()
sp, bx, ps = ctypes.c_int(), ctypes.c_int64(), ctypes.c_int()
_lib.check(L.akb_fresnel_last_timing(ctypes.byref(p), ctypes.byref(t), ctypes.byref(sp), ctypes.byref(bx), ctypes.byref(ps)), "t")
print(f"device part: pair kernel {p.value * 1e3:.1f} us, pack+pair+reduce {t.value * 1e3:.1f} us, splits {sp.value}, "
      f"detector blocks {bx.value}, blocks/SM {ps.value}, variant {L.akb_fresnel_variant_name().decode()}")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    akb.fresnel_sum(*dev)
    with torch.cuda.graph(g, stream=s):
        out = akb.fresnel_sum(*dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    g.replay()
torch.cuda.synchronize()
print(f"CUDA-graph replay of the device call: {(time.perf_counter() - t0) / 50 * 1e6:.1f} us")
