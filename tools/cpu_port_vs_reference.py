#!/usr/bin/env python
"""BUILD-CONTAINER ONLY (needs /root/reference): time the reference's own numba path
(CPU0402.compute_u_parallel via forward_propagation_numpy_batch) next to the C/OpenMP port in oracle/
on identical inputs and thread counts, and check they agree bit for bit.  Evidence for using the port
as bench.py's CPU arm on the GPU box, where the reference tree does not exist.
Usage: python tools/cpu_port_vs_reference.py [n_src] [G]"""
import contextlib
import io
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import numpy as np  # noqa: E402

import _refload  # noqa: E402
import oracle  # noqa: E402
from akbraytracing_b200 import workloads  # noqa: E402

n_src = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ref = _refload.load_cpu0402()
import numba  # noqa: E402

threads = numba.get_num_threads()
c = workloads.c1_patch(n_src=n_src, G=G)
args = (c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], c["k"], c["ds"])
oracle.build()


def best_of(fn, reps=3):
    best, out = 1e30, None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


with contextlib.redirect_stdout(io.StringIO()):
    ref.forward_propagation_numpy_batch(*args)  # JIT compile
    t_ref, u_ref = best_of(lambda: ref.forward_propagation_numpy_batch(*args))
t_port, u_port = best_of(lambda: oracle.fresnel_sum(*args, nthreads=threads))
terms = len(c["x"]) * len(c["sx"])
print(f"{terms:.3g} terms, {threads} threads (numba {numba.__version__}, numpy {np.__version__})")
print(f"reference numba path : {t_ref * 1e3:9.1f} ms  {terms / t_ref:.3e} terms/s")
print(f"oracle C/OpenMP port : {t_port * 1e3:9.1f} ms  {terms / t_port:.3e} terms/s   ({t_ref / t_port:.2f}x the reference's speed)")
print(f"bit-identical: {np.array_equal(u_ref, u_port)}   max |diff| / max |u| = {np.abs(u_ref - u_port).max() / np.abs(u_ref).max():.2e}")
