#!/usr/bin/env python
"""Where the host-buffer (e2e) call of the sharded path spends its time: run under torchrun on >= 2 GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/e2e_breakdown.py [G=2048] [N=20000]

A small source set keeps the kernel short so the copies dominate; the copies do not depend on N except for the
56 bytes per source.  "staged" = what the call did before round 2's last change: every rank uploads ALL detector
coordinates and reads the field back into a fresh pageable array.  "product" = fresnel_sum_sharded on NumPy buffers now.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from akbraytracing_b200 import wavecalc  # noqa: E402


def main():
    G = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    rng = np.random.default_rng(3)
    yy, zz = np.meshgrid(np.linspace(-1e-5, 1e-5, G), np.linspace(-1e-5, 1e-5, G), indexing="ij")
    x, y, z = np.full(G * G, 0.3), yy.ravel().copy(), zz.ravel().copy()
    sx, sy, sz = rng.normal(0, 1e-3, N), rng.normal(0, 1e-3, N), rng.normal(0, 1e-3, N)
    su = rng.normal(size=N) + 1j * rng.normal(size=N)
    ds = np.full(N, 1e-9)
    k = 2 * np.pi / 1e-10

    def staged():
        t = [time.perf_counter()]
        dx, dy, dz = (torch.as_tensor(a).to(dev) for a in (x, y, z))
        torch.cuda.synchronize(); t.append(time.perf_counter())
        s = [torch.as_tensor(a).to(dev) for a in (sx, sy, sz, su, ds)]
        torch.cuda.synchronize(); t.append(time.perf_counter())
        out = wavecalc.fresnel_sum_sharded(dx, dy, dz, s[0], s[1], s[2], s[3], k, s[4])
        torch.cuda.synchronize(); t.append(time.perf_counter())
        res = out.cpu().numpy()
        t.append(time.perf_counter())
        return res, np.diff(t) * 1e3

    def product():
        t0 = time.perf_counter()
        res = wavecalc.fresnel_sum_sharded(x, y, z, sx, sy, sz, su, k, ds)
        return res, (time.perf_counter() - t0) * 1e3

    for _ in range(3):
        a, _t = staged()
        b, _t = product()
    assert np.array_equal(a, b)
    rows_s, rows_p = [], []
    for _ in range(7):
        dist.barrier(); torch.cuda.synchronize()
        _, t = staged(); rows_s.append(t)
        dist.barrier(); torch.cuda.synchronize()
        _, t = product(); rows_p.append(t)
    s_med = np.median(np.array(rows_s), axis=0)
    p_med = float(np.median(rows_p))
    both = torch.tensor([s_med.sum(), p_med], device=dev)
    dist.all_reduce(both, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world {dist.get_world_size()}  {G}x{G} detectors, {N} sources (median of 7, ms; rank 0 phases, totals = max over ranks)")
        print(f"  staged : H2D all detectors {s_med[0]:.2f} | H2D sources {s_med[1]:.2f} | kernels + all-gather {s_med[2]:.2f} | "
              f"D2H pageable {s_med[3]:.2f} | total {both[0].item():.2f}")
        print(f"  product: total {both[1].item():.2f}  (copies + host work = {both[1].item() - s_med[2]:.2f} vs {both[0].item() - s_med[2]:.2f})")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
