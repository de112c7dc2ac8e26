#!/usr/bin/env python
"""The reference's whole Wavecalc workflow at 1e6 points per mirror, on the device:
trace the AKB four-mirror chain (1000 x 1000 rays) -> write the hand-off folder saveWaveData would write -> run the stage
chain of Wavecalc_raytrace_fromData_CPU0402.py:247-375 (source -> M1 -> M2 -> M3 -> M4 -> Image -> Image2: three
mirror-to-mirror stages of 1e12 terms each + two focal grids) with every field resident in HBM.
Timed with the reference's roundings (phase_mode='faithful') and with the per-stage choice ('auto'); the two results are
compared, and the last stage is spot-checked against the CPU oracle fed with the device-computed M4 field.

    python tools/full_chain.py [n] [G]                 (1 GPU)
    python -m torch.distributed.run --nproc-per-node N ... tools/full_chain.py   (stages sharded over N GPUs)
"""
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import akbraytracing_b200 as akb  # noqa: E402
import oracle  # noqa: E402
from akbraytracing_b200 import workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
G = int(sys.argv[2]) if len(sys.argv) > 2 else 512
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()
    finally:
        os.dup2(saved, 1)
        os.close(saved)

coeffs, neg, plane, ray, src = workloads.chain_inputs("c4", n, dev)
t0 = time.perf_counter()
tr = akb.trace_chain(coeffs, neg, plane, ray, src)
torch.cuda.synchronize()
trace_s = time.perf_counter() - t0
folder = tempfile.mkdtemp(prefix="akb_full_chain_") if rank == 0 else None
if world > 1:
    box = [folder]
    dist.broadcast_object_list(box, src=0)
    folder = box[0]
if rank == 0:
    t0 = time.perf_counter()
    akb.write_handoff(folder, src[:, 0], [tr["points"][k] for k in range(4)], (n, n), tr["det"], det_defocus=tr["det"],
                      option_HighNA=True, focus_shape=(G, G), defocus=1e-3)
    handoff_s = time.perf_counter() - t0
if world > 1:
    dist.barrier()
terms = 1.0 * n * n + 3.0 * (n * n) ** 2 + 2.0 * (n * n) * G * G
res = {}
for mode in ("faithful", "auto"):
    akb.run_stage_chain(folder, phase_mode=mode, keep_on_device=True) if mode == "faithful" and n <= 200 else None  # warm-up for small runs
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    out = akb.run_stage_chain(folder, phase_mode=mode, keep_on_device=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    res[mode] = (time.perf_counter() - t0, out)
if rank == 0:
    f, a = res["faithful"][1], res["auto"][1]
    dev_rel = {k: float(torch.linalg.vector_norm(a[k] - f[k]) / torch.linalg.vector_norm(f[k])) for k in f}
    # last stage against the oracle, fed with the device-computed M4 field
    h = akb.load_handoff(folder)
    grid = np.array(h["gridImage"])
    for r in range(3):
        m = np.mean(grid[r, :]); grid[r, :] = (grid[r, :] - m) * 2.0 + m
    sel = np.sort(np.random.default_rng(0).choice(G * G, 32, replace=False))
    m4 = h["mirrors"][3]
    ref = oracle.fresnel_sum(grid[0][sel], grid[1][sel], grid[2][sel], m4[0], m4[1], m4[2], f["M4"].cpu().numpy(), 2 * np.pi / 13.5e-9, m4[3])
    got = f["Image"].cpu().numpy()[sel]
    print(json.dumps({
        "workflow": f"AKB Wavecalc chain, {n * n} points per mirror, {G}x{G} focal grids: source->M1->M2->M3->M4->Image->Image2",
        "n_gpus": world, "terms": terms, "trace_s": trace_s, "write_handoff_s": handoff_s,
        "chain_s_faithful": res["faithful"][0], "terms_per_s_faithful": terms / res["faithful"][0],
        "chain_s_auto": res["auto"][0], "terms_per_s_auto": terms / res["auto"][0],
        "auto_vs_faithful_rel_l2": dev_rel, "image_peak_same": int(a["Image"].abs().argmax()) == int(f["Image"].abs().argmax()),
        "image_stage_rel_l2_vs_oracle_32pts": float(np.linalg.norm(got - ref) / np.linalg.norm(ref)),
        "note": "chain_s includes reading the hand-off files (4 x 32 MB) and H2D; fields stay on the device between stages"}))
if world > 1:
    dist.destroy_process_group()
