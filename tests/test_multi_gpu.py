"""Multi-GPU parity tests (-m gpu; they un-skip when >= 2 devices are visible, e.g. `gpurun --gpus 2`).

Three shapes of the reference's multi-GPU call (GPU0402:64-136, GPU0402_multi.py:123-229):
  1. ONE process driving every visible device (what WaveField3D(device='cuda') reaches);
  2. one process per GPU under torch.distributed/NCCL -> akb_fresnel_sum_sharded with PyTorch's communicator
     (NumPy in / NumPy out, device in / device out, uneven shards, source broadcast, own communicator);
  3. a non-torch caller: one HOST THREAD per GPU (the threading model of GPU0402_multi.py:213-225), each binding
     the C-ABI with ctypes only: akb_nccl_unique_id / akb_nccl_comm_init / akb_fresnel_sum_sharded.
Every field is compared with the pinned CPU oracle and across ranks bit for bit."""
import ctypes
import os
import socket
import threading

import numpy as np
import pytest

import oracle
from conftest import rel_l2

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _need_two():
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 visible GPUs")
    from akbraytracing_b200 import build
    build.ensure_built()


def _case(M, N=3001, seed=0):
    rng = np.random.default_rng(seed)
    yy = np.linspace(-1e-6, 1e-6, M) + 3e-4
    x = np.full(M, 0.15); y = yy; z = rng.uniform(-1e-6, 1e-6, M)
    sx = rng.uniform(-1e-2, 1e-2, N); sy = rng.uniform(-1e-3, 1e-3, N); sz = rng.uniform(-1e-3, 1e-3, N)
    u = np.exp(2j * np.pi * rng.uniform(size=N)); ds = rng.uniform(1e-9, 2e-9, N)
    return dict(x=x, y=y, z=z, sx=sx, sy=sy, sz=sz, u=u, ds=ds, k=2 * np.pi / 13.5e-9)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


# ------------------------------------------------------------------ 1. one process, every visible device

@pytest.mark.parametrize("M", [4096, 4099, 3])
def test_single_process_multi_device(golden, M):
    _need_two()
    import torch
    import akbraytracing_b200 as akb
    c = _case(M)
    ref = oracle.fresnel_sum(c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], c["k"], c["ds"])
    got = akb.forward_propagation_cupy_batch_multi_gpu(c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], c["k"], c["ds"])
    assert isinstance(got, np.ndarray) and got.shape == (M,)
    assert rel_l2(got, ref) <= 1e-12
    one = akb.forward_propagation_cupy_batch(c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], c["k"], c["ds"])
    assert rel_l2(got, one) <= 1e-13
    t = {k: torch.as_tensor(v).cuda() for k, v in c.items() if k != "k"}
    dev = akb.forward_propagation_cupy_batch_multi_gpu(t["x"], t["y"], t["z"], t["sx"], t["sy"], t["sz"], t["u"], c["k"], t["ds"])
    assert dev.is_cuda and dev.device.index == 0 and np.array_equal(dev.cpu().numpy(), got)
    # WaveField3D(device='cuda') goes through the same call (GPU0402:52-57)
    back = akb.WaveField3D(len(c["sx"]), 13.5e-9, 1, 1, device="cuda")
    back.setdata(np.vstack([c["sx"], c["sy"], c["sz"]])); back.set_ds(c["ds"]); back.u = t["u"]
    front = akb.WaveField3D(M, 13.5e-9, 1, 1, device="cuda")
    front.setdata(np.vstack([c["x"], c["y"], c["z"]]))
    front.forward_propagation(back)
    assert rel_l2(front.u.cpu().numpy(), ref) <= 1e-12


# ------------------------------------------------------------------ 2. one process per GPU, torch.distributed + NCCL

def _nccl_worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    msgs = []
    try:
        import akbraytracing_b200 as akb
        from akbraytracing_b200 import wavecalc
        for M in (4096, 4099, 1):  # equal blocks (ncclAllGather), uneven tail (grouped broadcasts), fewer points than ranks
            c = _case(M)
            ref = oracle.fresnel_sum(c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], c["k"], c["ds"], nthreads=2)
            # NumPy in -> NumPy out under an NCCL group (used to raise on CPU tensors)
            got = akb.forward_propagation_cupy_batch_multi_gpu(c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], c["k"], c["ds"])
            ok = isinstance(got, np.ndarray) and got.shape == (M,) and rel_l2(got, ref) <= 1e-12
            # device in -> device out, bit-identical on every rank and to the NumPy form
            t = {k: torch.as_tensor(v).cuda() for k, v in c.items() if k != "k"}
            dev = akb.fresnel_sum_sharded(t["x"], t["y"], t["z"], t["sx"], t["sy"], t["sz"], t["u"], c["k"], t["ds"])
            ok = ok and dev.is_cuda and np.array_equal(dev.cpu().numpy(), got)
            both = [torch.empty_like(dev) for _ in range(world)]
            dist.all_gather(both, dev)
            ok = ok and all(torch.equal(b, both[0]) for b in both)
            # the sources live on rank 0 only (the reference keeps them on device 0): broadcast inside the call
            junk = {k: (torch.zeros_like(v) if rank else v) for k, v in t.items()}
            bc = akb.fresnel_sum_sharded(t["x"], t["y"], t["z"], junk["sx"], junk["sy"], junk["sz"], junk["u"], c["k"], junk["ds"],
                                         broadcast_sources=True)
            ok = ok and torch.equal(bc, dev)
            if rank:  # the caller's arrays on ranks > 0 are not overwritten
                ok = ok and float(junk["sx"].abs().max()) == 0.0
            # local block only
            loc = akb.fresnel_sum_sharded(t["x"], t["y"], t["z"], t["sx"], t["sy"], t["sz"], t["u"], c["k"], t["ds"], gather=False)
            b, n = wavecalc._split(M, world, rank)
            ok = ok and loc.shape[0] == n and torch.equal(loc, dev[b:b + n])
            msgs.append(f"M={M}:{'ok' if ok else 'BAD'}:{rel_l2(got, ref):.2e}")
        # through-focus stack (config C5): planes x pixels flattened and sharded over the ranks; PSF planes array_split
        G, P = 32, 3
        c = _case(G * G, seed=11)
        planes = 0.15 + np.linspace(-1e-3, 1e-3, P)
        stack = akb.fresnel_sum_planes(c["y"], c["z"], planes, c["sx"], c["sy"], c["sz"], c["u"], c["k"], c["ds"])
        ok = isinstance(stack, np.ndarray) and stack.shape == (P, G * G)
        for p in range(P):
            ref = oracle.fresnel_sum(np.full(G * G, planes[p]), c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], c["k"], c["ds"], nthreads=2)
            ok = ok and rel_l2(stack[p], ref) <= 1e-12
        res = akb.psf_stack(torch.as_tensor(stack).cuda(), (G, G), 13.5e-9, 1e-7, 0.3, pad_factor=2)
        b, n = wavecalc._split(P, world, rank)
        ok = ok and res["planes"] == list(range(b, b + n)) and tuple(res["I"].shape) == (n, 2 * G, 2 * G)
        msgs.append(f"throughfocus:{'ok' if ok else 'BAD'}")
        # a communicator made through the C-ABI instead of PyTorch's
        wavecalc._own_comms.clear()
        os.environ["AKB_OWN_NCCL_COMM"] = "1"
        c = _case(4099, seed=3)
        ref = oracle.fresnel_sum(c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], c["k"], c["ds"], nthreads=2)
        got = akb.fresnel_sum_sharded(c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], c["k"], c["ds"])
        msgs.append(f"owncomm:{'ok' if (len(wavecalc._own_comms) == 1 and rel_l2(got, ref) <= 1e-12) else 'BAD'}")
    except Exception as e:  # noqa: BLE001
        import traceback
        msgs.append("EXC " + traceback.format_exc())
    finally:
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as fh:
            fh.write("\n".join(msgs))
        dist.destroy_process_group()


def test_torch_distributed_nccl_sharded(tmp_path):
    _need_two()
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_nccl_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        text = open(tmp_path / f"rank{r}.txt").read()
        print(f"rank {r}: {text}")
        assert text and "BAD" not in text and "EXC" not in text, text


# ------------------------------------------------------------------ 3. non-torch caller: a host thread per GPU, ctypes only

def test_c_abi_sharded_from_host_threads():
    _need_two()
    import torch  # only to allocate / copy device buffers in this test; the calls under test are pure C-ABI
    from akbraytracing_b200 import _lib
    L = _lib.load()
    world = min(torch.cuda.device_count(), 4)
    M = 10_001
    c = _case(M, seed=5)
    ref = oracle.fresnel_sum(c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], c["k"], c["ds"])
    uid = ctypes.create_string_buffer(128)
    _lib.check(L.akb_nccl_unique_id(uid), "akb_nccl_unique_id")
    results, errors = [None] * world, []

    def work(rank):
        try:
            torch.cuda.set_device(rank)
            dev = torch.device("cuda", rank)
            comm = ctypes.c_void_p()
            _lib.check(L.akb_nccl_comm_init(ctypes.byref(comm), world, rank, uid), "akb_nccl_comm_init")
            t = {k: torch.as_tensor(v).to(dev) for k, v in c.items() if k != "k"}
            if rank:  # only rank 0 holds the back surface
                for key in ("sx", "sy", "sz", "u", "ds"):
                    t[key].zero_()
            out = torch.empty(M, dtype=torch.complex128, device=dev)
            st = torch.cuda.Stream(dev)
            p = _lib.dev_ptr
            rc = L.akb_fresnel_sum_sharded(comm, rank, world, p(t["x"]), p(t["y"]), p(t["z"]), M, p(t["sx"]), p(t["sy"]),
                                           p(t["sz"]), p(t["u"]), p(t["ds"]), len(c["sx"]), c["k"], p(out), 0, 1,
                                           ctypes.c_void_p(st.cuda_stream))
            _lib.check(rc, "akb_fresnel_sum_sharded")
            st.synchronize()
            results[rank] = out.cpu().numpy()
            _lib.check(L.akb_nccl_comm_destroy(comm), "akb_nccl_comm_destroy")
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    [t.start() for t in threads]
    [t.join(timeout=300) for t in threads]
    assert not errors, errors
    for r in range(world):
        assert results[r] is not None and np.array_equal(results[r], results[0])
    err = rel_l2(results[0], ref)
    print(f"C-ABI sharded over {world} host threads: rel-L2 vs oracle {err:.2e}")
    assert err <= 1e-12
