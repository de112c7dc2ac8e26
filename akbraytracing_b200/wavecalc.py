"""Path A host side: the reference's Wavecalc call signatures over the sm_100a pair-sum kernel.

Mirrors (same names, positional order, return shapes/dtypes):

* ``forward_propagation_numpy_batch``          Wavecalc_raytrace_fromData_CPU0402.py:87-124
* ``forward_propagation_cupy_batch``           Wavecalc_raytrace_fromData_GPU0402.py:139-201
* ``forward_propagation_cupy_batch_multi_gpu`` GPU0402:64-136, GPU0402_multi.py:123-229
* ``WaveField3D``                              CPU0402:17-52 / GPU0402:14-62

Array kinds: NumPy in -> NumPy out (host buffers, H2D/D2H inside the call, like the CPU script);
device arrays in -> torch CUDA tensor out (zero copy, asynchronous on torch's current stream, like
the CuPy scripts).  A device array is a torch CUDA tensor or ANY object with ``__cuda_array_interface__``
(the cp.ndarrays GPU0402:36-38 holds, Numba device arrays): its pointer is used in place, and the
returned torch tensor exposes the same interface, so ``cp.asarray(result)`` is free.  Work is ordered
on torch's current stream -- the legacy default stream unless the caller changed it, which is also
CuPy's default, so a CuPy caller's pending kernels are ordered before ours.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
import time

import numpy as np

from . import _lib
from ._lib import PHASE_EXACT, PHASE_FAITHFUL, PHASE_REFERENCED  # noqa: F401  (re-exported)

__all__ = ["WaveField3D", "forward_propagation_numpy_batch", "forward_propagation_cupy_batch",
           "forward_propagation_cupy_batch_multi_gpu", "fresnel_sum", "fresnel_sum_sharded",
           "PHASE_FAITHFUL", "PHASE_EXACT", "PHASE_REFERENCED"]


def _any_device(*arrays) -> bool:
    """True when any argument is a device-side array: a torch tensor or any object exposing
    ``__cuda_array_interface__`` (CuPy arrays, Numba device arrays: GPU0402:36-38 holds cp.ndarrays)."""
    return any(_lib.is_torch(a) or _lib.is_cuda_array(a) for a in arrays)


def _fresnel_host(x, y, z, sx, sy, sz, u, k, ds, mode, device=-1) -> np.ndarray:
    x, y, z = _lib.as_f64(x), _lib.as_f64(y), _lib.as_f64(z)
    sx, sy, sz = _lib.as_f64(sx), _lib.as_f64(sy), _lib.as_f64(sz)
    u = _lib.as_c128(u)
    if not (x.shape == y.shape == z.shape and x.ndim == 1):
        raise ValueError("x, y, z must be 1-D arrays of equal length")
    if not (sx.shape == sy.shape == sz.shape == u.shape and sx.ndim == 1):
        raise ValueError("u_back_x, u_back_y, u_back_z, u_back_u must be 1-D arrays of equal length")
    dsp = None
    if ds is not None:
        ds = np.broadcast_to(_lib.as_f64(ds), sx.shape)
        ds = np.ascontiguousarray(ds)
        dsp = _lib.host_ptr(ds)
    out = np.empty(x.shape[0], dtype=np.complex128)
    rc = _lib.load().akb_fresnel_sum_host(
        _lib.host_ptr(x), _lib.host_ptr(y), _lib.host_ptr(z), x.shape[0],
        _lib.host_ptr(sx), _lib.host_ptr(sy), _lib.host_ptr(sz), _lib.host_ptr(u), dsp, sx.shape[0],
        float(k), _lib.host_ptr(out), int(mode), int(device))
    _lib.check(rc, "akb_fresnel_sum_host")
    return out


def _fresnel_device(x, y, z, sx, sy, sz, u, k, ds, mode, device=None):
    import torch
    if device is None:
        for a in (x, y, z, sx, u):
            if _lib.is_cuda_array(a):
                a = _lib.from_cuda_array(a)
            if _lib.is_torch(a) and a.is_cuda:
                device = a.device
                break
        else:
            device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    x, y, z = (_lib.dev_f64(a, device) for a in (x, y, z))
    sx, sy, sz = (_lib.dev_f64(a, device) for a in (sx, sy, sz))
    u = _lib.dev_c128(u, device)
    if ds is not None:
        ds = _lib.dev_f64(ds, device)
        if ds.numel() != sx.numel():
            ds = ds.expand(sx.shape).contiguous()
    if not (x.shape == y.shape == z.shape and x.dim() == 1):
        raise ValueError("x, y, z must be 1-D arrays of equal length")
    if not (sx.shape == sy.shape == sz.shape == u.shape and sx.dim() == 1):
        raise ValueError("u_back_x, u_back_y, u_back_z, u_back_u must be 1-D arrays of equal length")
    out = torch.empty(x.shape[0], dtype=torch.complex128, device=device)
    with torch.cuda.device(device):
        rc = _lib.load().akb_fresnel_sum(
            _lib.dev_ptr(x), _lib.dev_ptr(y), _lib.dev_ptr(z), x.shape[0],
            _lib.dev_ptr(sx), _lib.dev_ptr(sy), _lib.dev_ptr(sz), _lib.dev_ptr(u),
            _lib.dev_ptr(ds) if ds is not None else None, sx.shape[0], float(k), _lib.dev_ptr(out), int(mode),
            _lib.torch_stream_ptr(device))
    _lib.check(rc, "akb_fresnel_sum")
    # inputs may be temporaries: the launch is stream ordered and torch's caching allocator
    # only reuses their memory on the same stream, so they stay valid for the kernel.
    return out


def fresnel_sum(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds=None, mode=PHASE_FAITHFUL, device=None):
    """u_i = sum_j (u_j ds_j) exp(-1j k r_ij)/r_ij on one B200 (CPU0402:71-85 + :102)."""
    if _any_device(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, ds):
        return _fresnel_device(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds, mode, device)
    return _fresnel_host(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds, mode, -1 if device is None else device)


def compute_u_parallel(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k):
    """Drop-in for the numba kernel CPU0402:71-85: the pair sum with the weights as given (no ``ds``)."""
    return fresnel_sum(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k)


def compute_u(i, x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k):
    """Drop-in for CPU0402:54-63: the field at detector point ``i`` alone (a complex scalar for NumPy input)."""
    u = fresnel_sum(x[i:i + 1] if i != -1 else x[-1:], y[i:i + 1] if i != -1 else y[-1:],
                    z[i:i + 1] if i != -1 else z[-1:], u_back_x, u_back_y, u_back_z, u_back_u, k)
    return u[0]


def forward_propagation_numpy_batch(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds, num_cores=None):
    """Drop-in for CPU0402:87-124.  ``num_cores`` is accepted and ignored (it is ineffective in
    the reference too: the env var is written after numba is imported, CPU0402:105-107)."""
    del num_cores
    return fresnel_sum(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds)


def forward_propagation_cupy_batch(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds):
    """Drop-in for GPU0402:139-201 (single device).  No batching: nothing is materialised."""
    return fresnel_sum(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds)


# ---------------------------------------------------------------- multi-GPU (detector sharding)

def _split(total, parts, rank):
    """array_split block of `rank`: (begin, count) -- np.array_split / cp.array_split (GPU0402:77-79)."""
    base, extra = divmod(int(total), int(parts))
    return rank * base + min(rank, extra), base + (1 if rank < extra else 0)


def _gather_blocks(local, total, group):
    """All-gather of array_split blocks through torch.distributed on whatever device `local` lives on
    (equal-size collective: every block is padded to the largest).  Used where NCCL cannot carry the
    exchange (gloo groups); the NCCL path gathers in place inside akb_fresnel_sum_sharded."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if not _lib.is_torch(local):
        local = torch.as_tensor(np.ascontiguousarray(local))
    count = int(local.shape[0])
    width = -(-total // world) if total else 0
    send = torch.zeros(width, dtype=local.dtype, device=local.device)
    send[:count] = local
    recv = torch.empty(world * width, dtype=local.dtype, device=local.device)
    if width:
        dist.all_gather_into_tensor(recv.view(torch.float64) if recv.is_cuda else recv,
                                    send.view(torch.float64) if send.is_cuda else send, group=group)
    pieces = [recv[r * width:r * width + _split(total, world, r)[1]] for r in range(world)]
    return torch.cat(pieces) if pieces else recv


def _sharded_over_group(compute_local, x, y, z, rest, group, gather):
    """Host logic of the torch.distributed form: this rank's array_split block of the detector points goes to
    `compute_local`, the blocks are gathered with _gather_blocks.  (tests/test_sharded_gloo.py drives this
    with a CPU stand-in for `compute_local`; the product passes the CUDA call.)"""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    total = int(x.shape[0])
    begin, count = _split(total, world, rank)
    sl = slice(begin, begin + count)
    local = compute_local(x[sl], y[sl], z[sl], *rest)
    return _gather_blocks(local, total, group) if gather else local


_own_comms = {}  # id(group) -> ncclComm_t made by akb_nccl_comm_init (when PyTorch's cannot be borrowed)


def _nccl_comm(group, device):
    """An ncclComm_t spanning `group` on this rank's device: PyTorch's own communicator when the group runs on
    NCCL (ProcessGroupNCCL._comm_ptr), else one created through the C-ABI from a unique id that rank 0
    hands out over the group."""
    import ctypes
    import torch
    import torch.distributed as dist
    key = id(group) if group is not None else 0
    if key in _own_comms:
        return _own_comms[key]
    pg = group if group is not None else dist.group.WORLD
    try:
        if os.environ.get("AKB_OWN_NCCL_COMM") == "1":
            raise RuntimeError("own communicator requested")
        backend = pg._get_backend(torch.device(device))
        if hasattr(backend, "_comm_ptr"):
            ptr = int(backend._comm_ptr())
            if ptr:
                return ptr
    except Exception:  # noqa: BLE001 -- not an NCCL group, or a PyTorch without _comm_ptr: make our own
        pass
    L = _lib.load()
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    uid = ctypes.create_string_buffer(128)
    if rank == 0:
        _lib.check(L.akb_nccl_unique_id(uid), "akb_nccl_unique_id")
    box = [uid.raw]
    dist.broadcast_object_list(box, src=dist.get_global_rank(pg, 0) if group is not None else 0, group=group)
    uid = ctypes.create_string_buffer(box[0], 128)
    comm = ctypes.c_void_p()
    with torch.cuda.device(device):
        _lib.check(L.akb_nccl_comm_init(ctypes.byref(comm), world, rank, uid), "akb_nccl_comm_init")
    _own_comms[key] = comm.value
    return comm.value


def fresnel_sum_sharded(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds=None, mode=PHASE_FAITHFUL,
                        group=None, gather=True, device=None, broadcast_sources=False):
    """One process per GPU (GPU0402_multi.py:123-229 with processes instead of host threads): rank r computes the
    r-th ``array_split`` block of the detector points (GPU0402:77-79) against the full source set and the blocks are
    all-gathered (the NCCL replacement of ``cp.concatenate``, GPU0402:135).  Every rank passes the same full
    arrays and every rank returns the full field -- NumPy in, NumPy out; device arrays in, a torch CUDA tensor out.

    On an NCCL-capable group the whole call is ``akb_fresnel_sum_sharded`` (block kernel + in-place NCCL
    all-gather on the current stream).  On a group that cannot move device memory (gloo) the block is
    computed on this rank's GPU and gathered through the group.  ``gather=False`` returns the local block only.
    ``broadcast_sources``: take the source arrays from rank 0 (the reference keeps them on device 0)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("fresnel_sum_sharded needs an initialised torch.distributed process group")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    arrays = (x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, ds)
    was_numpy = not _any_device(*arrays)
    if device is None:
        device = next((a.device for a in arrays if _lib.is_torch(a) and a.is_cuda), None)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    total = int(x.shape[0])
    rest = (u_back_x, u_back_y, u_back_z, u_back_u, k, ds)
    backend = dist.get_backend(group)
    if "nccl" not in str(backend) or not gather:
        def compute_local(xs, ys, zs, *r):
            return fresnel_sum(xs, ys, zs, *r, mode=mode, device=device)
        out = _sharded_over_group(compute_local, x, y, z, rest, group, gather)
        return _lib.to_host(out) if (was_numpy and _lib.is_torch(out)) else out
    begin = 0
    if not _any_device(x, y, z):
        # host detector arrays: only this rank's block crosses PCIe (1/world of the coordinates); the entry point
        # indexes det[begin + i], so it gets the block's address moved back by `begin` elements
        x, y, z = (_lib.as_f64(a) for a in (x, y, z))
        if not (x.shape == y.shape == z.shape and x.ndim == 1):
            raise ValueError("x, y, z must be 1-D arrays of equal length")
        begin, count = _lib.shard_range(total, world, rank)
        x, y, z = (a[begin:begin + count] for a in (x, y, z))
    dx, dy, dz = (_lib.dev_f64(a, device) for a in (x, y, z))
    sx, sy, sz = (_lib.dev_f64(a, device) for a in (u_back_x, u_back_y, u_back_z))
    su = _lib.dev_c128(u_back_u, device)
    if not (dx.shape == dy.shape == dz.shape and dx.dim() == 1):
        raise ValueError("x, y, z must be 1-D arrays of equal length")
    if not (sx.shape == sy.shape == sz.shape == su.shape and sx.dim() == 1):
        raise ValueError("u_back_x, u_back_y, u_back_z, u_back_u must be 1-D arrays of equal length")
    sd = None
    if ds is not None:
        sd = _lib.dev_f64(ds, device)
        if sd.numel() != sx.numel():
            sd = sd.expand(sx.shape).contiguous()
    if broadcast_sources:  # the buffers are written by the broadcast: never alias the caller's arrays on ranks > 0
        sx, sy, sz, su = (t.clone() if rank else t for t in (sx, sy, sz, su))
        sd = sd.clone() if (sd is not None and rank) else sd
    out = torch.empty(total, dtype=torch.complex128, device=device)
    comm = _nccl_comm(group, device)
    with torch.cuda.device(device):
        rc = _lib.load().akb_fresnel_sum_sharded(
            comm, rank, world, *(ctypes.c_void_p(t.data_ptr() - 8 * begin) for t in (dx, dy, dz)), total, _lib.dev_ptr(sx), _lib.dev_ptr(sy), _lib.dev_ptr(sz), _lib.dev_ptr(su),
            _lib.dev_ptr(sd) if sd is not None else None, sx.shape[0], float(k), _lib.dev_ptr(out), int(mode),
            1 if broadcast_sources else 0, _lib.torch_stream_ptr(device))
    _lib.check(rc, "akb_fresnel_sum_sharded")
    return _lib.to_host(out) if was_numpy else out


def forward_propagation_cupy_batch_multi_gpu(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds, devices=None,
                                             mode=PHASE_FAITHFUL):
    """Drop-in for GPU0402:64-136 / GPU0402_multi.py:123-229.

    * under torch.distributed (one process per GPU, world > 1): ``fresnel_sum_sharded`` -- detector sharding,
      NCCL all-gather, the full field on every rank;
    * otherwise, from ONE process: the detector blocks of ``array_split`` are launched on every visible
      device back to back (the entry points are asynchronous, so no host thread per GPU is needed as in
      GPU0402_multi.py:213-225) and concatenated on the first device (GPU0402:135)."""
    import torch
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return fresnel_sum_sharded(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds, mode=mode)
    except ImportError:  # pragma: no cover
        pass
    was_numpy = not _any_device(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, ds)
    if devices is None:
        devices = list(range(_lib.device_count()))
    if len(devices) <= 1:
        return fresnel_sum(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds, mode=mode)
    total = int(x.shape[0])
    home = torch.device("cuda", devices[0])
    parts = []
    for r, d in enumerate(devices):
        dev = torch.device("cuda", d)
        b, c = _lib.shard_range(total, len(devices), r)
        sl = slice(b, b + c)
        parts.append(_fresnel_device(x[sl], y[sl], z[sl], u_back_x, u_back_y, u_back_z, u_back_u, k, ds, mode, dev))
    out = torch.cat([p.to(home, non_blocking=True) for p in parts])
    if was_numpy:
        return _lib.to_host(out)
    return out


# ---------------------------------------------------------------- the holder class

class WaveField3D:
    """Field samples on one surface (CPU0402:17-52 / GPU0402:14-62).

    ``device=None`` keeps NumPy arrays (CPU script flavour); ``device='cuda'`` keeps torch CUDA
    tensors resident between stages (CuPy script flavour), so a stage chain
    src -> M1 -> ... -> Image never leaves HBM.
    """

    def __init__(self, num, _lambda, wave_num_H, wave_num_V, device=None):
        self._device = device
        self.phase_mode = PHASE_FAITHFUL  # arithmetic of forward_propagation (the reference's roundings by default)
        if device is None:
            self.u = np.zeros(num, dtype=np.complex128)
            self.x = np.zeros(num, dtype=np.float64)
            self.y = np.zeros(num, dtype=np.float64)
            self.z = np.zeros(num, dtype=np.float64)
        else:
            import torch
            self.u = torch.zeros(num, dtype=torch.complex128, device=device)
            self.x = torch.zeros(num, dtype=torch.float64, device=device)
            self.y = torch.zeros(num, dtype=torch.float64, device=device)
            self.z = torch.zeros(num, dtype=torch.float64, device=device)
        self.lambda_ = np.float64(_lambda)
        self.wave_num_H = wave_num_H
        self.wave_num_V = wave_num_V

    def _own(self, row):
        if self._device is None:
            if _lib.is_torch(row):
                row = row.detach().cpu().numpy()
            return np.array(row, dtype=np.float64)
        return _lib.dev_f64(row, self._device).clone()

    def setdata(self, data):
        self.x = self._own(data[0, :])
        self.y = self._own(data[1, :])
        self.z = self._own(data[2, :])

    def set_ds(self, data):
        self.ds = self._own(data)

    def forward_propagation(self, u_back, num_cores=None):
        k = 2.0 * np.pi / self.lambda_  # CPU0402:39
        t0 = time.time()
        if self._device is None:
            if self.phase_mode == PHASE_FAITHFUL:
                self.u = forward_propagation_numpy_batch(self.x, self.y, self.z, u_back.x, u_back.y, u_back.z,
                                                         u_back.u, k, u_back.ds, num_cores=num_cores)
            else:
                self.u = fresnel_sum(self.x, self.y, self.z, u_back.x, u_back.y, u_back.z, u_back.u, k, u_back.ds,
                                     mode=self.phase_mode)
        else:
            self.u = forward_propagation_cupy_batch_multi_gpu(self.x, self.y, self.z, u_back.x, u_back.y,
                                                              u_back.z, u_back.u, k, u_back.ds, mode=self.phase_mode)
        print(f"計算時間: {time.time() - t0:.6f} 秒")
