"""Path A host side: the reference's Wavecalc call signatures over the sm_100a pair-sum kernel.

Mirrors (same names, positional order, return shapes/dtypes):

* ``forward_propagation_numpy_batch``          Wavecalc_raytrace_fromData_CPU0402.py:87-124
* ``forward_propagation_cupy_batch``           Wavecalc_raytrace_fromData_GPU0402.py:139-201
* ``forward_propagation_cupy_batch_multi_gpu`` GPU0402:64-136, GPU0402_multi.py:123-229
* ``WaveField3D``                              CPU0402:17-52 / GPU0402:14-62

Array kinds: NumPy in -> NumPy out (host buffers, H2D/D2H inside the call, like the CPU script);
torch CUDA tensors in -> torch CUDA tensor out (zero copy, asynchronous on the current stream,
like the CuPy scripts).  There is no CPU fallback.
"""
from __future__ import annotations

import time

import numpy as np

from . import _lib
from ._lib import PHASE_EXACT, PHASE_FAITHFUL, PHASE_REFERENCED  # noqa: F401  (re-exported)

__all__ = ["WaveField3D", "forward_propagation_numpy_batch", "forward_propagation_cupy_batch",
           "forward_propagation_cupy_batch_multi_gpu", "fresnel_sum", "fresnel_sum_sharded",
           "PHASE_FAITHFUL", "PHASE_EXACT", "PHASE_REFERENCED"]


def _any_torch(*arrays) -> bool:
    return any(_lib.is_torch(a) for a in arrays)


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
    if _any_torch(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, ds):
        return _fresnel_device(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds, mode, device)
    return _fresnel_host(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds, mode, -1 if device is None else device)


def forward_propagation_numpy_batch(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds, num_cores=None):
    """Drop-in for CPU0402:87-124.  ``num_cores`` is accepted and ignored (it is ineffective in
    the reference too: the env var is written after numba is imported, CPU0402:105-107)."""
    del num_cores
    return fresnel_sum(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds)


def forward_propagation_cupy_batch(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds):
    """Drop-in for GPU0402:139-201 (single device).  No batching: nothing is materialised."""
    return fresnel_sum(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds)


# ---------------------------------------------------------------- multi-GPU (detector sharding)

def fresnel_sum_sharded(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds=None, mode=PHASE_FAITHFUL,
                        group=None, compute=None, gather=True):
    """One process per GPU: rank r computes the r-th ``array_split`` block of detector points
    (GPU0402:77-79) against the full source set and the blocks are all-gathered (the NCCL
    replacement of ``cp.concatenate``, GPU0402:135).  Every rank returns the full field.

    ``compute`` exists for CPU tests of this host logic (gloo): it replaces the CUDA call.
    """
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("fresnel_sum_sharded needs an initialised torch.distributed process group")
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    total = int(x.shape[0])
    begin, count = _lib.shard_range(total, world, rank) if compute is None else _split(total, world, rank)
    sl = slice(begin, begin + count)
    fn = compute if compute is not None else (lambda *a: fresnel_sum(*a, mode=mode))
    local = fn(x[sl], y[sl], z[sl], u_back_x, u_back_y, u_back_z, u_back_u, k, ds)
    if not gather:
        return local
    if not _lib.is_torch(local):
        local = torch.as_tensor(np.ascontiguousarray(local))
    # equal-size all-gather: pad every block to the largest (first total%world blocks hold +1)
    width = -(-total // world) if total else 0
    send = torch.zeros(width, dtype=local.dtype, device=local.device)
    send[:count] = local
    recv = torch.empty(world * width, dtype=local.dtype, device=local.device)
    if width:
        dist.all_gather_into_tensor(recv.view(torch.float64) if recv.is_cuda else recv,
                                    send.view(torch.float64) if send.is_cuda else send, group=group)
    pieces = []
    for r in range(world):
        _, c = _split(total, world, r)
        pieces.append(recv[r * width:r * width + c])
    return torch.cat(pieces) if pieces else recv


def _split(total, parts, rank):
    base, extra = divmod(int(total), int(parts))
    return rank * base + min(rank, extra), base + (1 if rank < extra else 0)


def forward_propagation_cupy_batch_multi_gpu(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds, devices=None):
    """Drop-in for GPU0402:64-136 / GPU0402_multi.py:123-229.

    * under torch.distributed (one process per GPU, world > 1): detector sharding + all-gather;
    * otherwise, from ONE process: the detector blocks of ``array_split`` are launched on every
      visible device back to back (the entry points are asynchronous, so no host thread per GPU
      is needed as in GPU0402_multi.py:213-225) and gathered on the first device.
    """
    import torch
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return fresnel_sum_sharded(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds)
    except ImportError:  # pragma: no cover
        pass
    was_numpy = not _any_torch(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, ds)
    if devices is None:
        devices = list(range(_lib.device_count()))
    if len(devices) <= 1:
        return fresnel_sum(x, y, z, u_back_x, u_back_y, u_back_z, u_back_u, k, ds)
    total = int(x.shape[0])
    home = torch.device("cuda", devices[0])
    parts = []
    for r, d in enumerate(devices):
        dev = torch.device("cuda", d)
        b, c = _lib.shard_range(total, len(devices), r)
        sl = slice(b, b + c)
        parts.append(_fresnel_device(x[sl], y[sl], z[sl], u_back_x, u_back_y, u_back_z, u_back_u, k, ds,
                                     PHASE_FAITHFUL, dev))
    out = torch.cat([p.to(home, non_blocking=True) for p in parts])
    if was_numpy:
        return out.cpu().numpy()
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
            self.u = forward_propagation_numpy_batch(self.x, self.y, self.z, u_back.x, u_back.y, u_back.z,
                                                     u_back.u, k, u_back.ds, num_cores=num_cores)
        else:
            self.u = forward_propagation_cupy_batch_multi_gpu(self.x, self.y, self.z, u_back.x, u_back.y,
                                                              u_back.z, u_back.u, k, u_back.ds)
        print(f"計算時間: {time.time() - t0:.6f} 秒")
