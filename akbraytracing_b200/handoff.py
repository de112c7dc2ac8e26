"""Hand-off helpers between the ray path and the field path.

* ``calc_dS(points, ray_num_V, ray_num_H)``  AKB_raytrace_20250312.py:13418-13473 -- the O(N)
  Python double loop of the reference as one kernel launch;
* ``opl_to_field(opl, k, amp=None)``: ``amp * exp(-1j*k*opl)``, the field a traced wavefront
  carries onto the last mirror, using the same exact phase reduction as the pair-sum kernel.
"""
from __future__ import annotations

import numpy as np

from . import _lib

__all__ = ["calc_dS", "opl_to_field"]


def _device_of(*arrays):
    import torch
    for a in arrays:
        if _lib.is_torch(a) and a.is_cuda:
            return a.device
    return torch.device("cuda", torch.cuda.current_device())


def calc_dS(points, ray_num_V, ray_num_H):
    """Area element per point of a (3, nV*nH) cloud sampled on an nV x nH grid -> (nV, nH)."""
    import torch
    numpy_io = not _lib.is_torch(points)
    dev = _device_of(points)
    p = _lib.dev_f64(points, dev)
    nV, nH = int(ray_num_V), int(ray_num_H)
    if p.dim() != 2 or p.shape[0] < 3 or p.shape[1] != nV * nH:
        raise ValueError("points must have shape (3, ray_num_V*ray_num_H)")
    p = p[:3].contiguous()
    out = torch.empty(nV, nH, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().akb_calc_ds(_lib.dev_ptr(p), nV, nH, _lib.dev_ptr(out), _lib.torch_stream_ptr(dev))
    _lib.check(rc, "akb_calc_ds")
    return _lib.to_host(out) if numpy_io else out


def opl_to_field(opl, k, amp=None):
    """complex128[N] = amp * exp(-1j * k * opl)."""
    import torch
    numpy_io = not (_lib.is_torch(opl) or _lib.is_torch(amp))
    dev = _device_of(opl, amp)
    o = _lib.dev_f64(opl, dev).reshape(-1)
    a = _lib.dev_f64(amp, dev).reshape(-1) if amp is not None else None
    out = torch.empty(o.shape[0], dtype=torch.complex128, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.load().akb_opl_to_field(_lib.dev_ptr(o), _lib.dev_ptr(a) if a is not None else None, o.shape[0],
                                          float(k), _lib.dev_ptr(out), _lib.torch_stream_ptr(dev))
    _lib.check(rc, "akb_opl_to_field")
    return _lib.to_host(out) if numpy_io else out
