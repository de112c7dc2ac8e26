"""ctypes binding of libakb_b200.so (the C-ABI declared in include/akb_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, a RuntimeError is
raised.  PyTorch is used only as the owner of device buffers and streams.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libakb_b200.so")

PHASE_FAITHFUL = 0
PHASE_EXACT = 1
PHASE_REFERENCED = 2
FLAG_MISS, FLAG_ZERO_NORM, FLAG_MISS_MASK, NFLAGS = 0, 1, 2, 4
MAX_MIRRORS = 8

_c_i64 = ctypes.c_int64
_c_int = ctypes.c_int
_c_uint = ctypes.c_uint
_c_dbl = ctypes.c_double
_vp = ctypes.c_void_p

# name -> (restype, argtypes); one entry per symbol declared in include/akb_b200.h
SIGNATURES = {
    "akb_last_error": (ctypes.c_char_p, []),
    "akb_version": (_c_int, []),
    "akb_device_count": (_c_int, []),
    "akb_trim": (_c_int, [_c_int]),
    "akb_fresnel_sum": (_c_int, [_vp, _vp, _vp, _c_i64, _vp, _vp, _vp, _vp, _vp, _c_i64, _c_dbl, _vp, _c_int, _vp]),
    "akb_fresnel_sum_host": (_c_int, [_vp, _vp, _vp, _c_i64, _vp, _vp, _vp, _vp, _vp, _c_i64, _c_dbl, _vp, _c_int, _c_int]),
    "akb_fresnel_sum_planes": (_c_int, [_vp, _vp, _c_i64, _vp, _c_int, _vp, _vp, _vp, _vp, _vp, _c_i64, _c_dbl, _vp, _c_int, _vp]),
    "akb_fresnel_sum_sharded": (_c_int, [_vp, _c_int, _c_int, _vp, _vp, _vp, _c_i64, _vp, _vp, _vp, _vp, _vp, _c_i64, _c_dbl, _vp,
                                          _c_int, _c_int, _vp]),
    "akb_allgather_blocks": (_c_int, [_vp, _c_int, _c_int, _vp, _c_i64, _c_int, _vp]),
    "akb_nccl_unique_id": (_c_int, [_vp]),
    "akb_nccl_comm_init": (_c_int, [ctypes.POINTER(_vp), _c_int, _c_int, _vp]),
    "akb_nccl_comm_destroy": (_c_int, [_vp]),
    "akb_shard_range": (_c_int, [_c_i64, _c_int, _c_int, ctypes.POINTER(_c_i64), ctypes.POINTER(_c_i64)]),
    "akb_launch_count": (_c_i64, [_c_int]),
    "akb_fresnel_variant_name": (ctypes.c_char_p, []),
    "akb_fresnel_row_blocks": (_c_int, [ctypes.POINTER(ctypes.c_int64), _c_int]),
    "akb_fresnel_timing": (_c_int, [_c_int]),
    "akb_fresnel_last_timing": (_c_int, [ctypes.POINTER(_c_dbl), ctypes.POINTER(_c_dbl), ctypes.POINTER(_c_int),
                                          ctypes.POINTER(_c_i64), ctypes.POINTER(_c_int)]),
    "akb_mirr_ray_intersection": (_c_int, [_vp, _vp, _vp, _c_i64, _c_int, _vp, _vp, _vp]),
    "akb_norm_vector": (_c_int, [_vp, _vp, _c_i64, _vp, _c_uint, _vp, _vp]),
    "akb_reflect_ray": (_c_int, [_vp, _vp, _c_i64, _vp, _c_uint, _vp, _vp]),
    "akb_normalize_vector": (_c_int, [_vp, _c_i64, _vp, _c_uint, _vp, _vp]),
    "akb_plane_ray_intersection": (_c_int, [_vp, _vp, _vp, _c_i64, _vp, _vp]),
    "akb_intersect_reflect": (_c_int, [_vp, _vp, _vp, _c_i64, _c_int, _vp, _vp, _vp, _c_uint, _vp, _vp]),
    "akb_trace_chain": (_c_int, [_vp, _vp, _c_int, _vp, _vp, _vp, _c_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _c_uint, _vp, _vp]),
    "akb_wavefront_opl": (_c_int, [_vp, _vp, _vp, _c_int, _c_i64, _vp, _vp, _vp, _c_dbl, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "akb_trace_chain_batched": (_c_int, [_vp, _vp, _c_int, _vp, _c_int, _vp, _vp, _c_i64, _vp, _vp, _vp, _vp]),
    "akb_intersect_reflect_host":(_c_int, [_vp, _vp, _vp, _c_i64, _c_int, _vp, _vp, _vp, _vp, _c_int]),
    "akb_trace_chain_host": (_c_int, [_vp, _vp, _c_int, _vp, _vp, _vp, _c_i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _c_int]),
    "akb_calc_ds": (_c_int, [_vp, _c_i64, _c_i64, _vp, _vp]),
    "akb_opl_to_field": (_c_int, [_vp, _vp, _c_i64, _c_dbl, _vp, _vp]),
    "akb_fp64_peak_probe": (_c_int, [_c_int, ctypes.POINTER(_c_dbl), _vp]),
    "akb_hbm_copy_probe": (_c_int, [_c_i64, _c_int, ctypes.POINTER(_c_dbl), _vp]),
    "akb_selftest_sqrt": (_c_int, [_c_i64, _c_dbl, _c_dbl, ctypes.POINTER(_c_i64), ctypes.POINTER(_c_dbl), _vp]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load the CUDA library; fail loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m akbraytracing_b200.build` "
                "(nvcc, sm_100a). akbraytracing_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library drift apart
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().akb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def host_ptr(a: np.ndarray):
    return ctypes.c_void_p(a.ctypes.data)


def as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def as_c128(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.complex128)


def launch_count(reset: bool = False) -> int:
    return int(load().akb_launch_count(1 if reset else 0))


def device_count() -> int:
    n = load().akb_device_count()
    if n < 0:
        check(n, "akb_device_count")
    return n


def shard_range(total: int, nranks: int, rank: int):
    b, c = _c_i64(), _c_i64()
    check(load().akb_shard_range(int(total), int(nranks), int(rank), ctypes.byref(b), ctypes.byref(c)), "akb_shard_range")
    return int(b.value), int(c.value)


# ---- torch helpers (device buffers + streams only) -------------------------------------------

def is_torch(x) -> bool:
    mod = type(x).__module__
    return mod == "torch" or mod.startswith("torch.")


def is_cuda_array(x) -> bool:
    """A non-torch object exposing __cuda_array_interface__ (CuPy / Numba device arrays)."""
    return (not is_torch(x)) and hasattr(x, "__cuda_array_interface__")


def from_cuda_array(x):
    """Zero-copy torch view of a __cuda_array_interface__ object (same pointer, device taken from it)."""
    import torch
    return torch.as_tensor(x, device="cuda")


def torch_stream_ptr(device):
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def dev_ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def dev_f64(t, device=None):
    """A contiguous float64 CUDA tensor (no copy when it already is one)."""
    import torch
    if is_cuda_array(t):
        t = from_cuda_array(t)
    if not is_torch(t):
        t = torch.as_tensor(np.ascontiguousarray(t, dtype=np.float64))
    if device is None:
        device = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    return t.to(device=device, dtype=torch.float64).contiguous()


def to_host(t):
    """A CUDA tensor as a NumPy array.  Results of 1 MiB .. 1 GiB come back through page-locked memory from torch's
    caching host allocator: the copy runs at link speed and, after the first call, touches no fresh pages (a
    67 MB field into a new pageable array: 31 ms, this way: 3 ms; tools/e2e_breakdown.py).  The array keeps the
    block alive and returns it to the allocator when dropped."""
    import torch
    nbytes = t.numel() * t.element_size()
    if not t.is_cuda or nbytes < (1 << 20) or nbytes > (1 << 30):
        return t.detach().cpu().numpy()
    try:
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    except RuntimeError:  # no page-locked memory to be had (RLIMIT_MEMLOCK, container limits): an ordinary array
        return t.detach().cpu().numpy()
    host.copy_(t.detach())
    return host.numpy()


def dev_c128(t, device=None):
    import torch
    if is_cuda_array(t):
        t = from_cuda_array(t)
    if not is_torch(t):
        t = torch.as_tensor(np.ascontiguousarray(t, dtype=np.complex128))
    if device is None:
        device = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    return t.to(device=device, dtype=torch.complex128).contiguous()
