"""Through-focus stack (BASELINE config C5): the field of one source surface on P detector
planes x = x_p that share one (y, z) pixel grid, and the Fraunhofer PSF of every plane.

The reference computes one plane per run of the Wavecalc script (focal grid and one defocused grid,
Wavecalc_raytrace_fromData_CPU0402.py:330-370) and one PSF per call of ``compute_psf_fft`` (psf_fft.py:29-125, fed by
``psf_calc``, AKB_raytrace_20250312.py:1121-1200 with ``pad_factor=16``).  Here

* ``fresnel_sum_planes`` is one ``akb_fresnel_sum_planes`` launch over the plane-major flat detector set: every
  1024-point block of the pair kernel lies in one row of one plane and takes the planar-row loop (REFERENCED mode: with
  the row expansion).  Under torch.distributed the planes are split like ``array_split`` over the ranks (4 planes per
  GPU at C5 on 8 GPUs) and all-gathered in place.  (A kernel in which a thread keeps one pixel for
  four planes -- y and z terms of r^2 shared by four pairs, 24 instead of 25.5 FP64 instructions per pair -- was built
  and measured 8 % slower: its ten row loads per loop iteration stall on shared memory)
* ``psf_stack`` evaluates ``compute_psf_fft`` for a batch of planes with one batched ``torch.fft.fft2`` (a library
  call: north_star item 4 keeps the PSF out of the optimisation scope), plane p on rank p's ``array_split`` block.

Memory of the PSF step: the padded pupil of one plane is (pad_factor * n)^2 complex128 -- 64 MiB for n = 1024 at
pad_factor = 2, 4 GiB at the reference's pad_factor = 16 -- and about four such buffers are alive at the peak
(padded pupil, its shifted copy, the transform, the intensity), so ``psf_stack`` works through the planes in chunks
sized from the free device memory: at pad_factor = 16 a 180 GB B200 holds about 8 planes per chunk, and 32 planes on
8 GPUs are 4 planes = one chunk per GPU.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .psf import field_to_pupil
from .wavecalc import PHASE_FAITHFUL, _any_device, fresnel_sum_sharded

__all__ = ["fresnel_sum_planes", "psf_stack", "compute_psf_fft_batch"]


def _dist_world():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(), dist.get_rank()
    except ImportError:  # pragma: no cover
        pass
    return 1, 0


def fresnel_sum_planes(y, z, x_planes, u_back_x, u_back_y, u_back_z, u_back_u, k, ds=None, mode=PHASE_FAITHFUL,
                       device=None, group=None):
    """Field of the back surface on the planes x = x_planes[p], all sampled at the same pixels (y[i], z[i]).

    y, z: float64[M] (a meshgrid-ordered focal grid: y fastest, like ``np.meshgrid(y_grid, z_grid)`` flattened,
    AKB_raytrace_20250312.py:13581-13589); x_planes: float64[P].  Returns complex128 (P, M): NumPy for NumPy inputs, a
    torch CUDA tensor for device inputs.

    One ``akb_fresnel_sum_planes`` call (the plane-major flat detector set through the planar-row loop; REFERENCED mode:
    with the row expansion).  With an initialised
    torch.distributed NCCL group of more than one rank the PLANES are split like ``array_split`` over the ranks (fewer
    planes than ranks: the flat (plane, pixel) index instead) and all-gathered: every rank returns the full stack."""
    import torch
    arrays = (y, z, u_back_x, u_back_y, u_back_z, u_back_u, ds)
    was_numpy = not _any_device(*arrays)
    if device is None:
        device = next((a.device for a in arrays if _lib.is_torch(a) and a.is_cuda), None)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    yd, zd = _lib.dev_f64(y, device).reshape(-1), _lib.dev_f64(z, device).reshape(-1)
    if yd.shape != zd.shape:
        raise ValueError("y and z must have the same length")
    xp = torch.as_tensor(np.ascontiguousarray(np.asarray(x_planes, dtype=np.float64).reshape(-1)), device=device) \
        if not _lib.is_torch(x_planes) else x_planes.to(device=device, dtype=torch.float64).reshape(-1).contiguous()
    P, M = int(xp.shape[0]), int(yd.shape[0])
    world, rank = _dist_world()
    nccl = False
    if world > 1:
        import torch.distributed as dist
        nccl = "nccl" in str(dist.get_backend(group))
    if world > 1 and (P < world or not nccl):  # too few planes to give every rank one, or no NCCL: shard the flat index
        gx = xp.repeat_interleave(M)           # plane-major flat detector set: (p, i) -> p*M + i
        flat = fresnel_sum_sharded(gx, yd.repeat(P), zd.repeat(P), u_back_x, u_back_y, u_back_z, u_back_u, k, ds, mode=mode,
                                   device=device, group=group)
        out = flat.reshape(P, M)
        return _lib.to_host(out) if was_numpy else out
    sx, sy, sz = (_lib.dev_f64(a, device) for a in (u_back_x, u_back_y, u_back_z))
    su = _lib.dev_c128(u_back_u, device)
    if not (sx.shape == sy.shape == sz.shape == su.shape and sx.dim() == 1):
        raise ValueError("u_back_x, u_back_y, u_back_z, u_back_u must be 1-D arrays of equal length")
    sd = None
    if ds is not None:
        sd = _lib.dev_f64(ds, device)
        if sd.numel() != sx.numel():
            sd = sd.expand(sx.shape).contiguous()
    out = torch.empty(P, M, dtype=torch.complex128, device=device)
    b, c = _lib.shard_range(P, world, rank) if world > 1 else (0, P)
    L = _lib.load()
    with torch.cuda.device(device):
        st = _lib.torch_stream_ptr(device)
        if c > 0:
            rc = L.akb_fresnel_sum_planes(_lib.dev_ptr(yd), _lib.dev_ptr(zd), M, _lib.dev_ptr(xp[b:b + c]), c, _lib.dev_ptr(sx),
                                          _lib.dev_ptr(sy), _lib.dev_ptr(sz), _lib.dev_ptr(su),
                                          _lib.dev_ptr(sd) if sd is not None else None, sx.shape[0], float(k),
                                          _lib.dev_ptr(out[b:b + c]), int(mode), st)
            _lib.check(rc, "akb_fresnel_sum_planes")
        if world > 1:  # every plane is one item of 2*M doubles: in-place all-gather of the array_split blocks of planes
            from .wavecalc import _nccl_comm
            rc = L.akb_allgather_blocks(_nccl_comm(group, device), rank, world, _lib.dev_ptr(out), P, 2 * M, st)
            _lib.check(rc, "akb_allgather_blocks")
    return _lib.to_host(out) if was_numpy else out


def compute_psf_fft_batch(opd_m, amp, wavelength_m, pupil_dx_m, focal_length_m, pad_factor=2, window=None,
                          return_efield=False, pupil_dy_m=None, device="cuda"):
    """``compute_psf_fft`` (psf_fft.py:29-125) for a stack: opd_m, amp of shape (P, ny, nx); one batched fft2.
    Every plane is normalised by its own peak, exactly as P separate calls would be.
    Returns (I (P, py, px), x_im, y_im[, E])."""
    import torch
    from .psf import _hann2d
    numpy_io = not (_lib.is_torch(opd_m) or _lib.is_torch(amp))
    if tuple(opd_m.shape) != tuple(amp.shape) or len(opd_m.shape) != 3:
        raise ValueError("opd_m and amp must both have shape (P, ny, nx)")
    if pad_factor < 1 or int(pad_factor) != pad_factor:
        raise ValueError("pad_factor must be a positive integer")
    pad_factor = int(pad_factor)
    if _lib.is_torch(opd_m):
        device = opd_m.device
    dev = torch.device(device)
    opd = torch.as_tensor(opd_m, dtype=torch.float64).to(dev)
    A = torch.as_tensor(amp, dtype=torch.float64).to(dev)
    A = torch.where(torch.isfinite(A), A, torch.zeros_like(A))          # PSF:86-87
    opd = torch.where(torch.isfinite(opd), opd, torch.zeros_like(opd))
    U_p = torch.polar(A, (2.0 * np.pi / wavelength_m) * opd)            # PSF:89-90
    if window is not None:
        if str(window).lower() != "hann":
            raise ValueError(f"Unsupported window '{window}'. Options: 'hann' or None.")
        U_p = U_p * _hann2d(U_p.shape[1], U_p.shape[2], dev)
    P, ny, nx = U_p.shape
    if ny % 2 or nx % 2:                                                # ensure_even_size, PSF:6-18
        U_p = torch.nn.functional.pad(U_p, (0, nx % 2, 0, ny % 2))
        P, ny, nx = U_p.shape
    py, px = ny * pad_factor, nx * pad_factor
    oy, ox = (py - ny) // 2, (px - nx) // 2
    U_pad = torch.zeros(P, py, px, dtype=torch.complex128, device=dev)
    U_pad[:, oy:oy + ny, ox:ox + nx] = U_p
    dx = pupil_dx_m
    dy = dx if pupil_dy_m is None else pupil_dy_m
    dims = (-2, -1)
    U_im = torch.fft.fftshift(torch.fft.fft2(torch.fft.ifftshift(U_pad, dim=dims), dim=dims), dim=dims) * (dx * dy)  # PSF:110
    del U_pad
    x_im = wavelength_m * focal_length_m * np.fft.fftshift(np.fft.fftfreq(px, d=dx))    # PSF:112-115
    y_im = wavelength_m * focal_length_m * np.fft.fftshift(np.fft.fftfreq(py, d=dy))
    inten = U_im.real ** 2 + U_im.imag ** 2
    peak = inten.amax(dim=dims, keepdim=True)
    safe = torch.where(peak > 0, peak, torch.ones_like(peak))
    inten = inten / safe
    ef = U_im / torch.sqrt(safe) if return_efield else None
    if numpy_io:
        inten = _lib.to_host(inten)
        ef = _lib.to_host(ef) if ef is not None else None
    else:
        x_im, y_im = torch.as_tensor(x_im, device=dev), torch.as_tensor(y_im, device=dev)
    return (inten, x_im, y_im, ef) if return_efield else (inten, x_im, y_im)


def psf_stack(fields, grid_shape, wavelength_m, pupil_dx_m, focal_length_m, pad_factor=2, window=None,
              pupil_dy_m=None, planes_per_chunk=None, shard=True):
    """PSF of every plane of a through-focus stack.

    fields: complex128 (P, M) as returned by ``fresnel_sum_planes`` (torch CUDA tensor or NumPy), M = ny*nx with
    grid_shape = (ny, nx).  Each plane becomes the pupil (amp = |u|, opd = arg(u) lambda / 2 pi: ``field_to_pupil``,
    SURVEY.md D5) of ``compute_psf_fft``.  Under torch.distributed with ``shard=True`` rank r transforms the r-th
    ``array_split`` block of the planes.  Returns dict(planes = indices handled here, I (len(planes), py, px),
    x, y): NumPy for NumPy input, torch otherwise."""
    import torch
    numpy_io = not _lib.is_torch(fields)
    f = torch.as_tensor(fields)
    if not f.is_cuda:
        f = f.to(torch.device("cuda", torch.cuda.current_device()))
    ny, nx = (int(v) for v in grid_shape)
    P = int(f.shape[0])
    if f.shape[1] != ny * nx:
        raise ValueError("fields must have shape (P, ny*nx)")
    world, rank = _dist_world()
    b, c = _lib.shard_range(P, world, rank) if (shard and world > 1) else (0, P)
    mine = list(range(b, b + c))
    py, px = (ny + ny % 2) * int(pad_factor), (nx + nx % 2) * int(pad_factor)
    if planes_per_chunk is None:  # ~4 padded complex128 buffers alive per plane at the peak
        free, _ = torch.cuda.mem_get_info(f.device)
        planes_per_chunk = max(1, int(free * 0.8 // (4 * 16 * py * px)))
    out_I, x_im, y_im = [], None, None
    for s in range(0, len(mine), planes_per_chunk):
        idx = mine[s:s + planes_per_chunk]
        opd, amp = field_to_pupil(f[idx].reshape(len(idx), ny, nx), wavelength_m)
        inten, x_im, y_im = compute_psf_fft_batch(opd, amp, wavelength_m, pupil_dx_m, focal_length_m, pad_factor=pad_factor,
                                                  window=window, pupil_dy_m=pupil_dy_m)
        out_I.append(inten)
    if out_I:
        I = torch.cat(out_I)
    else:
        I = torch.empty(0, py, px, dtype=torch.float64, device=f.device)
    if numpy_io:
        return dict(planes=mine, I=_lib.to_host(I), x=None if x_im is None else _lib.to_host(x_im),
                    y=None if y_im is None else _lib.to_host(y_im))
    return dict(planes=mine, I=I, x=x_im, y=y_im)
