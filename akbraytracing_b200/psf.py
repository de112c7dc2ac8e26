"""Fraunhofer PSF of a pupil OPD/amplitude map on the device.

``compute_psf_fft`` keeps the signature and return values of psf_fft.py:29-125; the FFT is a plain
``torch.fft.fft2`` library call (north_star item 4: run on-device, not an optimisation target).
``field_to_pupil`` turns an assembled complex field into the (opd, amp) pair the function expects
(amp = |u|, opd = arg(u) * lambda / 2pi).
"""
from __future__ import annotations

import numpy as np

from . import _lib

__all__ = ["compute_psf_fft", "psf_to_db", "ensure_even_size", "field_to_pupil"]


def ensure_even_size(arr):
    """psf_fft.py:6-18: pad odd sides by one zero pixel; returns (array, crop-slices-or-None)."""
    import torch
    ny, nx = arr.shape
    py, px = ny % 2, nx % 2
    if py or px:
        return torch.nn.functional.pad(arr, (0, px, 0, py)), (slice(0, ny), slice(0, nx))
    return arr, None


def _hann2d(ny, nx, device):
    """psf_fft.py:20-27: separable Hann window with unit peak."""
    import torch
    wx = 0.5 - 0.5 * torch.cos(2 * np.pi * torch.arange(nx, dtype=torch.float64, device=device) / nx)
    wy = 0.5 - 0.5 * torch.cos(2 * np.pi * torch.arange(ny, dtype=torch.float64, device=device) / ny)
    w2 = torch.outer(wy, wx)
    return w2 / w2.max()


def compute_psf_fft(opd_m, amp, wavelength_m, pupil_dx_m, focal_length_m, pad_factor=2, window=None,
                    return_efield=False, pupil_dy_m=None, device="cuda"):
    """psf_fft.py:29-125 on `device`.  NumPy in -> NumPy out, torch in -> torch out."""
    import torch
    numpy_io = not (_lib.is_torch(opd_m) or _lib.is_torch(amp))
    if tuple(opd_m.shape) != tuple(amp.shape):
        raise ValueError("opd_m and amp must have the same shape")
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
    U_p = torch.polar(A, (2.0 * np.pi / wavelength_m) * opd)            # A * exp(i phase), PSF:89-90
    if window is not None:
        if str(window).lower() != "hann":
            raise ValueError(f"Unsupported window '{window}'. Options: 'hann' or None.")
        U_p = U_p * _hann2d(U_p.shape[0], U_p.shape[1], dev)
    U_p, _ = ensure_even_size(U_p)
    ny, nx = U_p.shape
    py, px = ny * pad_factor, nx * pad_factor
    oy, ox = (py - ny) // 2, (px - nx) // 2
    U_pad = torch.zeros(py, px, dtype=torch.complex128, device=dev)
    U_pad[oy:oy + ny, ox:ox + nx] = U_p
    dx = pupil_dx_m
    dy = dx if pupil_dy_m is None else pupil_dy_m
    U_im = torch.fft.fftshift(torch.fft.fft2(torch.fft.ifftshift(U_pad))) * (dx * dy)   # PSF:110
    x_im = wavelength_m * focal_length_m * np.fft.fftshift(np.fft.fftfreq(px, d=dx))    # PSF:112-115
    y_im = wavelength_m * focal_length_m * np.fft.fftshift(np.fft.fftfreq(py, d=dy))
    inten = U_im.real ** 2 + U_im.imag ** 2
    peak = float(inten.max())
    if peak > 0:
        inten = inten / peak
    ef = U_im / np.sqrt(peak if peak > 0 else 1.0) if return_efield else None
    if numpy_io:
        inten = _lib.to_host(inten)
        ef = _lib.to_host(ef) if ef is not None else None
    else:
        x_im, y_im = torch.as_tensor(x_im, device=dev), torch.as_tensor(y_im, device=dev)
    if return_efield:
        return inten, x_im, y_im, ef
    return inten, x_im, y_im


def psf_to_db(psf, floor_db=-60.0):
    """psf_fft.py:127-131."""
    if _lib.is_torch(psf):
        import torch
        return 10.0 * torch.log10(torch.clamp(psf, min=10.0 ** (floor_db / 10.0)))
    with np.errstate(divide="ignore"):
        return 10.0 * np.log10(np.maximum(psf, 10.0 ** (floor_db / 10.0)))


def field_to_pupil(u, wavelength_m, shape=None):
    """(opd, amp) of an assembled complex field: amp = |u|, opd = arg(u) * lambda / (2 pi)."""
    if _lib.is_torch(u):
        import torch
        amp, opd = u.abs(), torch.angle(u) * (wavelength_m / (2.0 * np.pi))
    else:
        amp, opd = np.abs(u), np.angle(u) * (wavelength_m / (2.0 * np.pi))
    if shape is not None:
        amp, opd = amp.reshape(shape), opd.reshape(shape)
    return opd, amp
