"""NumPy restatement of the hot-path arithmetic (TEST INFRASTRUCTURE, see oracle/__init__.py).

Written independently of the C restatement so the two can be cross-checked; both
are pinned against outputs of the reference itself (tests/golden/).  Citations:
CPU0402 = Wavecalc_raytrace_fromData_CPU0402.py, ER3D = EllipseRaytrace3D.py,
PSF = psf_fft.py, BIG = AKB_raytrace_20250312.py.
"""
from __future__ import annotations

import numpy as np


# ---------------------------------------------------------------- path A

def fresnel_sum(x, y, z, sx, sy, sz, u, k, ds=None, chunk: int = 256) -> np.ndarray:
    """CPU0402:54-63 (compute_u) evaluated for blocks of detector points.

    u_i = sum_j (u_j ds_j) / r_ij * exp(-i k r_ij); ds pre-multiplication CPU0402:102.
    """
    x = np.asarray(x, np.float64); y = np.asarray(y, np.float64); z = np.asarray(z, np.float64)
    sx = np.asarray(sx, np.float64); sy = np.asarray(sy, np.float64); sz = np.asarray(sz, np.float64)
    w = np.asarray(u, np.complex128) * (1.0 if ds is None else np.asarray(ds, np.float64))
    out = np.empty(x.shape[0], np.complex128)
    for lo in range(0, x.shape[0], chunk):
        hi = min(lo + chunk, x.shape[0])
        ddx = x[lo:hi, None] - sx[None, :]
        ddy = y[lo:hi, None] - sy[None, :]
        ddz = z[lo:hi, None] - sz[None, :]
        r = np.sqrt((ddx * ddx + ddy * ddy) + ddz * ddz)
        ph = -k * r
        kern = (1.0 / r) * (np.cos(ph) + 1j * np.sin(ph))
        out[lo:hi] = (kern * w[None, :]).sum(axis=1)
    return out


def array_split_bounds(total: int, parts: int):
    """Block boundaries of np/cp.array_split (GPU0402:77-79): the first total%parts
    blocks get one extra element."""
    base, extra = divmod(int(total), int(parts))
    bounds, start = [], 0
    for r in range(parts):
        n = base + (1 if r < extra else 0)
        bounds.append((start, n))
        start += n
    return bounds


# ---------------------------------------------------------------- path B

def quadric_value(co, pt):
    """F(x,y,z) of the 10-coefficient quadric (ER3D:19 / ER3D:27-28)."""
    a, b, c, d, e, f, g, h, i, j = [float(v) for v in co]
    x, y, z = pt
    return a * x * x + b * y * y + c * z * z + d * x * y + e * x * z + f * y * z + g * x + h * y + i * z + j


def _unit_columns(v):
    """ER3D:57-59: all-or-nothing column normalisation."""
    nrm = np.sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2])
    if np.all(nrm != 0):
        return v / nrm
    return v


def normalize_vector(v):
    return _unit_columns(np.array(v, dtype=np.float64))


def mirr_ray_intersection(co, ray, src, negative=False):
    """ER3D:18-45: roots of the quadric along src + t*ray, + root unless negative."""
    a, b, c, d, e, f, g, h, i, j = [np.float64(v) for v in co]
    L, M, N = (np.asarray(r, np.float64) for r in ray)
    P, Q, R = (np.asarray(r, np.float64) for r in src)
    qa = a * (L * L) + b * (M * M) + c * (N * N) + d * M * L + e * N * L + f * M * N
    qb = (2 * a * P * L + 2 * b * Q * M + 2 * c * R * N + d * (P * M + Q * L) + e * (P * N + R * L)
          + f * (R * M + Q * N) + g * L + h * M + i * N)
    qc = a * (P * P) + b * (Q * Q) + c * (R * R) + d * P * Q + e * P * R + f * Q * R + g * P + h * Q + i * R + j
    disc = qb * qb - 4 * qa * qc
    out = np.empty((3, L.shape[0]), np.float64)
    if not bool(np.all(disc > 0)):
        out[:] = np.nan
        return out
    root = np.sqrt(disc)
    t = (-qb - root) / (2 * qa) if negative else (-qb + root) / (2 * qa)
    out[0] = t * L + P
    out[1] = t * M + Q
    out[2] = t * N + R
    return out


def norm_vector(co, pt):
    """ER3D:61-71."""
    a, b, c, d, e, f, g, h, i, _ = [np.float64(v) for v in co]
    x, y, z = (np.asarray(r, np.float64) for r in pt)
    grad = np.empty((3, x.shape[0]), np.float64)
    grad[0] = 2 * a * x + d * y + e * z + g
    grad[1] = 2 * b * y + d * x + f * z + h
    grad[2] = 2 * c * z + e * x + f * y + i
    return _unit_columns(grad)


def reflect_ray(ray, nv):
    """ER3D:47-55."""
    ray = np.asarray(ray, np.float64); nv = np.asarray(nv, np.float64)
    dot = ray[0] * nv[0] + ray[1] * nv[1] + ray[2] * nv[2]
    return _unit_columns(ray - 2 * dot * nv)


def plane_ray_intersection(co, ray, src):
    """ER3D:145-157."""
    g, h, i, j = [np.float64(v) for v in co[6:10]]
    L, M, N = (np.asarray(r, np.float64) for r in ray)
    P, Q, R = (np.asarray(r, np.float64) for r in src)
    t = -(g * P + h * Q + i * R + j) / (g * L + h * M + i * N)
    return np.stack([t * L + P, t * M + Q, t * N + R])


# ---------------------------------------------------------------- 'ray_wave' tail (row f-4)

def rotate_vectors(vector, theta_y, theta_z):
    """BIG:917-931: R_y @ (R_z @ v)."""
    R_y = np.array([[np.cos(theta_y), 0, np.sin(theta_y)], [0, 1, 0], [-np.sin(theta_y), 0, np.cos(theta_y)]])
    R_z = np.array([[np.cos(theta_z), -np.sin(theta_z), 0], [np.sin(theta_z), np.cos(theta_z), 0], [0, 0, 1]])
    return R_y @ (R_z @ vector)


def rotate_points(points, focus_apprx, theta_y, theta_z):
    """BIG:933-944."""
    shifted = points - focus_apprx[:, np.newaxis]
    rotated = rotate_vectors(shifted, theta_y, theta_z)
    rotated += focus_apprx[:, np.newaxis]
    return rotated


def wavefront_opl(last_point, last_dir, dist, plane_x, plane2_x, theta_y, theta_z, pivot):
    """The tail of plot_result_debug(p,'ray_wave'): BIG:3589-3598 (rotation into the detector frame, plane),
    BIG:3617-3631 (second plane, dist4tofocus, totalDist, totalDist2)."""
    v = rotate_vectors(np.asarray(last_dir, np.float64), theta_y, theta_z)
    p = rotate_points(np.asarray(last_point, np.float64), np.asarray(pivot, np.float64), theta_y, theta_z)
    out = {"point": p, "dir": v}
    for tag, px in (("", plane_x), ("2", plane2_x)):
        co = np.zeros(10)
        co[6] = 1
        co[9] = -px
        det = plane_ray_intersection(co, v, p)
        total = dist[0]
        for d in dist[1:]:
            total = total + d
        out["det" + tag] = det
        out["opl" + tag] = total + np.linalg.norm(det - p, axis=0)
    return out


# ---------------------------------------------------------------- PSF (convenience row f-3)

def compute_psf_fft(opd_m, amp, wavelength_m, pupil_dx_m, focal_length_m, pad_factor=2, window=None,
                    return_efield=False, pupil_dy_m=None):
    """PSF:29-125 restated: pupil field A*exp(i 2pi/lambda opd) -> (hann) -> even size ->
    centred zero pad -> shifted FFT * dx*dy -> |.|^2 / max."""
    opd_m = np.asarray(opd_m); amp = np.asarray(amp)
    if opd_m.shape != amp.shape:
        raise ValueError("opd_m and amp must have the same shape")
    if pad_factor < 1 or int(pad_factor) != pad_factor:
        raise ValueError("pad_factor must be a positive integer")
    A = np.where(np.isfinite(amp), amp, 0.0).astype(float)
    W = np.where(np.isfinite(opd_m), opd_m, 0.0).astype(float)
    field = A * np.exp(1j * ((2.0 * np.pi / wavelength_m) * W))
    if window is not None:
        if str(window).lower() != "hann":
            raise ValueError(f"Unsupported window '{window}'. Options: 'hann' or None.")
        ny, nx = field.shape
        hy = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(ny) / ny)
        hx = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(nx) / nx)
        h2 = np.outer(hy, hx)
        field = field * (h2 / h2.max())
    ny, nx = field.shape
    field = np.pad(field, ((0, ny % 2), (0, nx % 2)))
    ny, nx = field.shape
    py, px = ny * pad_factor, nx * pad_factor
    oy, ox = (py - ny) // 2, (px - nx) // 2
    big = np.zeros((py, px), complex)
    big[oy:oy + ny, ox:ox + nx] = field
    dy = pupil_dx_m if pupil_dy_m is None else pupil_dy_m
    img = np.fft.fftshift(np.fft.fft2(np.fft.ifftshift(big))) * (pupil_dx_m * dy)
    x_im = wavelength_m * focal_length_m * np.fft.fftshift(np.fft.fftfreq(px, d=pupil_dx_m))
    y_im = wavelength_m * focal_length_m * np.fft.fftshift(np.fft.fftfreq(py, d=dy))
    inten = np.abs(img) ** 2
    peak = inten.max()
    if peak > 0:
        inten = inten / peak
    if return_efield:
        return inten, x_im, y_im, img / np.sqrt(peak if peak > 0 else 1.0)
    return inten, x_im, y_im
