"""Synthetic workloads of BASELINE.json's configs, built on the device from the geometry the
reference's own design code produced (akbraytracing_b200/data/geometry.npz, written by
tests/golden/make_golden.py -- coefficient sets, detector plane and launch-angle tangents of the
kept second pass of the reference drivers at 1000 x 1000 rays).

C1  1e4 source points -> 64x64 focal grid           (synthetic mirror patch)
C2  single elliptical mirror (ER3D ell_v), 3163^2 ~ 1e7 rays, intersect + reflect
C3  KB two-mirror chain (BIG KB_debug geometry), 1000^2 rays -> 512x512 focal grid
C4  AKB four-mirror Wolter III+I chain (BIG plot_result_debug), 1000^2 rays -> 2048x2048
C5  C4's last mirror -> 32 defocus planes x 1024x1024

torch is used only to create/hold device buffers (linspace, tile, meshgrid).
"""
from __future__ import annotations

import os

import numpy as np

from . import handoff, raytrace

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "geometry.npz")
_geo = None

WAVELENGTH_EUV = 13.5e-9   # CPU0402:243 (option_HighNA)
WAVELENGTH_XRAY = 1.35e-9  # CPU0402:245


def geometry():
    global _geo
    if _geo is None:
        z = np.load(_DATA)
        _geo = {k.replace("__", "/"): z[k] for k in z.files}
    return _geo


def _resample(t, n):
    """n launch tangents from the stored 1000: identity at n == len(t), else linear resampling."""
    if n == t.shape[0]:
        return t
    return np.interp(np.linspace(0.0, 1.0, n), np.linspace(0.0, 1.0, t.shape[0]), t)


def launch_grid(tan_h, tan_v, device):
    """Unit launch vectors of an nV x nH ray grid: (1, tan h, tan v) normalised, h fastest
    (BIG:2710-2717).  Returns a (3, nV*nH) CUDA tensor."""
    import torch
    th = torch.as_tensor(np.ascontiguousarray(tan_h), device=device)
    tv = torch.as_tensor(np.ascontiguousarray(tan_v), device=device)
    nH, nV = th.shape[0], tv.shape[0]
    raw = torch.empty(3, nV * nH, dtype=torch.float64, device=device)
    raw[0] = 1.0
    raw[1] = th.repeat(nV)
    raw[2] = tv.repeat_interleave(nH)
    return raytrace.normalize_vector(raw)


def chain_inputs(tag, n, device):
    """(coeffs[K,10], negative[K], plane[10], ray (3,n*n), source (3,n*n)) of config C3 ('c3') or C4 ('c4')."""
    import torch
    g = geometry()
    tan_h, tan_v = _resample(g[f"{tag}/tan_h"], n), _resample(g[f"{tag}/tan_v"], n)
    ray = launch_grid(tan_h, tan_v, device)
    src = torch.as_tensor(g[f"{tag}/source_point"], device=device).reshape(3, 1).expand(3, n * n).contiguous()
    return g[f"{tag}/coeffs"], [bool(b) for b in g[f"{tag}/negative"]], g[f"{tag}/plane"], ray, src


def focal_grid(det, G, half=1e-6, planes=None):
    """G x G detector grid centred on the traced spot (BIG:13570-13591): y,z in centre +/- half,
    x = mean of the traced detector x.  ``planes``: optional x offsets -> stacked grids."""
    import torch
    y, z = det[1], det[2]
    yc = (y.min() + y.max()) / 2
    zc = (z.min() + z.max()) / 2
    yg = torch.linspace(float(yc - half), float(yc + half), G, dtype=torch.float64, device=det.device)
    zg = torch.linspace(float(zc - half), float(zc + half), G, dtype=torch.float64, device=det.device)
    zz, yy = torch.meshgrid(zg, yg, indexing="ij")  # np.meshgrid(y_grid, z_grid) layout: y fastest
    x0 = float(det[0].mean())
    offs = [0.0] if planes is None else list(planes)
    xs = torch.cat([torch.full((G * G,), x0 + float(o), dtype=torch.float64, device=det.device) for o in offs])
    return xs, yy.reshape(-1).repeat(len(offs)), zz.reshape(-1).repeat(len(offs))


def traced_field_inputs(tag, n, G, wavelength=WAVELENGTH_EUV, device="cuda", planes=None):
    """Trace config `tag` and return everything one Fresnel stage last-mirror -> focal grid needs:
    dict(det_x, det_y, det_z, src_x, src_y, src_z, u, ds, k, trace)."""
    coeffs, neg, plane, ray, src = chain_inputs(tag, n, device)
    tr = raytrace.trace_chain(coeffs, neg, plane, ray, src, want_dist=True)
    last = tr["points"][-1]
    k = 2.0 * np.pi / wavelength
    opl = tr["dist"].sum(dim=0)
    u = handoff.opl_to_field(opl, k)
    ds = handoff.calc_dS(last, n, n).reshape(-1)
    dx, dy, dz = focal_grid(tr["det"], G, planes=planes)
    return dict(det_x=dx, det_y=dy, det_z=dz, src_x=last[0].contiguous(), src_y=last[1].contiguous(),
                src_z=last[2].contiguous(), u=u, ds=ds, k=k, trace=tr)


def c1_patch(n_src=10_000, G=64, wavelength=WAVELENGTH_EUV, seed=0):
    """Host (NumPy) inputs of config C1: mirror patch 30 mm x 5 mm x 2 mm at x ~ 146 m, G x G
    detector square of 2 um side 0.15 m downstream, random-phase unit sources."""
    rng = np.random.default_rng(seed)
    sx = 146.0 + rng.uniform(-0.015, 0.015, n_src)
    sy = rng.uniform(-0.0025, 0.0025, n_src)
    sz = rng.uniform(-0.001, 0.001, n_src)
    u = np.exp(2j * np.pi * rng.uniform(0, 1, n_src))
    ds = rng.uniform(0.5e-9, 1.5e-9, n_src)
    yy, zz = np.meshgrid(np.linspace(-1e-6, 1e-6, G), np.linspace(-1e-6, 1e-6, G))
    x = np.full(G * G, 146.15)
    return dict(x=x, y=yy.ravel(), z=zz.ravel(), sx=sx, sy=sy, sz=sz, u=u, ds=ds, k=2.0 * np.pi / wavelength)


def c2_rays(n=3163, device="cuda"):
    """Config C2: (coeffs, ray (3,n*n), source (3,n*n)) for the single elliptical mirror ell_v."""
    import torch
    g = geometry()
    ay = np.linspace(g["c2/angle_y"][0], g["c2/angle_y"][1], n)
    az = np.linspace(g["c2/angle_z"][0], g["c2/angle_z"][1], n)
    ray = launch_grid(np.tan(ay), np.tan(az), device)
    src = torch.zeros(3, n * n, dtype=torch.float64, device=device)
    return g["c2/coeffs"], ray, src
