"""CPU oracle for the two hot paths (TEST INFRASTRUCTURE -- never the product path).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  The product
package ``akbraytracing_b200`` never does, and fails loudly when its CUDA library
is missing.

Two restatements of the reference's arithmetic live here:

* ``akb_oracle.c``  -- plain C (OpenMP over detector points), loaded with ctypes;
* ``numpy_port.py`` -- NumPy, used to cross-check the C on small cases.

Parity pinning: the reference has no tests or golden vectors of its own
(SURVEY.md section 4).  Both restatements are pinned against outputs of the reference
itself, generated in the build container by ``tests/golden/make_golden.py`` and
committed under ``tests/golden/*.npz`` (see ``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "akb_oracle.c")
_LIB = os.path.join(_HERE, "liborc.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile akb_oracle.c -> liborc.so (no FMA contraction, OpenMP)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
               "-o", _LIB, _SRC, "-lm"]
        subprocess.run(cmd, check=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        i64, dp, dbl, ci = ctypes.c_int64, ctypes.POINTER(ctypes.c_double), ctypes.c_double, ctypes.c_int
        L.orc_fresnel_sum.argtypes = [i64, dp, dp, dp, i64, dp, dp, dp, dp, dp, dbl, dp, ci]
        L.orc_fresnel_sum.restype = None
        L.orc_mirr_ray_intersection.argtypes = [dp, dp, dp, i64, ci, dp]
        L.orc_mirr_ray_intersection.restype = i64
        L.orc_normalize_vector.argtypes = [dp, i64]
        L.orc_normalize_vector.restype = i64
        L.orc_norm_vector.argtypes = [dp, dp, i64, dp]
        L.orc_norm_vector.restype = i64
        L.orc_reflect_ray.argtypes = [dp, dp, i64, dp]
        L.orc_reflect_ray.restype = i64
        L.orc_plane_ray_intersection.argtypes = [dp, dp, dp, i64, dp]
        L.orc_plane_ray_intersection.restype = None
        L.orc_segment_length.argtypes = [dp, dp, i64, dp]
        L.orc_segment_length.restype = None
        L.orc_calc_dS.argtypes = [dp, i64, i64, dp]
        L.orc_calc_dS.restype = None
        L.orc_max_threads.restype = ci
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def max_threads() -> int:
    return int(lib().orc_max_threads())


# ---------------------------------------------------------------- path A

def fresnel_sum(x, y, z, sx, sy, sz, u, k, ds=None, nthreads: int = 0) -> np.ndarray:
    """u_i = sum_j u_j ds_j exp(-i k r_ij)/r_ij  (CPU0402:71-124)."""
    x, px = _d(x); y, py = _d(y); z, pz = _d(z)
    sx, psx = _d(sx); sy, psy = _d(sy); sz, psz = _d(sz)
    u = np.ascontiguousarray(u, dtype=np.complex128)
    pu = u.view(np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    if ds is None:
        pds = None
    else:
        ds, pds = _d(ds)
    out = np.empty(x.shape[0], dtype=np.complex128)
    po = out.view(np.float64).ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    lib().orc_fresnel_sum(x.shape[0], px, py, pz, sx.shape[0], psx, psy, psz, pu, pds, float(k), po,
                          int(nthreads))
    return out


# ---------------------------------------------------------------- path B

def mirr_ray_intersection(coeffs, ray, source, negative=False) -> np.ndarray:
    co, pco = _d(coeffs); ray, pr = _d(ray); source, ps = _d(source)
    out = np.empty_like(source)
    lib().orc_mirr_ray_intersection(pco, pr, ps, ray.shape[1], int(bool(negative)),
                                    out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out


def normalize_vector(vector) -> np.ndarray:
    v = np.array(vector, dtype=np.float64, order="C", copy=True)
    lib().orc_normalize_vector(v.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), v.shape[1])
    return v


def norm_vector(coeffs, point) -> np.ndarray:
    co, pco = _d(coeffs); point, pp = _d(point)
    out = np.empty_like(point)
    lib().orc_norm_vector(pco, pp, point.shape[1], out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out


def reflect_ray(ray, N) -> np.ndarray:
    ray, pr = _d(ray); N, pn = _d(N)
    out = np.empty_like(ray)
    lib().orc_reflect_ray(pr, pn, ray.shape[1], out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out


def plane_ray_intersection(coeffs, ray, source) -> np.ndarray:
    co, pco = _d(coeffs); ray, pr = _d(ray); source, ps = _d(source)
    out = np.empty_like(source)
    lib().orc_plane_ray_intersection(pco, pr, ps, ray.shape[1],
                                     out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out


def segment_length(a, b) -> np.ndarray:
    a, pa = _d(a); b, pb = _d(b)
    out = np.empty(a.shape[1], dtype=np.float64)
    lib().orc_segment_length(pa, pb, a.shape[1], out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out


def trace_chain(coeffs_list, negative_list, plane_coeffs, ray, source):
    """K-mirror chain + detector plane, the call sequence of BIG:2881-2905 (AKB) and
    BIG:11039-11054 (KB): intersect -> normal -> reflect per mirror, segment lengths,
    then plane_ray_intersection.  Returns dict(points=[K], normals=[K], reflect=[K],
    dist=[K], det=(3,N))."""
    pts, nrm, refl, dist = [], [], [], []
    cur_ray = np.ascontiguousarray(ray, dtype=np.float64)
    cur_src = np.ascontiguousarray(source, dtype=np.float64)
    for co, neg in zip(coeffs_list, negative_list):
        p = mirr_ray_intersection(co, cur_ray, cur_src, neg)
        n = norm_vector(co, p)
        r = reflect_ray(cur_ray, n)
        dist.append(segment_length(cur_src, p))
        pts.append(p); nrm.append(n); refl.append(r)
        cur_ray, cur_src = r, p
    det = plane_ray_intersection(plane_coeffs, cur_ray, cur_src) if plane_coeffs is not None else None
    return dict(points=pts, normals=nrm, reflect=refl, dist=dist, det=det)


def calc_dS(points, ray_num_V, ray_num_H) -> np.ndarray:
    points, pp = _d(points)
    out = np.empty((ray_num_V, ray_num_H), dtype=np.float64)
    lib().orc_calc_dS(pp, int(ray_num_V), int(ray_num_H), out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    return out
