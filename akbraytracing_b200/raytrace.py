"""Path B host side: the reference's ray / quadric-mirror call signatures over sm_100a kernels.

Mirrors (same names, positional order, (3,N) float64 row-major arrays in and out):

* ``mirr_ray_intersection(coeffs, ray, source, negative=False)``  EllipseRaytrace3D.py:18-45
* ``norm_vector(coeffs, point)``                                  ER3D:61-71
* ``reflect_ray(ray, N)``                                         ER3D:47-55
* ``normalize_vector(vector)``                                    ER3D:57-59
* ``plane_ray_intersection(coeffs, ray, source)``                 ER3D:145-157
* ``ell`` (``coeffs``, ``calc_reflect``) / ``PlanePoints``        ER3D:207-263
* ``trace_chain``: the fused K-mirror call sequence of AKB_raytrace_20250312.py:2881-2905 /
  :11039-11054 (no single reference function; one kernel here)

NumPy in -> NumPy out (H2D/D2H inside the call); torch CUDA tensors in -> torch CUDA out.
Reference semantics kept in this layer: if ANY ray has not(D>0) the whole
intersection result is NaN (ER3D:31-33); if ANY column norm is 0 the array is returned
un-normalised (ER3D:57-59).  No CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import _lib

__all__ = ["mirr_ray_intersection", "norm_vector", "reflect_ray", "normalize_vector",
           "plane_ray_intersection", "intersect_reflect", "trace_chain", "trace_chain_batched", "wavefront_opl",
           "rotation_matrices", "ell", "PlanePoints",
           "Ell_define", "calcEll_Yvalue", "shift_x"]


# ---------------------------------------------------------------- plumbing

def _coeffs(c) -> np.ndarray:
    c = np.ascontiguousarray(np.asarray(c, dtype=np.float64).reshape(-1))
    if c.shape[0] != 10:
        raise ValueError("coeffs must hold the 10 quadric coefficients [a..j]")
    return c


class _Ctx:
    """Device staging for one call: uploads NumPy inputs, allocates outputs, reads flags.
    Device arrays = torch CUDA tensors or any object with __cuda_array_interface__ (used in place)."""

    def __init__(self, *arrays):
        import torch
        self.torch = torch
        arrays = [_lib.from_cuda_array(a) if _lib.is_cuda_array(a) else a for a in arrays]
        self.numpy_io = not any(_lib.is_torch(a) for a in arrays)
        dev = None
        for a in arrays:
            if _lib.is_torch(a) and a.is_cuda:
                dev = a.device
                break
        self.device = dev if dev is not None else torch.device("cuda", torch.cuda.current_device())
        self.stream = _lib.torch_stream_ptr(self.device)
        self._flags = None

    @property
    def flags(self):
        """int32[NFLAGS] status block (the entry points zero it on the stream themselves)."""
        if self._flags is None:
            self._flags = self.torch.empty(_lib.NFLAGS, dtype=self.torch.int32, device=self.device)
        return self._flags

    def rays(self, a):
        t = _lib.dev_f64(a, self.device)
        if t.dim() != 2 or t.shape[0] != 3:
            raise ValueError("ray arrays must have shape (3, N)")
        return t

    def pair(self, a, b):
        """Two (3, N) arrays of one call.  Like NumPy in the reference, a (3, 1) array broadcasts against
        (3, N) (one launch point / one direction for every ray); any other mismatch is an error -- the kernels
        index both arrays with the same N."""
        ta, tb = self.rays(a), self.rays(b)
        na, nb = ta.shape[1], tb.shape[1]
        if na != nb:
            if na == 1:
                ta = ta.expand(3, nb).contiguous()
            elif nb == 1:
                tb = tb.expand(3, na).contiguous()
            else:
                raise ValueError(f"ray arrays have {na} and {nb} columns: they must match (or one of them be (3, 1))")
        return ta, tb

    def empty(self, *shape):
        return self.torch.empty(*shape, dtype=self.torch.float64, device=self.device)

    def read_flags(self):
        return [int(v) for v in self.flags.cpu().tolist()]  # synchronises the stream

    def out(self, t):
        return _lib.to_host(t) if self.numpy_io else t

    def nan_like(self, t):
        return self.out(self.torch.full_like(t, float("nan")))


def _run(ctx, name, *args):
    with ctx.torch.cuda.device(ctx.device):
        rc = getattr(_lib.load(), name)(*args)
    _lib.check(rc, name)


# ---------------------------------------------------------------- the five free functions

def mirr_ray_intersection(coeffs, ray, source, negative=False):
    co = _coeffs(coeffs)
    ctx = _Ctx(ray, source)
    r, s = ctx.pair(ray, source)
    p = ctx.empty(3, r.shape[1])
    _run(ctx, "akb_mirr_ray_intersection", _lib.host_ptr(co), _lib.dev_ptr(r), _lib.dev_ptr(s), r.shape[1],
         int(bool(negative)), _lib.dev_ptr(p), _lib.dev_ptr(ctx.flags), ctx.stream)
    if ctx.read_flags()[_lib.FLAG_MISS]:
        return ctx.nan_like(p)  # ER3D:31-33
    return ctx.out(p)


def _normalising(ctx, name, n, args_before, out, args_after=()):
    """Run a normalising op; on a zero norm re-run it un-normalised (ER3D:59)."""
    skip = 0
    for _ in range(2):
        _run(ctx, name, *args_before, n, _lib.dev_ptr(out), skip, _lib.dev_ptr(ctx.flags), ctx.stream, *args_after)
        if skip or not ctx.read_flags()[_lib.FLAG_ZERO_NORM]:
            break
        skip = 1
    return ctx.out(out)


def norm_vector(coeffs, point):
    co = _coeffs(coeffs)
    ctx = _Ctx(point)
    p = ctx.rays(point)
    return _normalising(ctx, "akb_norm_vector", p.shape[1], (_lib.host_ptr(co), _lib.dev_ptr(p)), ctx.empty(3, p.shape[1]))


def reflect_ray(ray, N):
    ctx = _Ctx(ray, N)
    r, n = ctx.pair(ray, N)
    return _normalising(ctx, "akb_reflect_ray", r.shape[1], (_lib.dev_ptr(r), _lib.dev_ptr(n)), ctx.empty(3, r.shape[1]))


def normalize_vector(vector):
    ctx = _Ctx(vector)
    v = ctx.rays(vector)
    return _normalising(ctx, "akb_normalize_vector", v.shape[1], (_lib.dev_ptr(v),), ctx.empty(3, v.shape[1]))


def plane_ray_intersection(coeffs, ray, source):
    co = _coeffs(coeffs)
    ctx = _Ctx(ray, source)
    r, s = ctx.pair(ray, source)
    p = ctx.empty(3, r.shape[1])
    _run(ctx, "akb_plane_ray_intersection", _lib.host_ptr(co), _lib.dev_ptr(r), _lib.dev_ptr(s), r.shape[1],
         _lib.dev_ptr(p), ctx.stream)
    return ctx.out(p)


# ---------------------------------------------------------------- fused forms

def intersect_reflect(coeffs, ray, source, negative=False, want_normal=True, check=True):
    """intersect -> normal -> reflect in one pass (ell.calc_reflect, ER3D:241-245).
    Returns (points, normal-or-None, reflect).  ``check=False`` skips the flag read-back (no
    synchronisation; misses then stay per-ray NaN, the mpmath flavour of III_I:296-301)."""
    co = _coeffs(coeffs)
    ctx = _Ctx(ray, source)
    r, s = ctx.pair(ray, source)
    n = r.shape[1]
    p, rf = ctx.empty(3, n), ctx.empty(3, n)
    nv = ctx.empty(3, n) if want_normal else None
    skip = 0
    for _ in range(3):
        _run(ctx, "akb_intersect_reflect", _lib.host_ptr(co), _lib.dev_ptr(r), _lib.dev_ptr(s), n, int(bool(negative)),
             _lib.dev_ptr(p), _lib.dev_ptr(nv) if nv is not None else None, _lib.dev_ptr(rf), skip,
             _lib.dev_ptr(ctx.flags), ctx.stream)
        if not check:
            return ctx.out(p), (ctx.out(nv) if nv is not None else None), ctx.out(rf)
        flags = ctx.read_flags()
        zero = flags[_lib.FLAG_ZERO_NORM] & ~skip
        if not zero:
            break
        skip |= zero & -zero
    if flags[_lib.FLAG_MISS]:
        return ctx.nan_like(p), (ctx.nan_like(nv) if nv is not None else None), ctx.nan_like(rf)
    return ctx.out(p), (ctx.out(nv) if nv is not None else None), ctx.out(rf)


def trace_chain(coeffs_list, negative_list, plane_coeffs, ray, source, want_normals=False, want_reflects=False,
                want_dist=True, check=True, want_opl=False):
    """K mirrors + detector plane + segment lengths + optical path in ONE kernel (BIG:2881-2905, BIG:11039-11054,
    BIG:3621-3623).

    Returns dict(points (K,3,N), normals/reflects (K,3,N) or None, last_reflect (3,N), det (3,N) or None,
    dist (K,N) or None, opl (N,) or None, flags).  ``opl`` = dist_0 + ... + dist_{K-1} (+ |det - P_K| with a
    plane), the reference's totalDist.  ``check=False`` skips the flag read-back: no host synchronisation, misses
    stay per-ray NaN (the mpmath flavour, III_I:296-301)."""
    K = len(coeffs_list)
    if not 1 <= K <= _lib.MAX_MIRRORS:
        raise ValueError(f"1..{_lib.MAX_MIRRORS} mirrors supported")
    co = np.ascontiguousarray(np.stack([_coeffs(c) for c in coeffs_list]))
    neg = np.ascontiguousarray(np.asarray([1 if b else 0 for b in negative_list], dtype=np.int32))
    if neg.shape[0] != K:
        raise ValueError("negative_list must have one entry per mirror")
    plane = _coeffs(plane_coeffs) if plane_coeffs is not None else None
    ctx = _Ctx(ray, source)
    r, s = ctx.pair(ray, source)
    n = r.shape[1]
    # one allocation for every output (rows of n doubles): points | normals | reflects | last | det | dist | opl | flags
    rows = 3 * K + (3 * K if want_normals else 0) + (3 * K if want_reflects else 0) + 3 + (3 if plane is not None else 0) \
        + (K if want_dist else 0) + (1 if want_opl else 0)
    slab = ctx.empty(rows * n + 2)  # + 16 bytes: the int32[4] status block
    at = [0]

    def take(nrows, *shape):
        v = slab[at[0] * n:(at[0] + nrows) * n].view(*shape)
        at[0] += nrows
        return v
    pts = take(3 * K, K, 3, n)
    nrm = take(3 * K, K, 3, n) if want_normals else None
    rfl = take(3 * K, K, 3, n) if want_reflects else None
    last = take(3, 3, n)
    det = take(3, 3, n) if plane is not None else None
    dist = take(K, K, n) if want_dist else None
    opl = take(1, n) if want_opl else None
    ctx._flags = slab[rows * n:].view(ctx.torch.int32)
    opt = lambda t: _lib.dev_ptr(t) if t is not None else None  # noqa: E731
    skip, flags = 0, [0] * _lib.NFLAGS
    for _ in range(2 * K + 1):
        _run(ctx, "akb_trace_chain", _lib.host_ptr(co), _lib.host_ptr(neg), K,
             _lib.host_ptr(plane) if plane is not None else None, _lib.dev_ptr(r), _lib.dev_ptr(s), n,
             _lib.dev_ptr(pts), opt(nrm), opt(rfl), _lib.dev_ptr(last), opt(det), opt(dist), opt(opl), skip,
             _lib.dev_ptr(ctx.flags), ctx.stream)
        if not check:
            break
        flags = ctx.read_flags()
        zero = flags[_lib.FLAG_ZERO_NORM] & ~skip
        if not zero:
            break
        skip |= zero & -zero
    if check and flags[_lib.FLAG_MISS]:
        # everything from the first mirror with a miss onwards is NaN in the reference
        first = next((k for k in range(K) if flags[_lib.FLAG_MISS_MASK] >> k & 1), 0)
        nan = float("nan")
        pts[first:] = nan
        for t in (nrm, rfl, dist):
            if t is not None:
                t[first:] = nan
        last[:] = nan
        for t in (det, opl):
            if t is not None:
                t[:] = nan
    o = lambda t: ctx.out(t) if t is not None else None  # noqa: E731
    return dict(points=o(pts), normals=o(nrm), reflects=o(rfl), last_reflect=o(last), det=o(det), dist=o(dist),
                opl=o(opl), flags=flags)


def rotation_matrices(theta_y, theta_z):
    """R_y, R_z of rotate_vectors (BIG:917-931), built on the host with NumPy exactly like the reference."""
    R_y = np.array([[np.cos(theta_y), 0, np.sin(theta_y)], [0, 1, 0], [-np.sin(theta_y), 0, np.cos(theta_y)]])
    R_z = np.array([[np.cos(theta_z), -np.sin(theta_z), 0], [np.sin(theta_z), np.cos(theta_z), 0], [0, 0, 1]])
    return R_y, R_z


def wavefront_opl(last_point, last_dir, dist, plane_x, plane2_x=None, theta_y=None, theta_z=None, pivot=None,
                  want_rotated=False):
    """The tail of ``plot_result_debug(p, 'ray_wave')`` (BIG:3516-3558, 3611-3631) in one launch.

    last_point, last_dir: (3,N) hit points / directions after the last mirror; dist: (K,N) segment lengths of
    ``trace_chain`` (or None).  With ``theta_y, theta_z, pivot`` the bundle is first rotated into the detector frame
    like the reference does (``rotate_vectors(v, theta_y, theta_z)`` / ``rotate_points(P, pivot, theta_y, theta_z)``).
    Returns dict(det, opl[, det2, opl2][, point, dir]): detector points on x = plane_x (and x = plane2_x),
    optical paths ``totalDist`` / ``totalDist2``."""
    ctx = _Ctx(last_point, last_dir, dist)
    p, v = ctx.pair(last_point, last_dir)
    n = p.shape[1]
    K = 0
    d = None
    if dist is not None:
        d = _lib.dev_f64(dist, ctx.device)
        if d.dim() != 2 or d.shape[1] != n:
            raise ValueError("dist must have shape (K, N)")
        K = d.shape[0]
    rot = (theta_y is not None) or (theta_z is not None)
    rz = ry = pv = None
    if rot:
        if pivot is None:
            raise ValueError("a rotation needs the pivot point (focus_apprx)")
        R_y, R_z = rotation_matrices(0.0 if theta_y is None else theta_y, 0.0 if theta_z is None else theta_z)
        ry, rz = np.ascontiguousarray(R_y, dtype=np.float64), np.ascontiguousarray(R_z, dtype=np.float64)
        pv = np.ascontiguousarray(np.asarray(pivot, dtype=np.float64).reshape(3))
    det, opl = ctx.empty(3, n), ctx.empty(n)
    det2 = ctx.empty(3, n) if plane2_x is not None else None
    opl2 = ctx.empty(n) if plane2_x is not None else None
    prot = ctx.empty(3, n) if want_rotated else None
    vrot = ctx.empty(3, n) if want_rotated else None
    import ctypes
    px2 = ctypes.c_double(float(plane2_x)) if plane2_x is not None else None
    opt = lambda t: _lib.dev_ptr(t) if t is not None else None  # noqa: E731
    hp = lambda a: _lib.host_ptr(a) if a is not None else None  # noqa: E731
    _run(ctx, "akb_wavefront_opl", _lib.dev_ptr(p), _lib.dev_ptr(v), opt(d), K, n, hp(rz), hp(ry), hp(pv),
         float(plane_x), ctypes.byref(px2) if px2 is not None else None, opt(prot), opt(vrot), _lib.dev_ptr(det),
         opt(det2), _lib.dev_ptr(opl), opt(opl2), ctx.stream)
    o = lambda t: ctx.out(t) if t is not None else None  # noqa: E731
    out = dict(det=o(det), opl=o(opl), det2=o(det2), opl2=o(opl2))
    if want_rotated:
        out.update(point=o(prot), dir=o(vrot))
    return out


def trace_chain_batched(coeffs_batch, negative_list, planes_batch, ray, source, want_det=True):
    """B geometries x one ray bundle in ONE launch (the auto_focus_NA scan pattern, BIG:12746-12895).

    coeffs_batch: (B, K, 10); planes_batch: (B, 10); ray, source: (3, n) shared by all geometries.
    Returns dict(det (B,3,n) or None, mean_y, std_y, mean_z, std_z (B,), miss (B,) int).  The std
    is the population standard deviation the scans use (np.std(detcenter[1, :]), BIG:12786-12787);
    like the reference's NaN fill, a geometry with missing rays gets NaN statistics."""
    co = np.ascontiguousarray(np.asarray(coeffs_batch, dtype=np.float64))
    if co.ndim != 3 or co.shape[2] != 10:
        raise ValueError("coeffs_batch must have shape (B, K, 10)")
    B, K = co.shape[0], co.shape[1]
    planes = np.ascontiguousarray(np.asarray(planes_batch, dtype=np.float64).reshape(B, 10))
    neg = np.ascontiguousarray(np.asarray([1 if b else 0 for b in negative_list], dtype=np.int32))
    if neg.shape[0] != K:
        raise ValueError("negative_list must have one entry per mirror")
    ctx = _Ctx(ray, source)
    r, s = ctx.pair(ray, source)
    n = r.shape[1]
    det = ctx.empty(B, 3, n)
    stats = ctx.empty(B, 4)
    miss = ctx.torch.empty(B, dtype=ctx.torch.int32, device=ctx.device)  # zeroed by the entry point
    _run(ctx, "akb_trace_chain_batched", _lib.host_ptr(co), _lib.host_ptr(neg), K, _lib.host_ptr(planes), B,
         _lib.dev_ptr(r), _lib.dev_ptr(s), n, _lib.dev_ptr(det), _lib.dev_ptr(stats), _lib.dev_ptr(miss), ctx.stream)
    bad = miss > 0
    stats[bad] = float("nan")
    if want_det:
        det[bad] = float("nan")
    o = ctx.out
    return dict(det=o(det) if want_det else None, mean_y=o(stats[:, 0]), std_y=o(stats[:, 1]), mean_z=o(stats[:, 2]),
                std_z=o(stats[:, 3]), miss=o(miss))


# ---------------------------------------------------------------- ER3D's small host-side classes
# O(1) closed-form design algebra: stays on the host, in NumPy, in the reference's operation
# order (the coefficients carry 1e-9 cancellation).

def Ell_define(l1, inc, l2):
    """ER3D:5-14."""
    sita1 = np.arctan(l2 * np.sin(2. * inc) / (l1 + l2 * np.cos(2. * inc)))
    a_ell = (l1 + l2) / 2.
    b_ell = np.sqrt(l1 * l2 * np.sin(inc) ** 2)
    sita3 = np.arcsin(l1 * np.sin(sita1) / l2)
    return a_ell, b_ell, sita1, sita3


def calcEll_Yvalue(a, b, x):
    """ER3D:15-16."""
    return np.sqrt(b ** 2. - (b * (x - np.sqrt(a ** 2. - b ** 2.)) / a) ** 2.)


def shift_x(coeffs, s):
    """ER3D:73-79."""
    a, b, c, d, e, f, g, h, i, j = coeffs
    return [a, b, c, d, e, f, g - 2 * a * s, h - d * s, i - e * s, j + a * s ** 2 - g * s]


class ell:
    """ER3D:207-245: an elliptical mirror with one focus at the origin."""

    def __init__(self, l1, l2, inc, mirr_length):
        self.a_ell, self.b_ell, self.sita1, self.sita3 = Ell_define(l1, inc, l2)
        self.f_ell = np.sqrt(self.a_ell ** 2 - self.b_ell ** 2)
        self.x_center = l1 * np.cos(self.sita1)
        self.y_center = l1 * np.sin(self.sita1)
        self.x1 = self.x_center - mirr_length / 2
        self.x2 = self.x_center + mirr_length / 2
        self.y1 = calcEll_Yvalue(self.a_ell, self.b_ell, self.x1)
        self.y2 = calcEll_Yvalue(self.a_ell, self.b_ell, self.x2)
        self.p1 = np.sqrt(self.x1 ** 2 + self.y1 ** 2)
        self.p2 = np.sqrt(self.x2 ** 2 + self.y2 ** 2)
        self.sita1_1 = np.arctan(self.y1 / self.x1)
        self.sita1_2 = np.arctan(self.y2 / self.x2)
        self.sita3_1 = np.arctan(self.y1 / (2 * self.f_ell - self.x1))
        self.sita3_2 = np.arctan(self.y2 / (2 * self.f_ell - self.x2))
        self.s0_prime_x1 = self.p1 * (np.cos(self.sita1_1) - np.cos(self.sita3_1))
        self.s0_prime_x2 = self.p2 * (np.cos(self.sita1_2) - np.cos(self.sita3_2))
        self.dist_s_f = self.f_ell * 2

    def coeffs(self, option):
        """ER3D:231-240 (like the reference, replaces this method by the coefficient list)."""
        co = np.zeros(10)
        co[0] = 1. / self.a_ell ** 2
        if option == 'y':
            co[1] = 1. / self.b_ell ** 2
        else:
            co[2] = 1. / self.b_ell ** 2
        co[9] = -1.
        self.coeffs = shift_x(co, self.f_ell)

    def calc_reflect(self, inc_vector, inc_points):
        """ER3D:241-245, one fused kernel."""
        self.points, self.N_ell, self.reflect = intersect_reflect(self.coeffs, inc_vector, inc_points)


class PlanePoints:
    """ER3D:246-263: the ray bundle on three detector planes x = position, position -/+ delta."""

    def __init__(self, position, delta, inc_ray, inc_points):
        for name, off in (("points0", 0.0), ("points1", delta), ("points2", -delta)):
            co = np.zeros(10)
            co[6] = 1.
            co[9] = -position + off
            setattr(self, name, plane_ray_intersection(co, inc_ray, inc_points))
