#!/usr/bin/env python
"""Generate the committed golden fixtures by RUNNING THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference has no tests or golden vectors, so the oracle is pinned
against these outputs of the reference's own functions:

* fresnel_ref.npz   forward_propagation_numpy_batch      (CPU0402:87-124, numba path)
                    + WaveField3D.forward_propagation     (CPU0402:38-52)
* ray_er3d_ref.npz  ell.calc_reflect / PlanePoints / free functions (ER3D:18-71,145-157,241-263)
* chain_akb_ref.npz every intersect/normal/reflect/plane call of the kept second pass of
                    plot_result_debug(p,'wave')           (BIG:2881-2905, Wolter III+I)
* chain_kb_ref.npz  same for KB_debug(zeros(26),1,1,'wave') (BIG:11039-11054)
* dS_ref.npz        calc_dS                               (BIG:13418-13473)
* psf_ref.npz       compute_psf_fft                       (PSF:29-125)
* geometry.npz      coefficient sets + launch-angle tangents for the full-size bench
                    configs C2/C3/C4 (values produced by the reference's geometry code;
                    lets the GPU box rebuild the exact reference ray sets without it)

Only arrays are stored; no reference source text is copied.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _refload as R  # noqa: E402


def synthetic_patch(rng, n_src, g, wavelength, x0=146.0, dist=0.15, half=1e-6):
    """Mirror patch 30 mm x 5 mm x 2 mm at x~146 m, g x g detector square `dist` downstream
    (the BASELINE.md section 2 probe geometry)."""
    sx = x0 + rng.uniform(-0.015, 0.015, n_src)
    sy = rng.uniform(-0.0025, 0.0025, n_src)
    sz = rng.uniform(-0.001, 0.001, n_src)
    u = np.exp(2j * np.pi * rng.uniform(0, 1, n_src))
    ds = rng.uniform(0.5e-9, 1.5e-9, n_src)
    yy, zz = np.meshgrid(np.linspace(-half, half, g), np.linspace(-half, half, g))
    x = np.full(g * g, x0 + dist)
    return dict(x=x, y=yy.ravel(), z=zz.ravel(), sx=sx, sy=sy, sz=sz, u=u, ds=ds,
                k=np.float64(2.0 * np.pi / wavelength))


def gen_fresnel(out):
    cpu = R.load_cpu0402()
    rng = np.random.default_rng(0)
    cases = {}
    cases["patch_euv"] = synthetic_patch(rng, 1000, 16, 13.5e-9)
    cases["patch_xray"] = synthetic_patch(rng, 1000, 16, 1.35e-9)
    # ragged / tiny shapes
    cases["ragged"] = synthetic_patch(rng, 37, 3, 13.5e-9)
    c = synthetic_patch(rng, 1, 2, 13.5e-9)
    cases["one_source"] = c
    c = synthetic_patch(rng, 129, 1, 1.35e-9)
    cases["one_detector"] = c
    # source -> M1 stage of the real chain: one point source at the origin, 146 m away
    # (k r ~ 7e10 / 7e11 rad: the reference's own double rounding of k*r matters here)
    for name, lam in (("src_to_m1_euv", 13.5e-9), ("src_to_m1_xray", 1.35e-9)):
        m = 600
        x = 146.0 + rng.uniform(-0.03, 0.03, m)
        y = rng.uniform(-0.004, 0.004, m)
        z = rng.uniform(-0.004, 0.004, m)
        cases[name] = dict(x=x, y=y, z=z, sx=np.zeros(1), sy=np.zeros(1), sz=np.zeros(1),
                           u=np.ones(1, complex), ds=np.ones(1), k=np.float64(2.0 * np.pi / lam))
    # mirror -> mirror stage scale (0.1 m apart), many sources
    m, n = 64, 3000
    c = dict(x=146.1 + rng.uniform(-0.02, 0.02, m), y=rng.uniform(-0.002, 0.002, m),
             z=rng.uniform(-0.002, 0.002, m),
             sx=146.0 + rng.uniform(-0.02, 0.02, n), sy=rng.uniform(-0.002, 0.002, n),
             sz=rng.uniform(-0.002, 0.002, n),
             u=rng.normal(size=n) + 1j * rng.normal(size=n), ds=rng.uniform(1e-9, 2e-9, n),
             k=np.float64(2.0 * np.pi / 1.35e-9))
    cases["mirror_to_mirror"] = c
    # a detector point coinciding with a source point: IEEE inf/nan kept
    c = synthetic_patch(rng, 8, 2, 13.5e-9)
    c["x"][1], c["y"][1], c["z"][1] = c["sx"][3], c["sy"][3], c["sz"][3]
    cases["coincident"] = c

    for name, c in cases.items():
        with R.quiet(), np.errstate(all="ignore"):
            ref = cpu.forward_propagation_numpy_batch(c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"],
                                                      c["u"], c["k"], c["ds"])
            ref_np = np.array([cpu.compute_u(i, c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"],
                                             c["u"] * c["ds"], c["k"]) for i in range(len(c["x"]))])
        for key, val in c.items():
            out[f"{name}/{key}"] = np.asarray(val)
        out[f"{name}/ref"] = ref
        out[f"{name}/ref_numpy"] = ref_np
        print(f"fresnel {name}: M={len(c['x'])} N={len(c['sx'])} |ref|max={np.nanmax(np.abs(ref)):.3e}")

    # the WaveField3D holder (CPU0402:17-52): lambda -> k, setdata, set_ds
    c = cases["patch_euv"]
    back = cpu.WaveField3D(len(c["sx"]), 13.5e-9, 1, 1)
    back.setdata(np.vstack([c["sx"], c["sy"], c["sz"], c["ds"]]))
    back.set_ds(c["ds"])
    back.u = c["u"].copy()
    front = cpu.WaveField3D(len(c["x"]), 13.5e-9, 16, 16)
    front.setdata(np.vstack([c["x"], c["y"], c["z"]]))
    with R.quiet():
        front.forward_propagation(back)
    out["wavefield/u"] = front.u
    out["wavefield/lambda"] = np.float64(13.5e-9)


def kb_geometry(er):
    """ER3D:309-328 (the __main__ design numbers), evaluated with the reference's own code."""
    l1h, l2h, inc_h, mlen_h, wd_v, inc_v, mlen_v = [np.float64(v) for v in
                                                    (146., 0.086, 0.214, 0.060, 0.0211, 0.21, 0.0232)]
    inc_h /= 20
    inc_v /= 20
    with R.quiet():
        ell_v = er.ell(l1h, l2h, inc_h, mlen_h)
        vir = (ell_v.s0_prime_x1 + ell_v.s0_prime_x2) / 2.
        res = er.KB_define(l1h, l2h, inc_h, mlen_h, wd_v, inc_v, mlen_v, gapf=vir)
        l1v, l2v = res[4], res[5]
        ell_h = er.ell(l1v, l2v, inc_v, mlen_v)
        ell_v.coeffs('y')
        ell_h.coeffs('z')
    return ell_v, ell_h


def er3d_rays(er, ell_v, ell_h, n):
    ay = np.linspace(ell_v.sita1_1, ell_v.sita1_2, n)
    az = np.linspace(ell_h.sita1_1, ell_h.sita1_2, n)
    yy, zz = np.meshgrid(ay, az)
    vec = np.zeros((3, n, n))
    vec[0] = 1
    vec[1] = np.tan(yy)
    vec[2] = np.tan(zz)
    vec = er.normalize_vector(vec.reshape(3, -1))
    return vec, np.zeros((3, n * n))


def gen_ray_er3d(out, geo):
    er = R.load_er3d()
    ell_v, ell_h = kb_geometry(er)
    n = 32
    vec, src = er3d_rays(er, ell_v, ell_h, n)
    co_v = np.array(ell_v.coeffs, dtype=np.float64)
    co_h = np.array(ell_h.coeffs, dtype=np.float64)
    out["single/coeffs"] = co_v
    out["single/ray"] = vec
    out["single/source"] = src
    ell_v.calc_reflect(vec, src)  # ER3D:241-245
    out["single/points"] = ell_v.points
    out["single/N_ell"] = ell_v.N_ell
    out["single/reflect"] = ell_v.reflect
    with R.quiet():
        pp = er.PlanePoints(ell_v.dist_s_f, 1e-8, ell_v.reflect, ell_v.points)  # ER3D:246-263
    out["single/plane_position"] = np.float64(ell_v.dist_s_f)
    out["single/plane_delta"] = np.float64(1e-8)
    out["single/points0"], out["single/points1"], out["single/points2"] = pp.points0, pp.points1, pp.points2
    # second mirror used stand-alone on the same rays (ER3D:363) -- option 'z' cylinder
    ell_h.calc_reflect(vec, src)
    out["single_z/coeffs"] = co_h
    out["single_z/points"], out["single_z/N_ell"], out["single_z/reflect"] = ell_h.points, ell_h.N_ell, ell_h.reflect
    # negative root of the same quadric
    out["negative/points"] = er.mirr_ray_intersection(co_v, vec, src, negative=True)
    # off-origin sources + a general (rotated) quadric: exercises every coefficient
    rng = np.random.default_rng(1)
    co_rot, _ = er.rotate_general_axis(list(co_v), [0.3, 1.0, 0.2], 1e-3, [ell_v.x_center, ell_v.y_center, 0.0])
    co_rot = np.array(co_rot, dtype=np.float64)
    src2 = rng.normal(scale=1e-3, size=src.shape)
    out["general/coeffs"] = co_rot
    out["general/source"] = src2
    p = er.mirr_ray_intersection(co_rot, vec, src2)
    nv = er.norm_vector(co_rot, p)
    rf = er.reflect_ray(vec, nv)
    out["general/points"], out["general/normal"], out["general/reflect"] = p, nv, rf
    # a miss: one ray pointing away -> the WHOLE result is NaN (ER3D:31-33)
    vec_miss = vec.copy()
    vec_miss[:, 5] = [1.0, 0.0, 0.0]  # the line y = 10 never meets the cylinder (|y| <= b)
    src_miss = src.copy()
    src_miss[:, 5] = [0.0, 10.0, 0.0]
    out["miss/ray"] = vec_miss
    out["miss/source"] = src_miss
    out["miss/points"] = er.mirr_ray_intersection(co_v, vec_miss, src_miss)
    with np.errstate(all="ignore"):
        out["miss/normal"] = er.norm_vector(co_v, out["miss/points"])
        out["miss/reflect"] = er.reflect_ray(vec_miss, out["miss/normal"])
    # all-or-nothing normalisation (ER3D:57-59)
    v = rng.normal(size=(3, 7))
    out["normalize/in"] = v
    out["normalize/out"] = er.normalize_vector(v.copy())
    v0 = v.copy()
    v0[:, 2] = 0.0
    out["normalize/in_zero"] = v0
    out["normalize/out_zero"] = er.normalize_vector(v0.copy())
    # full-size C2 description (single mirror ell_v, n=3163 -> 1.0005e7 rays)
    geo["c2/coeffs"] = co_v
    geo["c2/angle_y"] = np.array([ell_v.sita1_1, ell_v.sita1_2])
    geo["c2/angle_z"] = np.array([ell_h.sita1_1, ell_h.sita1_2])
    geo["c2/plane_position"] = np.float64(ell_v.dist_s_f)
    print("ray er3d: max|F(P)| =", np.abs(
        co_v[0] * ell_v.points[0] ** 2 + co_v[1] * ell_v.points[1] ** 2 + co_v[6] * ell_v.points[0] + co_v[9]).max())


def _record(ns, names):
    calls = []

    def wrap(name):
        orig = ns[name]

        def rec(*a, **k):
            res = orig(*a, **k)
            calls.append(dict(fn=name, args=[np.array(x, dtype=np.float64) for x in a],
                              negative=bool(k.get("negative", a[3] if (name == "mirr_ray_intersection" and len(a) > 3) else False)),
                              out=np.array(res, dtype=np.float64)))
            return res
        ns[name] = rec
    for nm in names:
        wrap(nm)
    return calls


HOT = ["mirr_ray_intersection", "norm_vector", "reflect_ray", "plane_ray_intersection", "normalize_vector"]


def _kept_pass(calls, n_rays, n_mirrors):
    """Pick the calls of the kept second pass: the LAST n_mirrors full-size intersect calls,
    their normal/reflect calls, the launch-vector normalisation before them and the plane
    call after them."""
    idx = [i for i, c in enumerate(calls) if c["fn"] == "mirr_ray_intersection" and c["args"][1].shape[1] == n_rays]
    idx = idx[-n_mirrors:]
    first = idx[0]
    launch = calls[first - 1]
    assert launch["fn"] == "normalize_vector" and launch["args"][0].shape[1] == n_rays
    mirrors = []
    for i in idx:
        inter = calls[i]
        # intersect, normalize(inside norm_vector), norm_vector, normalize(inside reflect), reflect
        nv = next(c for c in calls[i + 1:] if c["fn"] == "norm_vector")
        rf = next(c for c in calls[i + 1:] if c["fn"] == "reflect_ray")
        mirrors.append((inter, nv, rf))
    plane = next(c for c in calls[idx[-1] + 1:] if c["fn"] == "plane_ray_intersection" and c["args"][1].shape[1] == n_rays)
    return launch, mirrors, plane


def gen_chain(out, geo, kind):
    small, full = 33, 1000
    for n, store_arrays in ((small, True), (full, False)):
        ns = R.load_big(wave_num=n, option_AKB=(kind == "akb"))
        calls = _record(ns, HOT)
        with R.quiet(), np.errstate(all="ignore"):
            if kind == "akb":
                ret = ns["plot_result_debug"](np.array(R.AKB_ALIGNMENT_PRESET), "wave")
                n_mirrors = 4
            else:
                ret = ns["KB_debug"](np.zeros(26), 1, 1, "wave")
                n_mirrors = 2
        launch, mirrors, plane = _kept_pass(calls, n * n, n_mirrors)
        raw = launch["args"][0]  # un-normalised launch vectors: (1, tan h, tan v)
        tan_h = raw[1, :n].copy()
        tan_v = raw[2, ::n].copy()
        assert np.array_equal(raw[1], np.tile(tan_h, n)) and np.array_equal(raw[2], np.repeat(tan_v, n))
        coeffs = np.stack([m[0]["args"][0] for m in mirrors])
        negative = np.array([m[0]["negative"] for m in mirrors])
        src0 = mirrors[0][0]["args"][2]
        assert np.all(src0 == src0[:, :1])
        if store_arrays:
            out["tan_h"], out["tan_v"] = tan_h, tan_v
            out["coeffs"], out["negative"] = coeffs, negative
            out["plane"] = plane["args"][0]
            out["source_point"] = src0[:, 0].copy()
            out["ray0"] = launch["out"]
            for k, (inter, nv, rf) in enumerate(mirrors):
                out[f"P{k}"] = inter["out"]
                out[f"N{k}"] = nv["out"]
                out[f"R{k}"] = rf["out"]
                # inputs really chained? (reflect k-1 -> ray k, point k-1 -> source k)
                if k:
                    assert np.array_equal(inter["args"][1], mirrors[k - 1][2]["out"])
                    assert np.array_equal(inter["args"][2], mirrors[k - 1][0]["out"])
            out["det"] = plane["out"]
            prev = src0
            for k, (inter, _, _) in enumerate(mirrors):
                out[f"dist{k}"] = np.linalg.norm(inter["out"] - prev, axis=0)  # BIG:2884-2897
                prev = inter["out"]
            print(f"chain {kind}: n={n} mirrors={n_mirrors} negative={negative.tolist()} "
                  f"det x={plane['out'][0].mean():.6f} nan={np.isnan(plane['out']).any()}")
        else:
            tag = "c4" if kind == "akb" else "c3"
            geo[f"{tag}/coeffs"], geo[f"{tag}/negative"] = coeffs, negative
            geo[f"{tag}/plane"] = plane["args"][0]
            geo[f"{tag}/source_point"] = src0[:, 0].copy()
            geo[f"{tag}/tan_h"], geo[f"{tag}/tan_v"] = tan_h, tan_v
            # a few full-size spot values so the GPU-box rebuild can be checked end to end
            sel = np.array([0, 1, n - 1, n * n // 2 + 17, n * n - 1])
            geo[f"{tag}/spot_index"] = sel
            geo[f"{tag}/spot_det"] = plane["out"][:, sel]
            geo[f"{tag}/spot_last_point"] = mirrors[-1][0]["out"][:, sel]
            geo[f"{tag}/det_mean"] = plane["out"].mean(axis=1)
            print(f"geometry {tag}: n={n} nan={np.isnan(plane['out']).any()} det mean={plane['out'].mean(axis=1)}")
        if store_arrays and kind == "kb":
            # hand-off: calc_dS on the (unrotated) last-mirror cloud, BIG:13520-13551
            pts = mirrors[-1][0]["out"]
            with R.quiet():
                dS = ns["calc_dS"](pts, n, n)
            out["dS_points"] = pts
            out["dS"] = dS


def gen_dS(out):
    ns = R.load_big(wave_num=9)
    rng = np.random.default_rng(2)
    nV, nH = 7, 9
    yy, zz = np.meshgrid(np.linspace(-1e-2, 1e-2, nH), np.linspace(-2e-3, 2e-3, nV))
    pts = np.vstack([(146.0 + 0.3 * yy ** 2 + 0.1 * zz + 1e-5 * rng.normal(size=yy.shape)).ravel(),
                     (yy + 1e-5 * rng.normal(size=yy.shape)).ravel(),
                     (zz + 1e-5 * rng.normal(size=yy.shape)).ravel()])
    with R.quiet():
        out["points"] = pts
        out["nV"], out["nH"] = np.int64(nV), np.int64(nH)
        out["dS"] = ns["calc_dS"](pts, nV, nH)
    # degenerate 3x3 (only one interior point)
    pts3 = pts.reshape(3, nV, nH)[:, :3, :3].reshape(3, -1).copy()
    with R.quiet():
        out["points3"] = pts3
        out["dS3"] = ns["calc_dS"](pts3, 3, 3)


def gen_psf(out):
    psf = R.load_psf()
    rng = np.random.default_rng(3)
    ny, nx = 31, 32  # odd side exercises ensure_even_size (PSF:6-18)
    yy, xx = np.mgrid[0:ny, 0:nx]
    rr = np.hypot((yy - ny / 2) / (ny / 2), (xx - nx / 2) / (nx / 2))
    amp = (rr < 0.9).astype(float)
    opd = 3e-9 * (rr ** 2) + 1e-10 * rng.normal(size=rr.shape)
    opd[3, 4] = np.nan
    amp[5, 6] = np.nan
    out["opd"], out["amp"] = opd, amp
    for tag, kw in (("plain", dict(pad_factor=2)), ("hann", dict(pad_factor=3, window="hann", pupil_dy_m=1.5e-4))):
        I, x_im, y_im, E = psf.compute_psf_fft(opd, amp, 13.5e-9, 1e-4, 0.3, return_efield=True, **kw)
        out[f"{tag}/I"], out[f"{tag}/x"], out[f"{tag}/y"], out[f"{tag}/E"] = I, x_im, y_im, E


def gen_ray_mpmath(out, akb_chain):
    """The reference's mpmath twins (AKB_raytrace_III_I_20250710.py:275-313, 325-377, 413-467), executed
    at 30 digits on 48 rays of the AKB chain: a >FP64 truth for the 1e-12 gate, and the per-ray NaN
    flavour of a miss (III_I:296-301)."""
    import ast
    import mpmath
    from mpmath import mp
    mp.dps = 30
    path = os.path.join(R.REF, "AKB_raytrace_III_I_20250710.py")
    tree = ast.parse(open(path, encoding="utf-8").read(), filename=path)
    want = {"mirr_ray_intersection", "norm_vector", "normalize_vector", "reflect_ray", "plane_ray_intersection",
            "mpmath_norm"}
    defs = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    ns = {"np": np}
    for name in ("mp", "mpf", "sin", "cos", "tan", "sqrt", "pi", "fabs", "asin", "acos", "atan", "isnan", "nan", "matrix", "nstr"):
        ns[name] = getattr(mpmath, name)
    exec(compile(ast.Module(body=defs, type_ignores=[]), path, "exec"), ns)

    def to_mp(a):
        m = mpmath.matrix(a.shape[0], a.shape[1])
        for i in range(a.shape[0]):
            for j in range(a.shape[1]):
                m[i, j] = mpmath.mpf(float(a[i, j]))
        return m

    def to_np(m):
        return np.array([[float(m[i, j]) for j in range(m.cols)] for i in range(m.rows)])

    n = akb_chain["tan_h"].shape[0]
    sel = np.linspace(0, n * n - 1, 48).astype(int)
    ray = akb_chain["ray0"][:, sel]
    src = np.repeat(akb_chain["source_point"][:, None], len(sel), axis=1)
    out["ray"], out["source"] = ray, src
    out["coeffs"], out["negative"], out["plane"] = akb_chain["coeffs"], akb_chain["negative"], akb_chain["plane"]
    cur_ray, cur_src = to_mp(ray), to_mp(src)
    for k in range(4):
        co = [mpmath.mpf(float(v)) for v in akb_chain["coeffs"][k]]
        p = ns["mirr_ray_intersection"](co, cur_ray, cur_src, negative=bool(akb_chain["negative"][k]))
        nv = ns["norm_vector"](co, p)
        rf = ns["reflect_ray"](cur_ray, nv)
        out[f"P{k}"], out[f"N{k}"], out[f"R{k}"] = to_np(p), to_np(nv), to_np(rf)
        cur_ray, cur_src = rf, p
    det = ns["plane_ray_intersection"]([mpmath.mpf(float(v)) for v in akb_chain["plane"]], cur_ray, cur_src)
    out["det"] = to_np(det)
    # per-ray miss: ray 5 starts at the centre of the first (hyperbolic) quadric and runs along z,
    # between the two branches: D < 0 for this ray only
    ray_m, src_m = ray.copy(), src.copy()
    c0 = akb_chain["coeffs"][0]
    ray_m[:, 5] = [0.0, 0.0, 1.0]
    src_m[:, 5] = [-c0[6] / (2 * c0[0]), 0.0, 0.0]
    co0 = [mpmath.mpf(float(v)) for v in akb_chain["coeffs"][0]]
    pm = ns["mirr_ray_intersection"](co0, to_mp(ray_m), to_mp(src_m))
    out["miss_ray"], out["miss_source"], out["miss_P0"] = ray_m, src_m, to_np(pm)
    print("ray mpmath: miss column NaN:", np.isnan(out["miss_P0"][:, 5]).all(), "others finite:",
          np.isfinite(np.delete(out["miss_P0"], 5, axis=1)).all())


def gen_stagechain(out):
    """SURVEY 8(f-1), pinned by RUNNING the reference end to end for a 25 x 25-ray KB and AKB case:
      1. saveWaveData (BIG:13475-13654, executed from the AST-extracted defs; it ends in sys.exit()) writes the
         hand-off folder output_<timestamp>/: every file is recorded, together with what the tracer handed to it
         (mirror clouds, detector points of both planes) so that write_handoff can be tested on the same inputs;
      2. the folder is renamed to the name CPU0402:192 hard-codes and the UNMODIFIED
         Wavecalc_raytrace_fromData_CPU0402.py is run as __main__ (runpy, h5py stubbed) in that directory:
         its complex_data_*.npz and the grids it re-saves are recorded.
    wave_num is odd: saveWaveData only defines size_v1.. on its odd branch (BIG:13489-13504)."""
    import datetime as _dt
    import glob
    import runpy
    import shutil
    import tempfile
    n = 25
    for kind in ("kb", "akb"):
        ns = R.load_big(wave_num=n, option_AKB=(kind == "akb"))
        ns["datetime"] = _dt.datetime
        tracer = "plot_result_debug" if kind == "akb" else "KB_debug"
        seen = {}
        orig = ns[tracer]

        def spy(*a, _orig=orig, **k):
            seen["ret"] = _orig(*a, **k)
            return seen["ret"]
        ns[tracer] = spy
        params = np.array(R.AKB_ALIGNMENT_PRESET) if kind == "akb" else np.zeros(26)
        work = tempfile.mkdtemp(prefix=f"akb_handoff_{kind}_")
        cwd = os.getcwd()
        os.chdir(work)
        try:
            with R.quiet(), np.errstate(all="ignore"):
                try:
                    ns["saveWaveData"](params)
                except SystemExit:
                    pass
            (folder,) = glob.glob("output_*")
            ret = seen["ret"]
            if kind == "akb":  # BIG:13479: source, vmirr_hyp, hmirr_hyp, vmirr_ell, hmirr_ell, detcenter, detcenter2, nH, nV, ...
                clouds, det, det2, nH, nV = list(ret[1:5]), ret[5], ret[6], ret[7], ret[8]
            else:              # BIG:13485: source, vmirr_hyp, hmirr_hyp, detcenter, detcenter2, nH, nV, ...
                clouds, det, det2, nH, nV = list(ret[1:3]), ret[3], ret[4], ret[5], ret[6]
            out[f"{kind}/tracer_source"] = np.asarray(ret[0], dtype=np.float64)
            for i, c in enumerate(clouds):
                out[f"{kind}/tracer_M{i + 1}"] = np.asarray(c, dtype=np.float64)
            out[f"{kind}/tracer_det"], out[f"{kind}/tracer_det2"] = np.asarray(det), np.asarray(det2)
            out[f"{kind}/ray_num"] = np.array([nV, nH])
            out[f"{kind}/params"] = params
            out[f"{kind}/handoff_listing"] = np.array(sorted(os.listdir(folder)))
            for f in sorted(os.listdir(folder)):
                if f.endswith(".npy"):
                    out[f"{kind}/handoff/{f[:-4]}"] = np.load(os.path.join(folder, f))
            out[f"{kind}/handoff/conditions_txt"] = np.array(open(os.path.join(folder, "calculation_conditions.txt")).read())
            # ---- the Wavecalc script itself on that folder
            os.rename(folder, "output_20250404_sNAAKB701")  # CPU0402:192
            R._stub("h5py")
            with R.quiet() as log, np.errstate(all="ignore"):
                runpy.run_path(os.path.join(R.REF, "Wavecalc_raytrace_fromData_CPU0402.py"), run_name="__main__")
            (res,) = [d for d in glob.glob("output_*") if d != "output_20250404_sNAAKB701"]
            out[f"{kind}/wavecalc_listing"] = np.array(sorted(os.listdir(res)))
            for f in sorted(os.listdir(res)):
                path = os.path.join(res, f)
                if f.endswith(".npz"):
                    with np.load(path) as z:
                        out[f"{kind}/wavecalc/{f[:-4]}"] = z["data"]
                elif f.startswith("points_grid"):
                    out[f"{kind}/wavecalc/{f[:-4]}"] = np.load(path)
            img = out[f"{kind}/wavecalc/complex_data_Image"]
            print(f"stagechain {kind}: hand-off {list(out[f'{kind}/handoff_listing'])}; wavecalc wrote "
                  f"{list(out[f'{kind}/wavecalc_listing'])}; |Image| peak at {int(np.argmax(np.abs(img)))} of {img.size}")
        finally:
            os.chdir(cwd)
            shutil.rmtree(work, ignore_errors=True)


class _Stop(Exception):
    pass


def gen_ray_wave(out):
    """plot_result_debug(p, 'ray_wave') (BIG:3565-3631) run from the AST-extracted defs with a 33 x 33
    bundle.  The first griddata call of that branch (BIG:3665) is replaced by a spy that copies the caller's local
    variables -- the rotation angles into the detector frame, the rotated bundle, both detector planes, totalDist,
    totalDist2, DistError2 -- and stops the function there (what follows is interpolation, plotting, PSF)."""
    n = 33
    ns = R.load_big(wave_num=n, option_AKB=True)
    calls = _record(ns, HOT)
    grabbed = {}

    def spy(*a, **k):
        grabbed.update(sys._getframe(1).f_locals)
        raise _Stop()
    ns["griddata"] = spy
    with R.quiet(), np.errstate(all="ignore"):
        try:
            ns["plot_result_debug"](np.array(R.AKB_ALIGNMENT_PRESET), "ray_wave", option_save=False)
        except _Stop:
            pass
    L = grabbed
    N = n * n
    # the unrotated bundle after the 4th mirror: the LAST full-size intersect / reflect calls of the kept pass
    last_int = [c for c in calls if c["fn"] == "mirr_ray_intersection" and c["args"][1].shape[1] == N][-1]
    last_ref = [c for c in calls if c["fn"] == "reflect_ray" and c["args"][0].shape[1] == N][-1]
    assert np.array_equal(last_int["out"], L["hmirr_hyp0"])
    k = "akb"
    out[f"{k}/last_point"], out[f"{k}/last_dir"] = last_int["out"], last_ref["out"]
    out[f"{k}/dist"] = np.stack([L["dist0to1"], L["dist1to2"], L["dist2to3"], L["dist3to4"]])
    out[f"{k}/theta_y"], out[f"{k}/theta_z"] = np.float64(-L["theta_y"]), np.float64(-L["theta_z"])  # rotate_*(.., -theta_y, -theta_z), BIG:3589-3591
    out[f"{k}/pivot"] = np.asarray(L["focus_apprx"], dtype=np.float64)
    out[f"{k}/plane_x"] = np.float64(L["s2f_middle"] + L["defocus"])
    out[f"{k}/plane2_x"] = np.float64(L["s2f_middle"] + L["defocus"] + L["defocusWave"])
    out[f"{k}/point_rot"], out[f"{k}/dir_rot"] = L["hmirr_hyp"], L["reflect4"]
    out[f"{k}/det"], out[f"{k}/det2"] = L["detcenter"], L["detcenter2"]
    out[f"{k}/opl"], out[f"{k}/opl2"] = L["totalDist"], L["totalDist2"]
    out[f"{k}/DistError2"] = L["DistError2"]
    print(f"ray_wave akb: n={n} theta_y={L['theta_y']:.3e} theta_z={L['theta_z']:.3e} totalDist mean {np.nanmean(L['totalDist']):.9f} "
          f"std {np.nanstd(L['totalDist']):.3e}  defocusWave {L['defocusWave']}")


SEPARATE = {"stagechain_ref": gen_stagechain, "ray_wave_ref": gen_ray_wave}  # fixtures added after round 1: `make_golden.py stagechain_ref`


def main():
    assert R.available(), "needs /root/reference (build container only)"
    if len(sys.argv) > 1:  # regenerate only the named fixtures
        for name in sys.argv[1:]:
            d = {}
            SEPARATE[name](d)
            path = os.path.join(HERE, name + ".npz")
            np.savez_compressed(path, **{k.replace("/", "__"): v for k, v in d.items()})
            print("wrote", path, os.path.getsize(path), "bytes")
        return
    geo = {}
    files = {}
    for name, fn in (("fresnel_ref", gen_fresnel), ("dS_ref", gen_dS), ("psf_ref", gen_psf)):
        files[name] = {}
        fn(files[name])
    files["ray_er3d_ref"] = {}
    gen_ray_er3d(files["ray_er3d_ref"], geo)
    for kind in ("akb", "kb"):
        files[f"chain_{kind}_ref"] = {}
        gen_chain(files[f"chain_{kind}_ref"], geo, kind)
    files["ray_mp_ref"] = {}
    gen_ray_mpmath(files["ray_mp_ref"], files["chain_akb_ref"])
    files["geometry"] = geo
    pkg_data = os.path.join(os.path.dirname(os.path.dirname(HERE)), "akbraytracing_b200", "data")
    os.makedirs(pkg_data, exist_ok=True)
    for name, d in files.items():
        # geometry.npz ships with the package (bench / workload builders read it on the GPU box)
        path = os.path.join(pkg_data if name == "geometry" else HERE, name + ".npz")
        np.savez_compressed(path, **{k.replace("/", "__"): v for k, v in d.items()})
        print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
