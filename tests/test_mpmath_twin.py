"""The reference's SECOND implementation of the ray kernels: the mpmath twins of
AKB_raytrace_III_I_20250710.py (275-313, 325-377, 413-467), executed at 30 digits by
tests/golden/make_golden.py on 48 rays of the AKB four-mirror chain with the SAME 10-double
coefficient sets.

What they show: the FP64 (NumPy-order) arithmetic of the reference -- which the CUDA kernels
reproduce bit for bit -- is itself only conditioned to ~4e-12 (hit points) and ~7e-9 (directions)
through the four grazing-incidence quadrics (the gradient 2ax+g cancels ~7 digits at x = 146 m).
north_star's 1e-12 gate is therefore a statement against the reference's FP64 path, not against
exact arithmetic.  They also pin the per-ray NaN flavour of a miss (III_I:296-301).
"""
import numpy as np
import pytest

import oracle
from conftest import per_ray_rel


def _chain_inputs(g):
    return list(g["coeffs"]), [bool(b) for b in g["negative"]], g["plane"], g["ray"], g["source"]


def _worst(out_points, out_dirs, det, g):
    pts = max(per_ray_rel(out_points[k], g[f"P{k}"]) for k in range(4))
    dirs = max(max(per_ray_rel(out_dirs[0][k], g[f"N{k}"]), per_ray_rel(out_dirs[1][k], g[f"R{k}"])) for k in range(4))
    return max(pts, per_ray_rel(det, g["det"])), dirs


def test_fp64_reference_arithmetic_vs_mpmath_truth(golden):
    g = golden("ray_mp_ref")
    coeffs, neg, plane, ray, src = _chain_inputs(g)
    out = oracle.trace_chain(coeffs, neg, plane, ray, src)
    pts, dirs = _worst(out["points"], (out["normals"], out["reflect"]), out["det"], g)
    print(f"FP64 (NumPy order) vs 30-digit mpmath, 4 mirrors + plane: points {pts:.2e}, directions {dirs:.2e}")
    assert per_ray_rel(out["points"][0], g["P0"]) <= 1e-14     # first mirror: well conditioned
    assert pts <= 2e-11 and dirs <= 5e-8                        # the chain: conditioning, not a bug


@pytest.mark.gpu
def test_kernel_equals_fp64_reference_and_tracks_mpmath(golden):
    import torch
    assert torch.cuda.is_available()
    import akbraytracing_b200 as akb
    g = golden("ray_mp_ref")
    coeffs, neg, plane, ray, src = _chain_inputs(g)
    out = akb.trace_chain(coeffs, neg, plane, ray, src, want_normals=True, want_reflects=True)
    ref = oracle.trace_chain(coeffs, neg, plane, ray, src)
    for k in range(4):   # bit-identical to the reference's FP64 arithmetic ...
        assert np.array_equal(out["points"][k], ref["points"][k])
        assert np.array_equal(out["normals"][k], ref["normals"][k])
        assert np.array_equal(out["reflects"][k], ref["reflect"][k])
    assert np.array_equal(out["det"], ref["det"])
    pts, dirs = _worst(out["points"], (out["normals"], out["reflects"]), out["det"], g)
    assert pts <= 2e-11 and dirs <= 5e-8   # ... hence exactly as close to exact arithmetic as the reference is


@pytest.mark.gpu
def test_per_ray_nan_flavour_matches_mpmath_twin(golden):
    """mpmath flavour: only the missing ray is NaN (III_I:296-301); NumPy flavour: everything is."""
    import akbraytracing_b200 as akb
    g = golden("ray_mp_ref")
    co = g["coeffs"][0]
    p, _, _ = akb.intersect_reflect(co, g["miss_ray"], g["miss_source"], check=False)
    assert np.isnan(g["miss_P0"][:, 5]).all()
    assert np.array_equal(np.isnan(p), np.isnan(g["miss_P0"]))
    ok = ~np.isnan(g["miss_P0"][0])
    assert per_ray_rel(p[:, ok], g["miss_P0"][:, ok]) <= 1e-14
    p_all = akb.mirr_ray_intersection(co, g["miss_ray"], g["miss_source"])
    assert np.isnan(p_all).all()  # ER3D:31-33 / BIG:456-459
