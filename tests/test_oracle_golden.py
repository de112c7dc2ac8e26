"""Pin the CPU oracle (oracle/) against outputs of the REFERENCE ITSELF.

The fixtures under tests/golden/*.npz were produced by tests/golden/make_golden.py, which
runs the reference's own functions in the build container.  CPU-only, no GPU.
"""
import numpy as np
import pytest

import oracle
from oracle import numpy_port as npp
from conftest import per_ray_rel, rel_l2

FRESNEL_CASES = ["patch_euv", "patch_xray", "ragged", "one_source", "one_detector",
                 "src_to_m1_euv", "src_to_m1_xray", "mirror_to_mirror", "coincident"]


@pytest.mark.parametrize("name", FRESNEL_CASES)
def test_fresnel_c_oracle_matches_reference(golden, name):
    c = golden("fresnel_ref").group(name)
    with np.errstate(all="ignore"):
        got = oracle.fresnel_sum(c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], float(c["k"]), c["ds"])
    ref = c["ref"]  # forward_propagation_numpy_batch (numba), CPU0402:87-124
    finite = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), finite)  # same inf/nan pattern
    # same operation order, no FMA: the C restatement reproduces numba to the last bit or two
    assert rel_l2(got[finite], ref[finite]) <= 1e-14
    # the reference's own pure-NumPy twin (CPU0402:54-63) agrees with its numba path
    assert rel_l2(c["ref_numpy"][finite], ref[finite]) <= 1e-13


@pytest.mark.parametrize("name", FRESNEL_CASES)
def test_fresnel_numpy_port_matches_reference(golden, name):
    c = golden("fresnel_ref").group(name)
    with np.errstate(all="ignore"):
        got = npp.fresnel_sum(c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], float(c["k"]), c["ds"])
    finite = np.isfinite(c["ref"])
    assert rel_l2(got[finite], c["ref"][finite]) <= 1e-13


def test_fresnel_threads_do_not_change_result(golden):
    c = golden("fresnel_ref").group("patch_euv")
    a = oracle.fresnel_sum(c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], float(c["k"]), c["ds"], nthreads=1)
    b = oracle.fresnel_sum(c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], float(c["k"]), c["ds"], nthreads=4)
    assert np.array_equal(a, b)


def test_wavefield_golden_is_k_from_lambda(golden):
    g = golden("fresnel_ref")
    c = g.group("patch_euv")
    k = 2.0 * np.pi / float(g["wavefield/lambda"])  # CPU0402:39
    got = oracle.fresnel_sum(c["x"], c["y"], c["z"], c["sx"], c["sy"], c["sz"], c["u"], k, c["ds"])
    assert rel_l2(got, g["wavefield/u"]) <= 1e-14


# ------------------------------------------------------------------ path B

def test_single_mirror_matches_er3d(golden):
    g = golden("ray_er3d_ref")
    co, ray, src = g["single/coeffs"], g["single/ray"], g["single/source"]
    for mod in (oracle, npp):
        p = mod.mirr_ray_intersection(co, ray, src)
        n = mod.norm_vector(co, p)
        r = mod.reflect_ray(ray, n)
        assert np.array_equal(p, g["single/points"])
        assert np.array_equal(n, g["single/N_ell"])
        assert np.array_equal(r, g["single/reflect"])
    # the 'z' cylinder and the negative root
    assert np.array_equal(oracle.mirr_ray_intersection(g["single_z/coeffs"], ray, src), g["single_z/points"])
    assert np.array_equal(oracle.mirr_ray_intersection(co, ray, src, negative=True), g["negative/points"])


def test_plane_points_match_er3d(golden):
    g = golden("ray_er3d_ref")
    pos, delta = float(g["single/plane_position"]), float(g["single/plane_delta"])
    for off, key in ((0.0, "points0"), (delta, "points1"), (-delta, "points2")):
        co = np.zeros(10)
        co[6] = 1.0
        co[9] = -pos + off  # ER3D:248-261
        for mod in (oracle, npp):
            got = mod.plane_ray_intersection(co, g["single/reflect"], g["single/points"])
            assert np.array_equal(got, g["single/" + key])


def test_general_quadric_matches_er3d(golden):
    g = golden("ray_er3d_ref")
    co, ray, src = g["general/coeffs"], g["single/ray"], g["general/source"]
    for mod in (oracle, npp):
        p = mod.mirr_ray_intersection(co, ray, src)
        n = mod.norm_vector(co, p)
        r = mod.reflect_ray(ray, n)
        assert per_ray_rel(p, g["general/points"]) <= 1e-15
        assert per_ray_rel(n, g["general/normal"]) <= 1e-15
        assert per_ray_rel(r, g["general/reflect"]) <= 1e-15
    assert np.array_equal(oracle.mirr_ray_intersection(co, ray, src), g["general/points"])


def test_miss_is_all_nan(golden):
    g = golden("ray_er3d_ref")
    for mod in (oracle, npp):
        with np.errstate(all="ignore"):
            p = mod.mirr_ray_intersection(g["single/coeffs"], g["miss/ray"], g["miss/source"])
            n = mod.norm_vector(g["single/coeffs"], p)
            r = mod.reflect_ray(g["miss/ray"], n)
        assert np.isnan(g["miss/points"]).all() and np.isnan(p).all()  # ER3D:31-33
        assert np.isnan(g["miss/normal"]).all() and np.isnan(n).all()
        assert np.isnan(g["miss/reflect"]).all() and np.isnan(r).all()


def test_normalize_all_or_nothing(golden):
    g = golden("ray_er3d_ref")
    for mod in (oracle, npp):
        assert np.array_equal(mod.normalize_vector(g["normalize/in"]), g["normalize/out"])
        # one zero column -> input returned unchanged (ER3D:57-59)
        assert np.array_equal(mod.normalize_vector(g["normalize/in_zero"]), g["normalize/out_zero"])
    assert np.array_equal(g["normalize/out_zero"], g["normalize/in_zero"])


@pytest.mark.parametrize("kind,n_mirrors", [("akb", 4), ("kb", 2)])
def test_chain_matches_driver(golden, kind, n_mirrors):
    """Every intersect/normal/reflect/plane call of the kept pass of the reference driver
    (BIG:2881-2905 AKB Wolter III+I, BIG:11039-11054 KB)."""
    g = golden(f"chain_{kind}_ref")
    n = g["tan_h"].shape[0]
    raw = np.vstack([np.ones(n * n), np.tile(g["tan_h"], n), np.repeat(g["tan_v"], n)])
    ray0 = oracle.normalize_vector(raw)
    assert np.array_equal(ray0, g["ray0"])
    src = np.repeat(g["source_point"][:, None], n * n, axis=1)
    out = oracle.trace_chain(list(g["coeffs"]), list(g["negative"]), g["plane"], ray0, src)
    assert list(g["negative"]) == ([False, False, False, True] if kind == "akb" else [False, False])
    for k in range(n_mirrors):
        assert np.array_equal(out["points"][k], g[f"P{k}"])
        assert np.array_equal(out["normals"][k], g[f"N{k}"])
        assert np.array_equal(out["reflect"][k], g[f"R{k}"])
        assert np.allclose(out["dist"][k], g[f"dist{k}"], rtol=1e-15, atol=0)
    assert np.array_equal(out["det"], g["det"])


def test_calc_dS_matches_driver(golden):
    g = golden("dS_ref")
    got = oracle.calc_dS(g["points"], int(g["nV"]), int(g["nH"]))
    assert np.allclose(got, g["dS"], rtol=1e-12, atol=0)  # BIG:13418-13473
    got3 = oracle.calc_dS(g["points3"], 3, 3)
    assert np.allclose(got3, g["dS3"], rtol=1e-12, atol=0)
    k = golden("chain_kb_ref")
    n = k["tan_h"].shape[0]
    assert np.allclose(oracle.calc_dS(k["dS_points"], n, n), k["dS"], rtol=1e-9, atol=0)


def test_psf_numpy_port_matches_reference(golden):
    g = golden("psf_ref")
    for tag, kw in (("plain", dict(pad_factor=2)), ("hann", dict(pad_factor=3, window="hann", pupil_dy_m=1.5e-4))):
        I, x, y, E = npp.compute_psf_fft(g["opd"], g["amp"], 13.5e-9, 1e-4, 0.3, return_efield=True, **kw)
        assert np.allclose(I, g[f"{tag}/I"], rtol=1e-12, atol=1e-18)
        assert np.array_equal(x, g[f"{tag}/x"]) and np.array_equal(y, g[f"{tag}/y"])
        assert np.allclose(E, g[f"{tag}/E"], rtol=1e-10, atol=1e-16)


def test_array_split_bounds():
    for total in (0, 1, 7, 4096, 4097):
        for parts in (1, 2, 3, 8):
            b = npp.array_split_bounds(total, parts)
            ref = np.array_split(np.arange(total), parts)
            assert [n for _, n in b] == [len(r) for r in ref]
            assert all(s == (r[0] if len(r) else s) for (s, _), r in zip(b, ref))


# ------------------------------------------------------------------ rows f-1 / f-4: reference END-TO-END runs

@pytest.mark.parametrize("kind,K", [("kb", 2), ("akb", 4)])
def test_oracle_stage_chain_reproduces_the_reference_script(golden, kind, K):
    """The unmodified Wavecalc_raytrace_fromData_CPU0402.py was run as __main__ on a folder written by the
    reference's saveWaveData (tests/golden/make_golden.py stagechain_ref).  The oracle, chained the way the script
    chains its stages (CPU0402:247-375: ds of a surface = row 3 of its points file, focal grid stretched by 2,
    defocused grid 'stretched' by 1), reproduces every complex_data_*.npz bit for bit."""
    g = golden("stagechain_ref")
    from akbraytracing_b200.stagechain import parse_conditions
    cond = parse_conditions(str(g[f"{kind}/handoff/conditions_txt"]))
    assert cond["option_AKB"] == (K == 4) and cond["option_HighNA"] is True
    nV, nH = (int(v) for v in g[f"{kind}/ray_num"])
    assert (cond["ray_num_H1"], cond["ray_num_V1"], cond["pix_y"], cond["pix_z"]) == (nH, nV, nH, nV)
    k = 2.0 * np.pi / np.float64(13.5e-9)
    src = g[f"{kind}/handoff/points_source"]
    u = np.ones(1, complex); bx, by, bz = (np.array([v]) for v in src); ds = np.ones(1)
    for i in range(K):
        pts = g[f"{kind}/handoff/points_M{i + 1}"]
        u = oracle.fresnel_sum(pts[0], pts[1], pts[2], bx, by, bz, u, k, ds)
        assert np.array_equal(u, g[f"{kind}/wavecalc/complex_data_M{i + 1}"]), f"stage M{i + 1}"
        bx, by, bz, ds = pts[0], pts[1], pts[2], pts[3]
    for name, grid_name, scale in (("Image", "points_gridImage", 2.0), ("Image2", "points_gridDefocus", 1.0)):
        grid = np.array(g[f"{kind}/handoff/{grid_name}"])
        for r in range(3):
            m = np.mean(grid[r, :])
            grid[r, :] = (grid[r, :] - m) * scale + m if scale != 1.0 else (grid[r, :] - m) + m
        saved = "points_gridImage" if name == "Image" else "points_gridImage2"
        assert np.array_equal(grid, g[f"{kind}/wavecalc/{saved}"])
        f = oracle.fresnel_sum(grid[0], grid[1], grid[2], bx, by, bz, u, k, ds)
        assert np.array_equal(f, g[f"{kind}/wavecalc/complex_data_{name}"]), name
    # the hand-off arrays themselves: dS row = oracle.calc_dS of the cloud, rows 0-2 = what the tracer returned
    for i in range(K):
        pts = g[f"{kind}/handoff/points_M{i + 1}"]
        assert np.array_equal(pts[:3], g[f"{kind}/tracer_M{i + 1}"])
        assert np.allclose(pts[3].reshape(nV, nH), oracle.calc_dS(pts[:3], nV, nH), rtol=1e-12, atol=0)


def test_oracle_ray_wave_tail_reproduces_the_reference_driver(golden):
    """plot_result_debug(p,'ray_wave') run from the reference (make_golden.py ray_wave_ref): rotation into the
    detector frame, both planes, totalDist / totalDist2 from the NumPy restatement."""
    w = golden("ray_wave_ref")
    k = "akb"
    got = npp.wavefront_opl(w[f"{k}/last_point"], w[f"{k}/last_dir"], w[f"{k}/dist"], float(w[f"{k}/plane_x"]),
                                   float(w[f"{k}/plane2_x"]), float(w[f"{k}/theta_y"]), float(w[f"{k}/theta_z"]), w[f"{k}/pivot"])
    for key, ref in (("point", "point_rot"), ("dir", "dir_rot"), ("det", "det"), ("det2", "det2"), ("opl", "opl"), ("opl2", "opl2")):
        assert np.array_equal(got[key], w[f"{k}/{ref}"]), key
    # segment lengths of the chain = what the fused kernel's dist output must give (oracle.trace_chain)
    c = golden("chain_akb_ref")
    assert w[f"{k}/dist"].shape[0] == 4
