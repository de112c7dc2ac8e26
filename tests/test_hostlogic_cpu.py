"""CPU-only tests of host logic that sits beside the hot path: hand-off file parsing, the PSF
restatement on torch (run on the CPU device here, on cuda in test_gpu_parity.py)."""
import numpy as np
import pytest


def test_parse_conditions_follows_reference_substring_rules():
    from akbraytracing_b200.stagechain import parse_conditions
    text = ("Conditions\n====================\ntime: 20250404_101112\n"
            "grid pitch_y: 3.9e-09\ngrid size_y: 2e-06\n"
            "grid pix_y: 64\ngrid pix_z: 48\ngrid pix_H1: 101\ngrid pix_V1: 99\ngrid pix_H2: 51\ngrid pix_V2: 49\n"
            "option_AKB: True\noption_HighNA: False\ndefocusForWave: 0.001\n")
    c = parse_conditions(text)
    assert (c["pix_y"], c["pix_z"]) == (64, 48)
    assert (c["ray_num_H1"], c["ray_num_V1"], c["ray_num_H2"], c["ray_num_V2"]) == (101, 99, 51, 49)  # CPU0402:224-231
    assert c["option_AKB"] is True and c["option_HighNA"] is False
    # older files only carry pix_y / pix_z: they seed the ray counts too (CPU0402:210-217)
    c = parse_conditions("grid pix_y: 33\ngrid pix_z: 31\noption_AKB: False\noption_HighNA: True\n")
    assert (c["ray_num_H1"], c["ray_num_H2"], c["ray_num_V1"], c["ray_num_V2"]) == (33, 33, 31, 31)
    # 'grid pix_H:' / 'grid pix_V:' cross-assign (CPU0402:218-223)
    c = parse_conditions("grid pix_H: 7\ngrid pix_V: 9\n")
    assert (c["ray_num_H1"], c["ray_num_V2"], c["ray_num_V1"], c["ray_num_H2"]) == (7, 7, 9, 9)


def test_focal_grids_layout():
    from akbraytracing_b200.stagechain import focal_grids
    det = np.array([[1.0, 1.2, 1.1], [-2e-6, 4e-6, 0.0], [5e-6, 7e-6, 6e-6]])
    grid, yg, zg = focal_grids(det, 5, 3)
    assert grid.shape == (3, 15) and np.allclose(grid[0], 1.1)
    assert np.allclose(yg, np.linspace(1e-6 - 1e-6, 1e-6 + 1e-6, 5)) and np.allclose(zg, np.linspace(5e-6, 7e-6, 3))
    assert np.array_equal(grid[1, :5], yg) and np.array_equal(grid[2, ::5], zg)  # np.meshgrid(y, z) order: y fastest


@pytest.mark.parametrize("tag,kw", [("plain", dict(pad_factor=2)),
                                    ("hann", dict(pad_factor=3, window="hann", pupil_dy_m=1.5e-4))])
def test_psf_matches_reference_golden_on_cpu_device(golden, tag, kw):
    from akbraytracing_b200.psf import compute_psf_fft, psf_to_db
    g = golden("psf_ref")
    I, x, y, E = compute_psf_fft(g["opd"], g["amp"], 13.5e-9, 1e-4, 0.3, return_efield=True, device="cpu", **kw)
    assert np.allclose(I, g[f"{tag}/I"], rtol=1e-10, atol=1e-16)
    assert np.array_equal(x, g[f"{tag}/x"]) and np.array_equal(y, g[f"{tag}/y"])
    assert np.allclose(E, g[f"{tag}/E"], rtol=1e-9, atol=1e-15)
    assert psf_to_db(I).min() >= -60.0 - 1e-9
    with pytest.raises(ValueError):
        compute_psf_fft(g["opd"], g["amp"][:-1], 13.5e-9, 1e-4, 0.3, device="cpu")
    with pytest.raises(ValueError):
        compute_psf_fft(g["opd"], g["amp"], 13.5e-9, 1e-4, 0.3, window="hamming", device="cpu")


def test_batched_psf_matches_reference_golden_on_cpu_device(golden):
    """compute_psf_fft_batch (through-focus PSF stack, one batched fft2) against the reference's psf_fft outputs."""
    from akbraytracing_b200.throughfocus import compute_psf_fft_batch
    g = golden("psf_ref")
    opd, amp = np.stack([g["opd"], g["opd"]]), np.stack([g["amp"], g["amp"] * 2.0])
    I, x, y, E = compute_psf_fft_batch(opd, amp, 13.5e-9, 1e-4, 0.3, pad_factor=2, return_efield=True, device="cpu")
    assert I.shape[0] == 2 and np.allclose(I[0], g["plain/I"], rtol=1e-10, atol=1e-16)
    assert np.allclose(I[1], g["plain/I"], rtol=1e-10, atol=1e-16)   # each plane normalised by its own peak
    assert np.array_equal(x, g["plain/x"]) and np.array_equal(y, g["plain/y"])
    assert np.allclose(E[0], g["plain/E"], rtol=1e-9, atol=1e-15)
    with pytest.raises(ValueError):
        compute_psf_fft_batch(g["opd"], g["amp"], 13.5e-9, 1e-4, 0.3, device="cpu")   # needs (P, ny, nx)


def test_auto_phase_mode_per_stage():
    """run_stage_chain(phase_mode='auto'): FAITHFUL where the other modes would leave the 1e-7 margin (the 146 m source ->
    M1 stage), REFERENCED for a detector plane, EXACT for an irregular detector set (a mirror)."""
    from akbraytracing_b200 import _lib
    from akbraytracing_b200.stagechain import auto_phase_mode
    k = 2 * np.pi / 13.5e-9
    rng = np.random.default_rng(0)
    src = np.zeros((3, 1))
    m1 = np.vstack([146.0 + rng.uniform(-0.03, 0.03, 50), rng.uniform(-2e-3, 2e-3, 50), rng.uniform(-2e-3, 2e-3, 50)])
    m2 = m1 + np.array([[0.12], [0.0], [0.0]])
    yy, zz = np.meshgrid(np.linspace(-1e-6, 1e-6, 8), np.linspace(-1e-6, 1e-6, 8))
    plane = np.vstack([np.full(64, 146.3), yy.ravel(), zz.ravel()])
    assert auto_phase_mode(k, m1, src) == _lib.PHASE_FAITHFUL        # k r 2^-52 = 1.5e-5
    assert auto_phase_mode(k, m2, m1) == _lib.PHASE_EXACT            # 0.18 m between two mirrors: 1.9e-8
    assert auto_phase_mode(k, plane, m2) == _lib.PHASE_REFERENCED    # a plane x = const
    assert auto_phase_mode(10 * k, plane, m2, tol=1e-9) == _lib.PHASE_FAITHFUL


def test_handoff_downsampling_is_every_second_sample():
    """downsample_array_3_n (AKB_raytrace_20250312.py:13336-13356): count = flag // 2 passes of [::2] per axis."""
    from akbraytracing_b200.stagechain import _downsample
    nV, nH = 9, 13
    cloud = np.arange(3 * nV * nH, dtype=np.float64).reshape(3, nV * nH)
    g = cloud.reshape(3, nV, nH)
    out, sv, sh = _downsample(cloud, nV, nH, 0, 0)
    assert (sv, sh) == (nV, nH) and np.array_equal(out, cloud)
    out, sv, sh = _downsample(cloud, nV, nH, 2, 0)
    assert (sv, sh) == (nV, 7) and np.array_equal(out, g[:, :, ::2].reshape(3, -1))
    out, sv, sh = _downsample(cloud, nV, nH, 4, 2)
    assert (sv, sh) == (5, 4) and np.array_equal(out, g[:, ::2, ::4].reshape(3, -1))


def test_to_host_of_a_cpu_tensor_is_a_plain_array():
    """_lib.to_host: page-locked staging is for CUDA tensors only; anything else converts like .cpu().numpy()."""
    import torch
    from akbraytracing_b200 import _lib
    t = torch.arange(6, dtype=torch.float64).reshape(2, 3).requires_grad_(False)
    a = _lib.to_host(t)
    assert isinstance(a, np.ndarray) and a.shape == (2, 3) and a.dtype == np.float64
    assert np.array_equal(a, np.arange(6.0).reshape(2, 3))
    c = _lib.to_host(torch.ones(4, dtype=torch.complex128))
    assert c.dtype == np.complex128 and c.sum() == 4
