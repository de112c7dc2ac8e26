"""Stage-chain driver and hand-off file formats.

The reference hands the traced mirror point clouds from the ray drivers to the Wavecalc scripts
through a folder of files (writer: ``saveWaveData``, AKB_raytrace_20250312.py:13475-13654; reader
and stage chain: Wavecalc_raytrace_fromData_CPU0402.py:190-377):

    points_source.npy            (3,)      the point source
    points_M1.npy .. points_M4.npy (4, N)  x, y, z, dS per mirror sample (M3/M4 only if option_AKB)
    points_gridImage.npy         (3, M)    focal-plane grid
    points_gridDefocus.npy       (3, M)    defocused grid
    calculation_conditions.txt             "grid pix_y: ..", "grid pix_H1: ..", "option_AKB: ..", ...
    complex_data_<stage>.npz['data']       complex128 field per stage (output)

Mirror order: the chain follows the FILE order M1 -> M2 -> M3 -> M4 exactly as the reference's reader
does (CPU0402:276-328).  The reference's writer names the files in alternating order (M1 = vmirr_hyp,
M2 = hmirr_hyp, M3 = vmirr_ell, M4 = hmirr_ell; BIG:13479, 13523-13554), which for the Wolter III+I
layout is not the physical beam path.  This driver is deliberately bug-compatible: it
reproduces what the reference computes from a given folder; a caller who wants the physical order
passes ``mirror_clouds`` to ``write_handoff`` in that order.

``run_stage_chain`` walks source -> M1 -> M2 (-> M3 -> M4) -> Image -> Image2 with every field
resident in HBM (one ``akb_fresnel_sum`` per stage, no file or host round trip in between) and
writes the same ``complex_data_*.npz`` files.  ``write_handoff`` produces the folder from traced
clouds with ``calc_dS`` evaluated on the device.
"""
from __future__ import annotations

import os

import numpy as np

from . import handoff
from .wavecalc import WaveField3D

__all__ = ["parse_conditions", "load_handoff", "write_handoff", "run_stage_chain", "focal_grids", "auto_phase_mode"]

WAVELENGTH_HIGH_NA = 13.5e-9         # CPU0402:243
WAVELENGTH_LOW_NA = 13.5e-9 * 1e-1   # CPU0402:245


def parse_conditions(text: str) -> dict:
    """The substring matching of CPU0402:208-237, including its quirks: 'grid pix_y' also seeds
    ray_num_H1/H2, 'grid pix_H' seeds H1 and V2, 'grid pix_V' seeds V1 and H2."""
    c = {}
    for line in text.splitlines():
        if ":" not in line:
            continue
        val = line.split(":")[1].strip()
        if "grid pix_y:" in line:
            c["pix_y"] = c["ray_num_H1"] = c["ray_num_H2"] = int(val)
        elif "grid pix_z:" in line:
            c["pix_z"] = c["ray_num_V1"] = c["ray_num_V2"] = int(val)
        elif "grid pix_H:" in line:
            c["ray_num_H1"] = c["ray_num_V2"] = int(val)
        elif "grid pix_V:" in line:
            c["ray_num_V1"] = c["ray_num_H2"] = int(val)
        elif "grid pix_H1:" in line:
            c["ray_num_H1"] = int(val)
        elif "grid pix_V1:" in line:
            c["ray_num_V1"] = int(val)
        elif "grid pix_H2:" in line:
            c["ray_num_H2"] = int(val)
        elif "grid pix_V2:" in line:
            c["ray_num_V2"] = int(val)
        elif "option_AKB:" in line:
            c["option_AKB"] = val.lower() == "true"
        elif "option_HighNA:" in line:
            c["option_HighNA"] = val.lower() == "true"
    return c


def load_handoff(folder: str) -> dict:
    """Read a hand-off folder (CPU0402:192-195, 276, 293-298, 314, 351)."""
    with open(os.path.join(folder, "calculation_conditions.txt"), "r") as fh:
        cond = parse_conditions(fh.read())
    out = {"conditions": cond, "source": np.load(os.path.join(folder, "points_source.npy"))}
    names = ["M1", "M2"] + (["M3", "M4"] if cond.get("option_AKB") else [])
    out["mirrors"] = [np.load(os.path.join(folder, f"points_{n}.npy")) for n in names]
    out["gridImage"] = np.load(os.path.join(folder, "points_gridImage.npy"))
    p = os.path.join(folder, "points_gridDefocus.npy")
    out["gridDefocus"] = np.load(p) if os.path.exists(p) else None
    return out


def focal_grids(det, size_h, size_v, ysize=1e-6, zsize=1e-6):
    """Detector grid of saveWaveData (BIG:13570-13591): y, z centred on the traced spot,
    +-ysize / +-zsize, meshgrid(y, z) flattened, x = mean traced x.  det: (3, N) array."""
    det = np.asarray(det)
    y, z = det[1], det[2]
    y_grid = np.linspace((y.min() + y.max()) / 2 - ysize, (y.min() + y.max()) / 2 + ysize, size_h)
    z_grid = np.linspace((z.min() + z.max()) / 2 - zsize, (z.min() + z.max()) / 2 + zsize, size_v)
    yy, zz = np.meshgrid(y_grid, z_grid)
    x = np.full(yy.size, det[0].mean())
    return np.vstack([x, yy.ravel(), zz.ravel()]), y_grid, z_grid


def _downsample(cloud, nV, nH, down_h, down_v):
    """downsample_array_3_n (BIG:13336-13356): keep every 2nd column down_h//2 times, every 2nd row down_v//2 times."""
    g = cloud[:3].reshape(3, nV, nH)
    for _ in range(int(down_h) // 2):
        g = g[:, :, ::2]
    for _ in range(int(down_v) // 2):
        g = g[:, ::2, :]
    return g.reshape(3, -1), g.shape[1], g.shape[2]


def write_handoff(folder, source_point, mirror_clouds, grid_shape, det, det_defocus=None, option_AKB=None,
                  option_HighNA=True, focus_shape=None, defocus=0.0, ysize=1e-6, zsize=1e-6,
                  downsample=(0, 0, 0, 0, 0, 0), initial_params=None, timestamp=None, option_2mirror=True,
                  option_avrgsplt=False):
    """Write the folder saveWaveData writes (BIG:13475-13654) from traced clouds: same files, same array
    contents (dS from the device kernel akb_calc_ds), same calculation_conditions.txt lines in the same order.

    mirror_clouds: list of (3, nV*nH) hit-point arrays (NumPy or torch CUDA) in FILE order M1.. (the reference's
    alternating naming: vmirr_hyp, hmirr_hyp, vmirr_ell, hmirr_ell, BIG:13479); grid_shape: (nV, nH) of the ray
    grid; det / det_defocus: the tracer's detector points (3, N) on the focal / defocused plane;
    defocus: the reference's signed ``defocusForWave`` -- the defocused grid is written only when
    ``abs(defocus) > 1e-9`` (BIG:13593) and its half-size is ``2e-7 + defocus*0.082*2`` (0.01 for low NA) with the
    sign kept (BIG:13594-13599); downsample: (h1, v1, h2, v2, h_f, v_f) as the module globals of BIG:13489-13498
    (odd mirrors use h1/v1, even ones h2/v2).  ``focus_shape`` (extension, not in the reference) overrides the
    focal grid's pixel counts, which the reference takes from the ray grid."""
    os.makedirs(folder, exist_ok=True)
    nV, nH = (int(v) for v in grid_shape)
    if option_AKB is None:
        option_AKB = len(mirror_clouds) == 4

    def host(a):
        return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)

    dh1, dv1, dh2, dv2, dhf, dvf = downsample
    np.save(os.path.join(folder, "points_source.npy"), np.asarray(host(source_point), dtype=np.float64).reshape(3, -1)[:, 0])
    sizes = {}
    for k, cloud in enumerate(mirror_clouds):
        cl, sv, sh = _downsample(host(cloud), nV, nH, *((dh1, dv1) if k % 2 == 0 else (dh2, dv2)))
        sizes[k % 2] = (sv, sh)
        dS = host(handoff.calc_dS(cl, sv, sh))                                # BIG:13520-13551
        np.save(os.path.join(folder, f"points_M{k + 1}.npy"), np.vstack((cl, dS.flatten())))
    det_d, fV, fH = _downsample(host(det), nV, nH, dhf, dvf)
    if focus_shape is not None:
        fV, fH = (int(v) for v in focus_shape)
    grid, y_grid, z_grid = focal_grids(det_d, fH, fV, ysize, zsize)           # BIG:13570-13591
    np.save(os.path.join(folder, "points_gridImage.npy"), grid)
    if det_defocus is not None and abs(defocus) > 1e-9:                      # BIG:13593
        ys = 2e-7 + defocus * (0.082 if option_HighNA else 0.01) * 2          # BIG:13594-13599 (signed)
        det2_d, _, _ = _downsample(host(det_defocus), nV, nH, dhf, dvf)
        grid2, _, _ = focal_grids(det2_d, fH, fV, ys, ys)
        np.save(os.path.join(folder, "points_gridDefocus.npy"), grid2)
    (sv1, sh1), (sv2, sh2) = sizes.get(0, (nV, nH)), sizes.get(1, (nV, nH))
    with open(os.path.join(folder, "calculation_conditions.txt"), "w") as fh:  # BIG:13630-13654
        fh.write("Conditions\n")
        fh.write("====================\n")
        if timestamp is not None:
            fh.write(f"time: {timestamp}\n")
        if initial_params is not None:
            p = np.asarray(initial_params)
            fh.write(f"params 0-1: {p[0:2]}\n")
            fh.write(f"params 2-7: {p[2:8]}\n")
            fh.write(f"params 8-13: {p[8:14]}\n")
            fh.write(f"params 14-19: {p[14:20]}\n")
            fh.write(f"params 20-26: {p[20:26]}\n")
        fh.write(f"grid pitch_y: {y_grid[1] - y_grid[0]}\n")
        fh.write(f"grid pitch_z: {z_grid[1] - z_grid[0]}\n")
        fh.write(f"grid size_y: {np.max(y_grid) - np.min(y_grid)}\n")
        fh.write(f"grid size_z: {np.max(z_grid) - np.min(z_grid)}\n")
        fh.write(f"grid pix_y: {fH}\n")
        fh.write(f"grid pix_z: {fV}\n")
        fh.write(f"grid pix_H1: {sh1}\n")
        fh.write(f"grid pix_V1: {sv1}\n")
        fh.write(f"grid pix_H2: {sh2}\n")
        fh.write(f"grid pix_V2: {sv2}\n")
        fh.write(f"option_AKB: {bool(option_AKB)}\n")
        fh.write(f"option_HighNA: {bool(option_HighNA)}\n")
        fh.write(f"defocusForWave: {defocus}\n")
        fh.write(f"calc both mirrors?: {bool(option_2mirror)}\n")
        fh.write(f"option_avrgsplt: {bool(option_avrgsplt)}\n")
        fh.write("====================\n")


def auto_phase_mode(k, front_xyz, back_xyz, tol=1e-7):
    """Phase arithmetic for one stage when exact reference roundings are not required.  The non-faithful modes deviate
    from the reference's roundings by about k*r*2^-52 rad per term (the reference's own rounding noise); they are
    chosen when that stays below ``tol`` for the largest distance between the bounding boxes of the two surfaces:
    AKB_PHASE_REFERENCED for a detector plane x = const (its row expansion makes it the fastest loop there: 20.75 FP64
    instructions per pair), AKB_PHASE_EXACT for an irregular detector set such as a mirror (26 instead of 29),
    else AKB_PHASE_FAITHFUL."""
    from ._lib import PHASE_EXACT, PHASE_FAITHFUL, PHASE_REFERENCED
    f, b = np.asarray(front_xyz)[:3], np.asarray(back_xyz)[:3]
    span = np.maximum(np.abs(f.max(axis=1) - b.min(axis=1)), np.abs(b.max(axis=1) - f.min(axis=1)))
    r_max = float(np.sqrt((span ** 2).sum()))
    if abs(k) * r_max * 2.0 ** -52 > tol:
        return PHASE_FAITHFUL
    return PHASE_REFERENCED if f.shape[1] > 1 and float(np.ptp(f[0])) == 0.0 else PHASE_EXACT


def run_stage_chain(folder, out_dir=None, device="cuda", image_scale=2.0, resume_dir=None, verbose=False,
                    phase_mode="faithful", keep_on_device=False):
    """The stage chain of CPU0402:247-375 with device-resident fields: source -> M1 -> M2 (-> M3 -> M4) ->
    Image -> Image2, one ``akb_fresnel_sum`` per stage, nothing leaves HBM between stages.

    Returns dict(stage name -> complex128 field): NumPy arrays, or the resident torch tensors with
    ``keep_on_device=True`` (then no field is copied to the host unless ``out_dir`` asks for the files).
    ``out_dir``: write what the reference script writes into its output_<timestamp> folder --
    complex_data_<stage>.npz, the points_*.npy copies, the stretched points_gridImage.npy, points_gridImage2.npy
    and calculation_conditions.txt.  ``resume_dir``: like the reference, a ``complex_data_Mk.npz`` found there is
    loaded instead of computing that stage (CPU0402:261-265).  ``image_scale``: the reference stretches the
    focal grid by 2 about its mean (CPU0402:330-334).  ``phase_mode``: 'faithful' (the reference's roundings,
    default), 'exact', 'referenced', or 'auto' (per stage, see auto_phase_mode)."""
    import contextlib
    import io
    import shutil
    from . import _lib
    h = load_handoff(folder)
    cond = h["conditions"]
    wavelength = WAVELENGTH_HIGH_NA if cond.get("option_HighNA") else WAVELENGTH_LOW_NA  # CPU0402:242-245
    quiet = contextlib.nullcontext() if verbose else contextlib.redirect_stdout(io.StringIO())
    modes = {"faithful": _lib.PHASE_FAITHFUL, "exact": _lib.PHASE_EXACT, "referenced": _lib.PHASE_REFERENCED}
    if phase_mode != "auto" and phase_mode not in modes:
        raise ValueError("phase_mode must be 'faithful', 'exact', 'referenced' or 'auto'")
    results = {}
    if out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        np.save(os.path.join(out_dir, "points_source.npy"), h["source"])                     # CPU0402:204-205
        shutil.copy(os.path.join(folder, "calculation_conditions.txt"), os.path.join(out_dir, "calculation_conditions.txt"))

    def save(name, field):
        on_dev = hasattr(field.u, "cpu")
        if out_dir is not None or not (keep_on_device and on_dev):
            arr = field.u.cpu().numpy() if on_dev else np.asarray(field.u)
        results[name] = field.u if (keep_on_device and on_dev) else arr
        if out_dir is not None:
            np.savez_compressed(os.path.join(out_dir, f"complex_data_{name}.npz"), data=arr)

    def resumed(name):
        if resume_dir is None:
            return None
        p = os.path.join(resume_dir, f"complex_data_{name}.npz")
        if os.path.exists(p):
            with np.load(p) as z:
                return z["data"]
        return None

    def propagate(field, back, front_xyz, back_xyz):
        k = 2.0 * np.pi / np.float64(wavelength)
        field.phase_mode = auto_phase_mode(k, front_xyz, back_xyz) if phase_mode == "auto" else modes[phase_mode]
        field.forward_propagation(back)

    with quiet:
        src_xyz = np.asarray(h["source"], dtype=np.float64).reshape(3, 1)
        src = WaveField3D(1, wavelength, 1, 1, device=device)                # CPU0402:247-256
        src.setdata(src_xyz)
        src.set_ds(np.ones(1))
        src.u = src.u + 1.0
        back, back_xyz = src, src_xyz
        dims = [(cond.get("ray_num_H1", 1), cond.get("ray_num_V1", 1)), (cond.get("ray_num_H2", 1), cond.get("ray_num_V2", 1))]
        for k, pts in enumerate(h["mirrors"]):
            name = f"M{k + 1}"
            if out_dir is not None:
                np.save(os.path.join(out_dir, f"points_{name}.npy"), pts)
            field = WaveField3D(pts.shape[1], wavelength, *dims[k % 2], device=device)
            field.setdata(pts)
            old = resumed(name)
            if old is not None:
                import torch
                field.u = torch.as_tensor(old).to(field.u.device) if device is not None else old
            else:
                propagate(field, back, pts, back_xyz)
            field.set_ds(pts[3, :])                                          # CPU0402:280,301,318,340
            save(name, field)
            back, back_xyz = field, pts
        grid = np.array(h["gridImage"], dtype=np.float64)
        for r in range(3):                                                   # CPU0402:330-334, row by row
            m = np.mean(grid[r, :])
            grid[r, :] = (grid[r, :] - m) * image_scale + m
        if out_dir is not None:
            np.save(os.path.join(out_dir, "points_gridImage.npy"), grid)
        img = WaveField3D(grid.shape[1], wavelength, cond.get("pix_y", 1), cond.get("pix_z", 1), device=device)
        img.setdata(grid)
        propagate(img, back, grid, back_xyz)
        save("Image", img)
        if h["gridDefocus"] is not None:
            g2 = np.array(h["gridDefocus"], dtype=np.float64)
            for r in range(3):                                               # CPU0402:354-358 (scale 1, not an identity in FP)
                m = np.mean(g2[r, :])
                g2[r, :] = (g2[r, :] - m) + m
            if out_dir is not None:
                np.save(os.path.join(out_dir, "points_gridImage2.npy"), g2)
            img2 = WaveField3D(g2.shape[1], wavelength, cond.get("pix_y", 1), cond.get("pix_z", 1), device=device)
            img2.setdata(g2)
            propagate(img2, back, g2, back_xyz)
            save("Image2", img2)
    return results
