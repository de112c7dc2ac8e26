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

__all__ = ["parse_conditions", "load_handoff", "write_handoff", "run_stage_chain", "focal_grids"]

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


def write_handoff(folder, source_point, mirror_clouds, grid_shape, det, det_defocus=None, option_AKB=None,
                  option_HighNA=True, focus_shape=None, defocus=0.0, ysize=1e-6, zsize=1e-6):
    """Write the folder saveWaveData writes (BIG:13515-13654) from traced clouds.

    mirror_clouds: list of (3, nV*nH) hit-point arrays (NumPy or torch CUDA) in file order M1..;
    grid_shape: (nV, nH) of the ray grid; det / det_defocus: traced detector points (3, N).
    dS comes from the device kernel (akb_calc_ds)."""
    os.makedirs(folder, exist_ok=True)
    nV, nH = grid_shape
    fV, fH = focus_shape if focus_shape is not None else grid_shape
    if option_AKB is None:
        option_AKB = len(mirror_clouds) == 4

    def host(a):
        return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)

    np.save(os.path.join(folder, "points_source.npy"), np.asarray(host(source_point), dtype=np.float64).reshape(3))
    for k, cloud in enumerate(mirror_clouds):
        dS = host(handoff.calc_dS(cloud, nV, nH))
        np.save(os.path.join(folder, f"points_M{k + 1}.npy"), np.vstack([host(cloud)[:3], dS.ravel()]))
    grid, y_grid, z_grid = focal_grids(host(det), fH, fV, ysize, zsize)
    np.save(os.path.join(folder, "points_gridImage.npy"), grid)
    if det_defocus is not None:
        ys = 2e-7 + abs(defocus) * (0.082 if option_HighNA else 0.01) * 2  # BIG:13594-13599
        grid2, _, _ = focal_grids(host(det_defocus), fH, fV, ys, ys)
        np.save(os.path.join(folder, "points_gridDefocus.npy"), grid2)
    with open(os.path.join(folder, "calculation_conditions.txt"), "w") as fh:  # BIG:13630-13654 (keys the reader uses)
        fh.write("Conditions\n====================\n")
        fh.write(f"grid pitch_y: {y_grid[1] - y_grid[0]}\n")
        fh.write(f"grid pitch_z: {z_grid[1] - z_grid[0]}\n")
        fh.write(f"grid pix_y: {fH}\n")
        fh.write(f"grid pix_z: {fV}\n")
        fh.write(f"grid pix_H1: {nH}\n")
        fh.write(f"grid pix_V1: {nV}\n")
        fh.write(f"grid pix_H2: {nH}\n")
        fh.write(f"grid pix_V2: {nV}\n")
        fh.write(f"option_AKB: {bool(option_AKB)}\n")
        fh.write(f"option_HighNA: {bool(option_HighNA)}\n")
        fh.write(f"defocusForWave: {defocus}\n")
        fh.write("====================\n")


def run_stage_chain(folder, out_dir=None, device="cuda", image_scale=2.0, resume_dir=None, verbose=False):
    """The stage chain of CPU0402:247-375 with device-resident fields.

    Returns dict(stage name -> complex128 NumPy field).  ``resume_dir``: like the reference, a
    ``complex_data_Mk.npz`` found there is loaded instead of computing that stage (CPU0402:261-265).
    ``image_scale``: the reference stretches the focal grid by 2 about its mean (CPU0402:330-334)."""
    import contextlib
    import io
    h = load_handoff(folder)
    cond = h["conditions"]
    wavelength = WAVELENGTH_HIGH_NA if cond.get("option_HighNA") else WAVELENGTH_LOW_NA  # CPU0402:242-245
    quiet = contextlib.nullcontext() if verbose else contextlib.redirect_stdout(io.StringIO())
    results = {}

    def save(name, field):
        arr = field.u.cpu().numpy() if hasattr(field.u, "cpu") else np.asarray(field.u)
        results[name] = arr
        if out_dir is not None:
            os.makedirs(out_dir, exist_ok=True)
            np.savez_compressed(os.path.join(out_dir, f"complex_data_{name}.npz"), data=arr)

    def resumed(name):
        if resume_dir is None:
            return None
        p = os.path.join(resume_dir, f"complex_data_{name}.npz")
        if os.path.exists(p):
            with np.load(p) as z:
                return z["data"]
        return None

    with quiet:
        src = WaveField3D(1, wavelength, 1, 1, device=device)               # CPU0402:247-256
        src.setdata(np.asarray(h["source"], dtype=np.float64).reshape(3, 1))
        src.set_ds(np.ones(1))
        src.u = src.u + 1.0
        back = src
        dims = [(cond.get("ray_num_H1", 1), cond.get("ray_num_V1", 1)), (cond.get("ray_num_H2", 1), cond.get("ray_num_V2", 1))]
        for k, pts in enumerate(h["mirrors"]):
            name = f"M{k + 1}"
            field = WaveField3D(pts.shape[1], wavelength, *dims[k % 2], device=device)
            field.setdata(pts)
            old = resumed(name)
            if old is not None:
                import torch
                field.u = torch.as_tensor(old).to(field.u.device) if device is not None else old
            else:
                field.forward_propagation(back)
            field.set_ds(pts[3, :])                                          # CPU0402:280,301,318,340
            save(name, field)
            back = field
        grid = np.array(h["gridImage"], dtype=np.float64)
        mean = grid.mean(axis=1, keepdims=True)
        grid = (grid - mean) * image_scale + mean                            # CPU0402:330-334
        img = WaveField3D(grid.shape[1], wavelength, cond.get("pix_y", 1), cond.get("pix_z", 1), device=device)
        img.setdata(grid)
        img.forward_propagation(back)
        save("Image", img)
        if h["gridDefocus"] is not None:
            g2 = np.array(h["gridDefocus"], dtype=np.float64)
            m2 = g2.mean(axis=1, keepdims=True)
            g2 = (g2 - m2) + m2                                              # CPU0402:354-358 (scale 1)
            img2 = WaveField3D(g2.shape[1], wavelength, cond.get("pix_y", 1), cond.get("pix_z", 1), device=device)
            img2.setdata(g2)
            img2.forward_propagation(back)
            save("Image2", img2)
    return results
