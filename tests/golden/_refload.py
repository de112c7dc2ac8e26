"""Load the reference's hot-path functions inside the BUILD container only.

This module is used by ``make_golden.py`` (and by nothing that runs on the GPU
box: ``/root/reference`` does not exist there).  It never copies reference
source; it imports / executes it in place:

* ``Wavecalc_raytrace_fromData_CPU0402`` and ``EllipseRaytrace3D`` are imported
  unmodified with the two packages they import but never use on the hot path
  (``h5py``: CPU0402:7, ``matplotlib``: ER3D:3) stubbed in ``sys.modules``.
* ``AKB_raytrace_20250312`` creates output directories and pulls plotting
  packages at import (BIG:102-114), so only its ``FunctionDef`` nodes are
  executed, into a namespace seeded with the module flags of BIG:65-100.
"""
from __future__ import annotations

import ast
import contextlib
import io
import os
import sys
import types

REF = os.environ.get("AKB_REFERENCE", "/root/reference")


def _stub(name: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__["__getattr__"] = lambda attr: _stub(name + "." + attr)
    sys.modules.setdefault(name, mod)
    return sys.modules[name]


def available() -> bool:
    return os.path.isdir(REF)


def load_cpu0402():
    _stub("h5py")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import Wavecalc_raytrace_fromData_CPU0402 as m  # noqa: N813
    return m


def load_er3d():
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import EllipseRaytrace3D as m  # noqa: N813
    return m


def load_psf():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import psf_fft as m
    return m


class _Anything:
    """Absorbs any plotting call made by the drivers (plt.figure(), ...)."""

    def __getattr__(self, _):
        return self

    def __call__(self, *a, **k):
        return self

    def __iter__(self):
        return iter(())


def load_big(wave_num: int = 65, option_AKB: bool = True, flags: dict | None = None) -> dict:
    """Namespace holding every function defined in AKB_raytrace_20250312.py.

    ``wave_num`` is the ray grid side used by option 'wave' (BIG:1890-1893).
    """
    import numpy as np
    from scipy.interpolate import interp1d, griddata

    path = os.path.join(REF, "AKB_raytrace_20250312.py")
    with open(path, "r", encoding="utf-8") as fh:
        tree = ast.parse(fh.read(), filename=path)

    ns: dict = {"__name__": "big_extracted", "np": np, "interp1d": interp1d, "griddata": griddata,
                "plt": _Anything(), "os": os, "sys": sys}
    for fn in ("abs", "sin", "cos", "tan", "arcsin", "arccos", "arctan", "sqrt"):
        ns[fn] = getattr(np, fn)
    ns["pi"] = np.pi
    ns.update(dict(
        optKBdesign=False, option_2mirror=True, option_rotate=True, option_avrgsplt=False,
        downsample_h1=0, downsample_v1=0, downsample_h2=0, downsample_v2=0,
        downsample_h_f=0, downsample_v_f=0,
        wave_num_H=wave_num, wave_num_V=wave_num, unit=wave_num,
        option_AKB=option_AKB, option_wolter_3_1=True, option_wolter_3_3_tandem=False,
        option_HighNA=True, option_energy="EUV", LowNAratio=1.0, defocusForWave=1e-3,
        option_mpmath=False, option_set=False, var_input=1, widesearch=False,
        KBdesign_7params=[np.float64(146.), np.float64(0.21), np.float64(0.16742), np.float64(0.180),
                          np.float64(0.030), np.float64(0.15525), np.float64(0.05)],
        directory_name="/tmp/akb_ref_out",
    ))
    if flags:
        ns.update(flags)

    defs = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)):
            defs.append(node)
        elif isinstance(node, ast.If) and isinstance(node.test, ast.Name) and node.test.id == "option_wolter_3_1":
            # BIG:1325 -- the Wolter III+I variant of plot_result_debug
            defs.extend(n for n in node.body if isinstance(n, ast.FunctionDef))
    mod = ast.Module(body=defs, type_ignores=[])
    exec(compile(mod, path, "exec"), ns)
    return ns


@contextlib.contextmanager
def quiet():
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield buf


AKB_ALIGNMENT_PRESET = [  # BIG:14586-14592 (values, not code)
    -5.73452570e-03, -2.87624337e-03, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0,
    1.05000000e-02, -3.59399021e-05, 0.0, 0.0, 0.0, 2.39536993e-06, 0.0, 0.0,
    0.0, 0.0, 0.0, 0.0, 1.05000000e-02, -3.59399021e-05, 0.0, 0.0, 0.0, 2.39536993e-06]
