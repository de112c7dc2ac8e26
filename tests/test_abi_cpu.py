"""CPU-only checks of the boundary: the C-ABI library loads, exports every symbol the header
declares, validates arguments and fails loudly without a GPU; host-side sharding logic."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from akbraytracing_b200 import build, _lib
    build.ensure_built()
    return _lib.load()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "akb_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(akb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    from akbraytracing_b200 import _lib
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/akb_b200.h but not exported"
    # and the ctypes table binds exactly the header
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_string(lib):
    assert lib.akb_version() >= 200
    assert isinstance(lib.akb_last_error(), bytes)


def test_shard_range_is_array_split(lib):
    from akbraytracing_b200 import _lib
    for total in (0, 1, 7, 4096, 4097, 2048 * 2048):
        for parts in (1, 2, 3, 4, 8):
            ref = np.array_split(np.arange(total), parts)  # the reference's cp.array_split (GPU0402:77-79)
            start = 0
            for r in range(parts):
                b, c = _lib.shard_range(total, parts, r)
                assert (b, c) == (start, len(ref[r]))
                start += c
    with pytest.raises(RuntimeError):
        _lib.shard_range(10, 0, 0)


def test_argument_validation_needs_no_gpu(lib):
    from akbraytracing_b200 import _lib
    rc = lib.akb_fresnel_sum(None, None, None, -1, None, None, None, None, None, 0, 1.0, None, 0, None)
    assert rc == -1 and b"non-negative" in lib.akb_last_error()
    rc = lib.akb_trace_chain(None, None, 0, None, None, None, 4, None, None, None, None, None, None, None, 0, None, None)
    assert rc == -1
    rc = lib.akb_fresnel_sum_sharded(None, 2, 2, None, None, None, 4, None, None, None, None, None, 1, 1.0, None, 0, 0, None)
    assert rc == -1 and b"rank" in lib.akb_last_error()
    rc = lib.akb_fresnel_sum_sharded(None, 0, 2, None, None, None, 4, None, None, None, None, None, 1, 1.0, None, 0, 0, None)
    assert rc == -1 and b"ncclComm_t" in lib.akb_last_error()
    rc = lib.akb_wavefront_opl(None, None, None, 0, 4, None, None, None, 0.0, None, None, None, None, None, None, None, None)
    assert rc == -1
    with pytest.raises(RuntimeError):
        _lib.check(rc, "akb_trace_chain")


def test_no_cpu_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import akbraytracing_b200 as akb
    x = np.zeros(4)
    with pytest.raises(RuntimeError):
        akb.forward_propagation_numpy_batch(x, x, x, x + 1, x, x, np.ones(4, complex), 1e6, np.ones(4))
    with pytest.raises(RuntimeError):
        akb._lib.device_count()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "akbraytracing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
                assert "liborc" not in text


def test_ell_host_algebra_matches_reference_coefficients(golden):
    """ell.__init__/coeffs are O(1) host algebra (ER3D:207-240): same 10 doubles as the reference."""
    import akbraytracing_b200 as akb
    g = golden("ray_er3d_ref")
    m = akb.ell(np.float64(146.), np.float64(0.086), np.float64(0.214) / 20, np.float64(0.060))
    m.coeffs('y')
    assert np.array_equal(np.asarray(m.coeffs, dtype=np.float64), g["single/coeffs"])
    assert float(m.dist_s_f) == float(g["single/plane_position"])
