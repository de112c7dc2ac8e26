"""CPU check of the instruction counts the roofline is quoted with (bench.py reads them from the loaded library with the
same code): the default FAITHFUL pair kernel executes 25.5 FP64 instructions per pair in its planar-row loop and 29 in
its general loop, and the built library really is sm_100a code with TMA bulk copies and mbarriers in it."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    from akbraytracing_b200 import build
    return build.ensure_built()


def test_default_pair_kernel_instruction_mix(lib_path):
    import sys
    sys.path.insert(0, ROOT)
    from tools import sass_cost
    c = sass_cost.default_kernel_loops(lib_path)
    row, gen = c["planar_row"], c["general"]
    assert row["fp64_instr_per_pair"] == 25.5 and gen["fp64_instr_per_pair"] == 29.0
    # 2*DFMA + DMUL + DADD per pair (the executed-flop figure of the roofline)
    assert row["exec_flop_per_pair"] == 2 * row["dfma_per_pair"] + row["dmul_per_pair"] + row["dadd_per_pair"]
    assert 38.0 <= row["exec_flop_per_pair"] <= 41.0 and 42.0 <= gen["exec_flop_per_pair"] <= 45.0
    assert row["three_read_per_pair"] <= 5.5 and gen["three_read_per_pair"] <= 5.5


def test_referenced_pair_kernel_instruction_mix(lib_path):
    """AKB_PHASE_REFERENCED: 20.75 FP64 instructions per pair on planar-row blocks (row expansion: one square root per
    four pairs), 26 in the general loop, and no square-root / reciprocal built-in call in either loop."""
    import sys
    sys.path.insert(0, ROOT)
    from tools import sass_cost
    c = sass_cost.kernel_loops(lib_path, sass_cost.REFERENCED_KERNEL)
    assert c["planar_row"]["fp64_instr_per_pair"] == 20.75 and c["general"]["fp64_instr_per_pair"] == 26.0


def test_library_is_sm100a_with_tma_and_mbarrier(lib_path):
    elf = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    sass = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    assert sass.count("UBLKCP") >= 10      # cp.async.bulk: the 1-D TMA bulk copy of source tiles
    assert sass.count("SYNCS") >= 20       # mbarrier arrive/try_wait
