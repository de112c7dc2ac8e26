"""bench.py contract checks.  CPU: the reference arm (the one leg of the benchmark that needs no GPU): it must
print exactly one JSON line on stdout carrying the keys the driver reads, time the CPU restatement with
every host core even when OMP_NUM_THREADS=1 is exported (torch.distributed.run does that), and stay silent
on ranks other than 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *args):
    env = dict(os.environ, AKB_BENCH_REF_STEP_S="0.3", **extra_env)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                        *args], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run({"OMP_NUM_THREADS": "1"})
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "fresnel_terms_per_s" and d["unit"] == "terms/s"
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert d["steps"] == 1 and d["warmup"] == 3 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("C4") and "model" not in d["config"] and d["scaling"] == "strong"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["unit"] == "terms/s" and cb["sample"]
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    assert cb["cores"] == ncores  # not the single thread OMP_NUM_THREADS=1 would give
    assert d["e2e"] == {"value": d["value"], "unit": "terms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert out.strip() == ""


def test_reference_arm_at_n2_times_the_c4_stage():
    """Every N is BASELINE config 4 (AKB trace, fixed 2048x2048 detector, strong scaling): the reference arm samples the
    same stage, on rank 0 only."""
    out = _run({"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"}, "--gpus", "2")
    d = json.loads([l for l in out.splitlines() if l.strip()][0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["scaling"] == "strong"
    assert d["config"]["workload"].startswith("C4") and d["config"]["detector_points"] == 2048 * 2048
    assert d["config"]["mirrors"] == 4 and d["config"]["terms_per_step"] == 1e6 * 2048 * 2048
    assert "C4 2048x2048" in d["cpu_baseline"]["sample"]


import pytest  # noqa: E402


@pytest.mark.gpu
def test_our_arm_prints_one_json_line_with_the_contract_keys():
    """The GPU arm on one device (1 timed step, CPU baseline skipped to keep the test short)."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--no-cpu-baseline"],
                       capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert key in d, key
    assert d["metric"] == "fresnel_terms_per_s" and d["n_gpus"] == 1 and d["scaling"] == "strong" and d["dtype"] == "f64"
    assert d["config"]["workload"].startswith("C4") and d["config"]["terms_per_step"] == 1e6 * 2048 * 2048
    p = d["parity"]
    assert p["rel_l2"] < 1e-11 and p["uneven_rel_l2"] < 1e-11 and p["peak_same"] is True and p["ranks_equal"] is None
    assert d["c3"]["terms_per_s"] > 1e11 and d["c3"]["rel_l2_vs_oracle"] < 1e-11
    assert d["value"] > 1e11 and d["gpu_launches"] > 0  # the CUDA path ran (a CPU fallback could not reach 1e11 terms/s)
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert 0.0 < r["frac"] < 1.0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert 0.9 < r["kernel_share_of_step"] <= 1.0
    e = d["e2e"]
    assert e["value"] > 1e11 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["max_abs_diff_vs_device_path"] == 0.0  # host-buffer call and device call: the same bits
    assert e["steps"] >= 1 and e["pageable"]["value"] > 1e11
    assert d["gpu_baseline"]["rel_l2_vs_fused_kernel"] < 1e-9
    assert d["roofline_ray"]["bound"] == "hbm" and 0.5 < d["roofline_ray"]["frac"] < 1.0
    assert "strided" in d["roofline_ray"]["kernel"]
    # the instruction counts in the roofline come from the loaded library's SASS
    assert r["fp64_instr_per_term"] == r["sass"]["planar_row"]["fp64_instr_per_pair"] == 25.5
    m = d["roofline_m2m"]
    assert m["bound"] == "fp64" and m["fp64_instr_per_term"] == 29.0 and 0.0 < m["frac"] < 1.0
    for mode in ("faithful", "exact", "referenced"):
        assert m["modes"][mode]["terms_per_s"] > 1e11
    assert m["modes"]["faithful"]["rel_l2_vs_oracle"] < 1e-11 and m["modes"]["exact"]["rel_l2_vs_oracle"] < 1e-6
    c = d["roofline_chain"]
    assert c["bound"] == "hbm" and c["cases"]["K2_segments"]["misses"] == 0 and c["cases"]["K4_opl"]["misses"] == 0
    assert 0.3 < c["frac"] < 1.0 and d["small_call_us"]["median"] > 0
