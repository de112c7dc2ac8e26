#!/usr/bin/env python
"""Write profiles/r02_configs.md from the committed bench / config JSON files of round 2."""
import json
import os

P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
ld = lambda f: json.load(open(os.path.join(P, f)))  # noqa: E731
b, n2, n4, n8 = ld("r02r_bench_n1.json"), ld("r02o_bench_n2.json"), ld("r02o_bench_n4.json"), ld("r02h_bench_n8.json")
c5 = ld("r02h_config_c5_n8.json")
c5 = c5[0] if isinstance(c5, list) else c5
fc = ld("r02p_full_chain_n1.json")
rows = []
rows.append(("C1: 1e4 rays → 64×64 grid (host-buffer call)",
             f"{b['small_call_us']['median']:.0f} µs per call = {b['small_call_us']['terms_per_s']:.3g} terms/s (round 1: 293 µs)",
             "full oracle parity ≤ 1e-12 (`tests/test_gpu_parity.py::test_c1_config_against_oracle`)", "`r02r_bench_n1.json` → `small_call_us`"))
r = b["roofline_ray"]
rows.append(("C2: single elliptical mirror, 1.0005e7 rays, intersect+normal+reflect",
             f"{r['kernel_ms']:.4f} ms = {r['rays_per_s']:.3g} rays/s = {r['achieved']:.0f} GB/s ({r['frac']:.2f} of the measured HBM peak); "
             f"without the normal {r['without_normal']['achieved']:.0f} GB/s",
             "strided subset bit-identical to the oracle, points on the quadric to 1e-13 (`test_c2_full_size_single_mirror`)",
             "`r02r_bench_n1.json` → `roofline_ray`"))
c3 = b["c3"]
rows.append(("C3: KB two-mirror trace 1e6 rays → 512×512 grid, 1 GPU",
             f"{c3['ms_per_step']:.1f} ms per stage (trace + dS + field + pair sum) = {c3['terms_per_s']:.4g} terms/s (round 1: 5.82e11); REFERENCED mode 6.97e11",
             f"rel-L2 vs oracle {c3['rel_l2_vs_oracle']:.1e} on 64 points", "`r02r_bench_n1.json` → `c3`; `r02_variants_ab.md` §6"))
rows.append(("C4: AKB four-mirror trace 1e6 rays × 2048×2048 detector, 1 / 2 / 4 / 8 GPUs (the bench headline)",
             f"{b['value']:.4g} / {n2['value']:.4g} / {n4['value']:.4g} / {n8['value']:.4g} terms/s "
             f"({b['ms_per_step'] / 1e3:.2f} / {n2['ms_per_step'] / 1e3:.2f} / {n4['ms_per_step'] / 1e3:.2f} / {n8['ms_per_step'] / 1e3:.3f} s per step); "
             f"e2e from NumPy buffers {b['e2e']['value']:.4g} / {n2['e2e']['value']:.4g} / {n4['e2e']['value']:.4g} / {n8['e2e']['value']:.4g}; "
             f"REFERENCED mode at N=1: {b['phase_modes']['referenced']['terms_per_s']:.4g} "
             f"({b['phase_modes']['referenced']['rel_l2_vs_faithful']:.1e} from the faithful field)",
             f"oracle parity on 256 points per rank: {b['parity']['rel_l2']:.1e} / {n2['parity']['rel_l2']:.1e} / {n4['parity']['rel_l2']:.1e} / "
             f"{n8['parity']['rel_l2']:.1e}, same peak, ranks bit-equal, uneven shards {n8['parity']['uneven_rel_l2']:.1e}",
             "`r02r_bench_n1.json`, `r02o_bench_n2.json`, `r02o_bench_n4.json`, `r02h_bench_n8.json`, `r02_scaling.md`"))
rows.append(("C5: through-focus stack 32 planes × 1024×1024 + PSF per plane, 8 GPUs (product calls `fresnel_sum_planes` + `psf_stack`)",
             f"{c5['ms'] / 1e3:.2f} s = {c5['terms_per_s']:.4g} terms/s; 32 PSFs (pad 2, 4 per GPU) {c5['psf_ms_all_planes']:.0f} ms; pad 16: 43 ms and 16 GiB per plane",
             f"rel-L2 vs oracle {c5['rel_l2_vs_oracle_subset']:.1e} on {c5['subset']} points, same peak", "`r02h_config_c5_n8.json`, `r02n_psf_pad_factor.log`"))
m = b["roofline_m2m"]["modes"]
rows.append(("M2M: AKB mirror 1 (1e6 points) → 262 144 points of mirror 2 (irregular detector set, the general loop)",
             f"faithful {m['faithful']['terms_per_s']:.4g}, exact {m['exact']['terms_per_s']:.4g}, referenced {m['referenced']['terms_per_s']:.4g} terms/s",
             f"vs oracle on 64 points: {m['faithful']['rel_l2_vs_oracle']:.1e} / {m['exact']['rel_l2_vs_oracle']:.1e} / {m['referenced']['rel_l2_vs_oracle']:.1e}",
             "`r02r_bench_n1.json` → `roofline_m2m`"))
rows.append(("the whole Wavecalc workflow at 1e6 points per mirror (AKB: source→M1→M2→M3→M4→Image→Image2, 3.52e12 terms), 1 GPU",
             f"{fc['chain_s_faithful']:.2f} s with the reference's roundings = {fc['terms_per_s_faithful']:.4g} terms/s incl. file reads; "
             f"{fc['chain_s_auto']:.2f} s with `phase_mode='auto'`",
             f"auto vs faithful ≤ {max(fc['auto_vs_faithful_rel_l2'].values()):.1e}, same peak; last stage vs oracle {fc['image_stage_rel_l2_vs_oracle_32pts']:.1e}",
             "`r02p_full_chain_n1.json`"))
ch = b["roofline_chain"]["cases"]
rows.append(("chain kernel at 1e7 rays (C-ABI)",
             f"K=2 with segment lengths {ch['K2_segments']['kernel_ms']:.3f} ms = {ch['K2_segments']['achieved']:.0f} GB/s ({ch['K2_segments']['frac']:.2f} of HBM, "
             f"FP64 pipe ≈ {ch['K2_segments']['fp64_pipe_frac_est']:.2f}); with the summed path {ch['K2_opl']['kernel_ms']:.3f} ms ({ch['K2_opl']['frac']:.2f}); "
             f"K=4 {ch['K4_segments']['kernel_ms']:.3f} ms ({ch['K4_segments']['frac']:.2f} of HBM, FP64 pipe ≈ {ch['K4_segments']['fp64_pipe_frac_est']:.2f})",
             "bit-identical to every call of the reference drivers' kept pass (`test_chain_bit_exact_vs_driver`), `opl` = `totalDist`",
             "`r02r_bench_n1.json` → `roofline_chain`"))
with open(os.path.join(P, "r02_configs.md"), "w") as fh:
    fh.write("# r02: BASELINE.json configs C1–C5 (+ the mirror-to-mirror stage, the whole workflow and the chain kernel) on B200, final state of round 2\n\n")
    fh.write("All numbers from the committed JSON files named in the last column (CUDA-event times, 1965 MHz, no throttle reasons); "
             "written by tools/configs_table.py.\n\n| config | time / throughput | parity | evidence |\n|---|---|---|---|\n")
    for row in rows:
        fh.write("| " + " | ".join(row) + " |\n")
print("wrote profiles/r02_configs.md")
