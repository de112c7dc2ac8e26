#!/bin/bash
# GPU-box session: parity tests, swizzle A/B, compute-sanitizer passes, configs C1..C4 on one GPU.
TAG=${TAG:-r01c}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
rm -f /tmp/variant_ref_*
for v in 0 5 0 5; do AKB_FRESNEL_VARIANT=$v python tools/variant_bench.py 512 0; done 2>&1 | tee gpurun_out/variants_$TAG.log
SEL="ragged or weight_phases or chain_bit_exact or single_mirror or calc_dS or batched or coincident or empty"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 86 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$SEL" \
    > gpurun_out/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck exit $?" | tee -a gpurun_out/sanitizer_memcheck_$TAG.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 86 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "ragged or batched" \
    > gpurun_out/sanitizer_racecheck_$TAG.log 2>&1; echo "racecheck exit $?" | tee -a gpurun_out/sanitizer_racecheck_$TAG.log
tail -3 gpurun_out/sanitizer_memcheck_$TAG.log gpurun_out/sanitizer_racecheck_$TAG.log
python tools/run_configs.py --configs c1,c2,c3,c4 > gpurun_out/configs_n1_$TAG.json 2> gpurun_out/configs_n1_$TAG.err; echo "configs exit $?"; cat gpurun_out/configs_n1_$TAG.json | cut -c1-400
