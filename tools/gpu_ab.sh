#!/bin/bash
# One GPU-box session: GPU parity tests, pair-kernel variant A/B on the C3 stage, microbenchmarks.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
rm -f /tmp/variant_ref_*
for v in 2 0 7 6 4 5; do AKB_FRESNEL_VARIANT=$v python tools/variant_bench.py 512 0; done 2>&1 | tee gpurun_out/variants.log
for m in 1 2; do rm -f /tmp/variant_ref_*; for v in 2 0 1 7; do AKB_FRESNEL_VARIANT=$v python tools/variant_bench.py 256 $m; done; done 2>&1 | tee -a gpurun_out/variants.log
./tools/ubench/fp64_dmma 2>&1 | tee gpurun_out/fp64_dmma.log
