#!/bin/bash
TAG=${TAG:-r01d}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
rm -f /tmp/variant_ref_*
for v in 2 0 4; do AKB_FRESNEL_VARIANT=$v python tools/variant_bench.py 512 0; done 2>&1 | tee gpurun_out/variants_$TAG.log
rm -f /tmp/variant_ref_*
for v in 2 0; do AKB_FRESNEL_VARIANT=$v python tools/variant_bench.py 256 1; done 2>&1 | tee -a gpurun_out/variants_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; cut -c1-250 gpurun_out/bench_$TAG.json
