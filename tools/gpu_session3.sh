#!/bin/bash
TAG=${TAG:-r01e}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
rm -f /tmp/variant_ref_*
for v in 2 1 5; do AKB_FRESNEL_VARIANT=$v python tools/variant_bench.py 256 2; done 2>&1 | tee gpurun_out/variants_$TAG.log
python tools/variant_bench.py 256 2 2>&1 | tee -a gpurun_out/variants_$TAG.log
