#!/bin/bash
TAG=${TAG:-r01g}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_verbose_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_verbose_$TAG.log
tail -3 gpurun_out/pytest_gpu_verbose_$TAG.log
rm -f /tmp/variant_ref_*
AKB_FRESNEL_VARIANT=2 python tools/variant_bench.py 512 0 2>&1 | tee gpurun_out/variants_$TAG.log
python tools/variant_bench.py 512 0 2>&1 | tee -a gpurun_out/variants_$TAG.log
python tools/variant_bench.py 256 0 2>&1 | tee -a gpurun_out/variants_$TAG.log
python tools/variant_bench.py 2048 0 2>&1 | tee -a gpurun_out/variants_$TAG.log
python tools/run_configs.py --configs c1,m2m 2>/dev/null | tee gpurun_out/configs_$TAG.json | cut -c1-300
