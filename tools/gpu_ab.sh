#!/bin/bash
# One GPU-box session: pair-kernel variant A/B on the C3 stage (needs an A/B build: AKB_AB_VARIANTS=1 python -m akbraytracing_b200.build --force).
# The first variant listed writes the reference field the others are compared with.
mkdir -p gpurun_out
rm -f /tmp/variant_ref_*
for m in ${MODES:-0}; do
  for v in ${VARIANTS:-4 0 6 7}; do
    AKB_FRESNEL_VARIANT=$v python tools/variant_bench.py ${GRID:-512} $m
    AKB_FRESNEL_VARIANT=$v python tools/variant_bench.py ${GRID:-512} $m general
  done
done 2>&1 | tee gpurun_out/variants.log
