#!/bin/bash
# One GPU-box session: pair-kernel variant A/B on the C3 stage (variant 2 = round-1 formulation first: it is the reference field)
mkdir -p gpurun_out
rm -f /tmp/variant_ref_*
for v in ${VARIANTS:-2 6 7 8 9 10}; do AKB_FRESNEL_VARIANT=$v python tools/variant_bench.py ${GRID:-512} 0; done 2>&1 | tee gpurun_out/variants.log
