#!/bin/bash
# One GPU-box session: parity tests, the bench line, the ncu launch list and one full capture of the pair kernel.
TAG=${TAG:-r01b}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; cat gpurun_out/bench_$TAG.json
rm -f /tmp/variant_ref_*
for m in 1 2; do rm -f /tmp/variant_ref_*; for v in 2 0 1 4; do AKB_FRESNEL_VARIANT=$v python tools/variant_bench.py 256 $m; done; done 2>&1 | tee gpurun_out/variants_modes_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:fresnel_pairs -s 3 -c 1 -f -o gpurun_out/${TAG}_bench_pairs \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full exit $?"
