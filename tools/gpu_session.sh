#!/bin/bash
# One GPU-box session: parity tests, the bench line, the ncu launch list and full captures of the pair and ray kernels.
TAG=${TAG:-r01d}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; cut -c1-200 gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "bench ref exit $?"; cut -c1-200 gpurun_out/bench_ref_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:fresnel_pairs -s 3 -c 1 -f -o gpurun_out/${TAG}_bench_pairs \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full exit $?"
ncu --set full --clock-control none --import-source on -k regex:intersect_reflect_strided -s 2 -c 2 -f -o gpurun_out/${TAG}_ray \
    python tools/ray_bench.py > gpurun_out/ncu_ray_$TAG.log 2>&1; echo "ncu ray exit $?"
