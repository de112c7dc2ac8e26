#!/bin/bash
# Short 1-GPU session: smoke, parity tests, both bench arms with their wall time, the ncu launch list of the bench command.
# SKIP_TESTS=1 skips the first two; SKIP_NCU=1 the last.
TAG=${TAG:-r02x}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_$TAG.log
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
fi
SECONDS=0
python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "bench ref exit $?, wall $SECONDS s"; cut -c1-160 gpurun_out/bench_ref_$TAG.json
SECONDS=0
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?, wall $SECONDS s"; cut -c1-200 gpurun_out/bench_$TAG.json
if [ -z "$SKIP_NCU" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches exit $?"
fi
