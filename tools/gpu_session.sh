#!/bin/bash
# One GPU-box session (1 GPU): smoke, parity tests, the bench line (both arms), the ncu launch list of the bench command,
# DRAM traffic of the bench's own pair-kernel launch (C4, 7 s: metrics-only pass), and a full capture of the same kernel on
# the C3-sized launch (0.42 s) plus the chain kernel.
TAG=${TAG:-r02}
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_$TAG.log
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
tail -3 gpurun_out/pytest_gpu_$TAG.log
python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "bench ref exit $?"; cut -c1-160 gpurun_out/bench_ref_$TAG.json
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; cut -c1-200 gpurun_out/bench_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1; echo "ncu launches exit $?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fresnel_pairs -s 1 -c 1 --csv \
    --log-file gpurun_out/traffic_c4_$TAG.csv python tools/profile_fresnel.py 2048 0 c4 > gpurun_out/ncu_traffic_$TAG.log 2>&1; echo "ncu traffic exit $?"
ncu --set full --clock-control none --import-source on -k regex:fresnel_pairs -s 1 -c 1 -f -o gpurun_out/${TAG}_pairs_c3 \
    python tools/profile_fresnel.py 512 0 c3 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full exit $?"
ncu --set full --clock-control none --import-source on -k regex:trace_chain_kernel -c 1 -f -o gpurun_out/${TAG}_chain \
    python tools/ray_bench.py > gpurun_out/ncu_chain_$TAG.log 2>&1; echo "ncu chain exit $?"
