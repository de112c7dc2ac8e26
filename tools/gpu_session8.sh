#!/bin/bash
# 8-GPU session: the bench line at N=8 and configs C4/C5 sharded over 8 ranks.
TAG=${TAG:-r01c}
N=${N:-8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 \
    > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err; echo "bench exit $?"; cat gpurun_out/bench_n${N}_$TAG.json | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/run_configs.py --configs c4,c5 \
    > gpurun_out/configs_n${N}_$TAG.json 2> gpurun_out/configs_n${N}_$TAG.err; echo "configs exit $?"; cat gpurun_out/configs_n${N}_$TAG.json | cut -c1-500
