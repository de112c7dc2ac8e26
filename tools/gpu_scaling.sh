#!/bin/bash
# Scaling session on an N-GPU box: the bench line at every N in $NS (weak scaling, one 512x512 block per rank).
TAG=${TAG:-r01f}
mkdir -p gpurun_out
for N in ${NS:-1 2 4 8}; do
  if [ "$N" = 1 ]; then
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n${N}_$TAG.json 2> gpurun_out/scale_n${N}_$TAG.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) bench.py --gpus $N --steps 3 --warmup 3 \
      > gpurun_out/scale_n${N}_$TAG.json 2> gpurun_out/scale_n${N}_$TAG.err
  fi
  echo "N=$N exit $?"; grep -o '"value": [0-9.e+]*' gpurun_out/scale_n${N}_$TAG.json | head -1
done
