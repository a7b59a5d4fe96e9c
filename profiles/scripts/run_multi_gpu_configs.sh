#!/bin/bash
# usage: run_multi.sh N
N=$1
cd /root/repo
for c in 2 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$c bench.py --gpus $N --config $c --no-cpu-baseline > gpurun_out/r2_c${c}_n${N}.log 2> gpurun_out/r2_c${c}_n${N}.err
  tail -1 gpurun_out/r2_c${c}_n${N}.log | cut -c1-250
done
