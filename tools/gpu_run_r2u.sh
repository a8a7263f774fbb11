#!/bin/bash
# Round 2: the N = 4 bench exactly as the driver launches it (weak scaling point between 2 and 8).
mkdir -p gpurun_out
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 \
  bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_4gpu_r02.json 2> gpurun_out/bench_4gpu_r02.err; echo "bench N=4 rc=$?"
tail -n 1 gpurun_out/bench_4gpu_r02.json | cut -c1-600
tail -n 5 gpurun_out/bench_4gpu_r02.err | cut -c1-300
