#!/bin/bash
# Round 2: the N = 2 bench exactly as the driver launches it (NCCL), both arms, plus the 2-rank GPU tests.
mkdir -p gpurun_out
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_r02.json 2> gpurun_out/bench_2gpu_r02.err; echo "bench N=2 rc=$?"
tail -n 1 gpurun_out/bench_2gpu_r02.json | cut -c1-3500
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_2gpu_ref_r02.json 2> gpurun_out/bench_2gpu_ref_r02.err; echo "reference arm N=2 rc=$?"
tail -n 1 gpurun_out/bench_2gpu_ref_r02.json | cut -c1-1200
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 3
B2E_ATT3=5 timeout -s KILL 300 python tools/att3_timeline.py > gpurun_out/att3_timeline_v5_split.log 2>&1; tail -n 4 gpurun_out/att3_timeline_v5_split.log | cut -c1-900
