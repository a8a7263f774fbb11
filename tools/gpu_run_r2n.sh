#!/bin/bash
# Round 2: four-warpgroup attention kernel (attention5.cuh): correctness, A/B against the shipping kernel, timeline.
# The first command is a short canary: a kernel that hangs must not hold the box for the whole call.
mkdir -p gpurun_out
timeout -s KILL 150 python tools/att_bench.py 64 > gpurun_out/att_bench_canary.log 2>&1; rc=$?; echo "canary rc=$rc"; tail -n 6 gpurun_out/att_bench_canary.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout -s KILL 400 python tools/att_bench.py 0,5,64,65,69,73 > gpurun_out/att_bench_split4.log 2>&1; echo "att bench rc=$?"; tail -n 24 gpurun_out/att_bench_split4.log
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention and not causal" > gpurun_out/pytest_r2n.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_r2n.log | cut -c1-220
B2E_ATT3=69 timeout -s KILL 200 python tools/att3_timeline.py > gpurun_out/att3_timeline_v69.log 2>&1; tail -n 3 gpurun_out/att3_timeline_v69.log | cut -c1-700
