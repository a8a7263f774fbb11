#!/bin/bash
# Round 2: padding-free layout -- the GPU suite and the bench (ragged variant is the line it should move).
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2f.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/pytest_r2f.log
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --json-out gpurun_out/bench_r2f.json > gpurun_out/bench_r2f.log 2>&1; echo "bench rc=$?"; tail -n 3 gpurun_out/bench_r2f.log | cut -c1-3000
