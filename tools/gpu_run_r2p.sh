#!/bin/bash
# Round 2: division-free item walker in the attention kernels: tests + timing with the clock.
mkdir -p gpurun_out
timeout -s KILL 150 python tools/att_bench.py 5 > gpurun_out/att_bench_walk.log 2>&1; rc=$?; echo "canary rc=$rc"; tail -n 4 gpurun_out/att_bench_walk.log
if [ $rc -ne 0 ]; then exit 1; fi
for v in 5 65; do timeout -s KILL 200 python tools/att_clock_probe.py $v 2>&1 | grep variant | tee -a gpurun_out/att_clock_probe_walk.log; done
timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_r2p.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/pytest_r2p.log | cut -c1-220
