#!/bin/bash
# Round 2: attention5 with one copy of each chunk body (65) and with the epilogue warpgroup (193).
mkdir -p gpurun_out
timeout -s KILL 150 python tools/att_bench.py 193 > gpurun_out/att_bench_canary2.log 2>&1; rc=$?; echo "canary rc=$rc"; tail -n 5 gpurun_out/att_bench_canary2.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout -s KILL 300 python tools/att_bench.py 5,65,193,197 > gpurun_out/att_bench_epi.log 2>&1; echo "att bench rc=$?"; grep "B=" gpurun_out/att_bench_epi.log
for v in 65 193; do timeout -s KILL 200 python tools/att_clock_probe.py $v 2>&1 | grep variant | cut -c1-200; done | tee gpurun_out/att_clock_probe_epi.log
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention and not causal and (shipping or epilogue)" > gpurun_out/pytest_r2x.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_r2x.log | cut -c1-200
B2E_ATT3=193 timeout -s KILL 200 python tools/att3_timeline.py > gpurun_out/att3_timeline_v193.log 2>&1; tail -n 2 gpurun_out/att3_timeline_v193.log | cut -c1-400
