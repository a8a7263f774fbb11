#!/bin/bash
# Round 2: attention5 without the timeline stamps in the production instantiation (321 = with stamps).
mkdir -p gpurun_out
timeout -s KILL 150 python tools/att_bench.py 65 > gpurun_out/att_bench_canary3.log 2>&1; rc=$?; echo "canary rc=$rc"; tail -n 4 gpurun_out/att_bench_canary3.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout -s KILL 300 python tools/att_bench.py 0,5,261,65,69,321 > gpurun_out/att_bench_nostamp2.log 2>&1; echo "att bench rc=$?"; grep "B=" gpurun_out/att_bench_nostamp2.log
timeout -s KILL 200 python tools/att_clock_probe.py 5 2>&1 | grep variant | cut -c1-200 | tee gpurun_out/att_clock_probe_nostamp_v5.log
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q -x -k "attention or packed or modernbert" > gpurun_out/pytest_r2z.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_r2z.log | cut -c1-200
B2E_ATT3=261 timeout -s KILL 200 python tools/att3_timeline.py > gpurun_out/att3_timeline_v261.log 2>&1; tail -n 2 gpurun_out/att3_timeline_v261.log | cut -c1-400
