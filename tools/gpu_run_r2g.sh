#!/bin/bash
# Round 2: fp32x2 softmax variants (micro-benchmark + kernel A/B), then the whole GPU suite.
mkdir -p gpurun_out
timeout -s KILL 120 tools/bin/exp_phase_bench > gpurun_out/exp_phase_bench_x2.log 2>&1; echo "exp bench rc=$?"; cat gpurun_out/exp_phase_bench_x2.log
timeout -s KILL 600 python tools/att_bench.py > gpurun_out/att_bench_x2.log 2>&1; echo "att bench rc=$?"; tail -n 30 gpurun_out/att_bench_x2.log
for v in 37 41; do B2E_ATT3=$v timeout -s KILL 300 python tools/att3_timeline.py > gpurun_out/att3_timeline_v$v.log 2>&1; tail -n 12 gpurun_out/att3_timeline_v$v.log; done
timeout -s KILL 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2g.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/pytest_r2g.log
