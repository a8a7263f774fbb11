#!/bin/bash
# Round 2, third GPU call: everything again with IEEE-half weights / activations (was bfloat16).
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | head -30
timeout -s KILL 900 python tools/drift_report.py gpurun_out/drift_report_f16.md > gpurun_out/drift3.log 2>&1; echo "drift rc=$?"; grep -E "pooled|oracle" gpurun_out/drift3.log
timeout -s KILL 600 python tools/stage_errors.py 1024 2 > gpurun_out/stage_errors_f16.log 2>&1; echo "stage rc=$?"; cat gpurun_out/stage_errors_f16.log
timeout -s KILL 600 python tools/att_bench.py 0,1,5 > gpurun_out/att_bench2.log 2>&1; echo "att rc=$?"; cat gpurun_out/att_bench2.log
for fl in 2 1 0; do B2E_ATT3=5 B2E_ATT3_FLAGS=$fl timeout -s KILL 120 python tools/att3_timeline.py 40 > gpurun_out/att3_timeline_v5_flags$fl.log 2>&1; echo "timeline flags=$fl rc=$?"; done
head -c 1500 gpurun_out/att3_timeline_v5_flags2.log
timeout -s KILL 900 python bench.py --steps 10 > gpurun_out/bench_f16.json 2> gpurun_out/bench_f16.err; echo "bench rc=$?"; cat gpurun_out/bench_f16.json; tail -n 3 gpurun_out/bench_f16.err
