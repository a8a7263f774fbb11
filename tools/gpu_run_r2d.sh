#!/bin/bash
# Round 2, fourth GPU call: two builds (half / bfloat16 storage), ModernBERT, binary retrieval, storage A/B.
mkdir -p gpurun_out
timeout -s KILL 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | head -40
timeout -s KILL 900 python bench.py --steps 10 > gpurun_out/bench_two_builds.json 2> gpurun_out/bench_two_builds.err; echo "bench rc=$?"; cat gpurun_out/bench_two_builds.json; tail -n 3 gpurun_out/bench_two_builds.err
timeout -s KILL 900 python tools/drift_report.py gpurun_out/drift_report_shipping.md bert,esm > gpurun_out/drift4.log 2>&1; echo "drift rc=$?"; grep -E "pooled" gpurun_out/drift4.log
