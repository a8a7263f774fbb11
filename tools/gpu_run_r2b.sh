#!/bin/bash
# Round 2, second GPU call: per-operator errors of the Mistral block, the exp-phase micro-benchmark, the
# attention softmax variants, then the whole -m gpu suite (no -x) and the drift report with the stable
# outlier model.
mkdir -p gpurun_out
timeout -s KILL 600 python tools/stage_errors.py 1024 2 > gpurun_out/stage_errors.log 2>&1; echo "stage rc=$?"; cat gpurun_out/stage_errors.log
timeout -s KILL 300 tools/bin/exp_phase_bench > gpurun_out/exp_phase_bench.log 2>&1; echo "expbench rc=$?"; cat gpurun_out/exp_phase_bench.log
timeout -s KILL 600 python tools/att_bench.py > gpurun_out/att_bench.log 2>&1; echo "att rc=$?"; cat gpurun_out/att_bench.log
timeout -s KILL 1800 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 40 gpurun_out/pytest_gpu.log | cut -c1-300
timeout -s KILL 900 python tools/drift_report.py gpurun_out/drift_report_v2.md > gpurun_out/drift2.log 2>&1; echo "drift rc=$?"; grep -E "pooled|oracle" gpurun_out/drift2.log
