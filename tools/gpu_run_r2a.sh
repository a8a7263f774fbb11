#!/bin/bash
# Round 2, first GPU call: the full -m gpu suite (incl. the new config-size parity and worker tests), the
# per-layer drift report, and both bench arms.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
timeout -s KILL 1500 python -m pytest tests -m gpu -q -x --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 30 gpurun_out/pytest_gpu.log
timeout -s KILL 900 python tools/drift_report.py gpurun_out/drift_report.md > gpurun_out/drift.log 2>&1; echo "drift rc=$?"; grep -E "pooled|oracle" gpurun_out/drift.log
timeout -s KILL 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
timeout -s KILL 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
