#!/bin/bash
# Round 2: vectorised rotary kernel -- parity suite and the C5 / C3 steps.
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_r2k.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/pytest_r2k.log
timeout -s KILL 300 python tools/bench_esm2.py 64 2>/dev/null | tail -n 1 | tee gpurun_out/c5_r2k.json
timeout -s KILL 600 python tools/bench_mistral.py 2>/dev/null | tail -n 1 | tee gpurun_out/c3_r2k.json
