#!/bin/bash
# Round 2: packed token layout for the Mistral path + four-warpgroup attention as the default: canary, the whole
# GPU suite, C3 ragged vs full, the default bench.
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "packed or mistral" > gpurun_out/pytest_r2r_canary.log 2>&1; rc=$?; echo "canary rc=$rc"; tail -n 6 gpurun_out/pytest_r2r_canary.log | cut -c1-220
if [ $rc -ne 0 ]; then exit 1; fi
timeout -s KILL 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2r.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/pytest_r2r.log | cut -c1-220
timeout -s KILL 600 python tools/bench_mistral.py 2>/dev/null | tail -n 1 | cut -c1-330 | tee gpurun_out/c3_r2r.json
timeout -s KILL 600 python tools/bench_mistral.py 16 4096 32 ragged 2>/dev/null | tail -n 1 | cut -c1-400 | tee gpurun_out/c3_ragged_r2r.json
B2E_PACKED=0 timeout -s KILL 600 python tools/bench_mistral.py 16 4096 32 ragged 2>/dev/null | tail -n 1 | cut -c1-400 | tee gpurun_out/c3_ragged_padded_r2r.json
