#!/bin/bash
# Round 2: tensor-core retrieval scan -- tests, then both scans side by side on 2 M x 768.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "topk" > gpurun_out/pytest_r2l.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/pytest_r2l.log | cut -c1-220
true
