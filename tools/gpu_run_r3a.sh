#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "profiling_instantiations or gemm_epilogues" > gpurun_out/pytest_r3a.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/pytest_r3a.log | cut -c1-220
timeout -s KILL 200 python tools/gemm_timeline.py > gpurun_out/gemm_timeline_r3a.log 2>&1; echo "timeline rc=$?"; tail -n 3 gpurun_out/gemm_timeline_r3a.log | cut -c1-300
