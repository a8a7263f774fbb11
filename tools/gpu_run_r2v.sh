#!/bin/bash
# Round 2: compute-sanitizer memcheck over the kernels that changed most this round (four-warpgroup attention,
# packed token layout for every family, tensor-core retrieval scan).
mkdir -p gpurun_out
timeout -s KILL 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --log-file gpurun_out/memcheck_r02.log \
  python -m pytest tests/test_gpu_kernels.py tests/test_gpu_parity.py -m gpu -q -x \
  -k "(attention_matches_reference and shipping) or (sliding_window and shipping) or packed_token_layout or tensor_core_scan or ubinary" \
  > gpurun_out/pytest_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -n 5 gpurun_out/pytest_memcheck.log | cut -c1-200
grep -c "Invalid\|Error" gpurun_out/memcheck_r02.log; tail -n 6 gpurun_out/memcheck_r02.log | cut -c1-200
