#!/bin/bash
# Round 2: ncu evidence -- launch list of the bench command, one full layer at the bench batch with
# `--set full`, and a source-level capture of the attention kernel (stall reasons per instruction).
mkdir -p gpurun_out
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_r02.log 2>&1; echo "ncu list rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k "regex:gemm|attention3|attn_prep|layernorm" --launch-skip 7 -c 7 -f \
  -o gpurun_out/layer_b512_r02 python tools/prof_kernels.py 512 > gpurun_out/ncu_full_r02.log 2>&1; echo "ncu full rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k "regex:attention3" --launch-skip 1 -c 1 -f \
  -o gpurun_out/att3_src_r02 python tools/prof_kernels.py 128 > gpurun_out/ncu_att3_r02.log 2>&1; echo "ncu att3 rc=$?"
ls -la gpurun_out | tail -n 6
