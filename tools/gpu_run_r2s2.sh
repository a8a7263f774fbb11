#!/bin/bash
# Round 2, final tree: ncu launch list of the bench command, one full layer under --set full, attention5 source capture.
mkdir -p gpurun_out
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final_r02.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_final.log 2>&1; echo "ncu list rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k "regex:gemm|attention|attn_prep|layernorm" --launch-skip 7 -c 7 -f \
  -o gpurun_out/layer_b512_final python tools/prof_kernels.py 512 > gpurun_out/ncu_full_final.log 2>&1; echo "ncu full rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k "regex:attention5" --launch-skip 1 -c 1 -f \
  -o gpurun_out/att5_src_final python tools/prof_kernels.py 128 > gpurun_out/ncu_att5_final.log 2>&1; echo "ncu att5 rc=$?"
