#!/bin/bash
# Round 2, final evidence on one GPU: the default bench, the ncu launch list of the bench command, one full layer at
# the bench batch under --set full, a source-level capture of the (four-warpgroup) attention kernel.
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 --json-out gpurun_out/bench_final_r02.json > gpurun_out/bench_final_r02.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/bench_final_r02.log | cut -c1-300
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final_r02.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu_final.log 2>&1; echo "ncu list rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k "regex:gemm|attention|attn_prep|layernorm" --launch-skip 7 -c 7 -f \
  -o gpurun_out/layer_b512_final python tools/prof_kernels.py 512 > gpurun_out/ncu_full_final.log 2>&1; echo "ncu full rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k "regex:attention5" --launch-skip 1 -c 1 -f \
  -o gpurun_out/att5_src_final python tools/prof_kernels.py 128 > gpurun_out/ncu_att5_final.log 2>&1; echo "ncu att5 rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
ls -la gpurun_out | grep final
