#!/bin/bash
# One gpurun call that produces every round-end artefact: GPU tests, smoke, both bench arms, the ncu
# launch list of the bench command, and an `ncu --set full` capture of one encoder layer at the bench
# batch.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout -s KILL 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k "regex:gemm|attention3|attn_prep|layernorm" --launch-skip 7 -c 7 -f \
  -o gpurun_out/layer_b512 python tools/prof_kernels.py 512 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out | tail -n 12
