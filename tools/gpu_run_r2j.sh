#!/bin/bash
# Round 2: parked waits for the attention kernel's single-thread roles (flag bit 2) A/B, and ncu launch lists of
# the C5 (ESM2-650M) and C3 (Mistral-7B) steps for their kernel shares.
mkdir -p gpurun_out
timeout -s KILL 600 python tools/att_bench.py 5 2,6,2,6 > gpurun_out/att_bench_park.log 2>&1; echo "att bench rc=$?"; grep "B=" gpurun_out/att_bench_park.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c5_r02.csv \
  python tools/bench_esm2.py 64 > gpurun_out/c5_under_ncu.log 2>&1; echo "ncu c5 rc=$?"
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_c3_r02.csv \
  python tools/bench_mistral.py > gpurun_out/c3_under_ncu.log 2>&1; echo "ncu c3 rc=$?"
ls -la gpurun_out/launches_c*_r02.csv
