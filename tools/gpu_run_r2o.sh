#!/bin/bash
# Round 2: attention time with the clock it ran at; retrieval incl. the binary branch.
mkdir -p gpurun_out
for v in 5 65; do timeout -s KILL 200 python tools/att_clock_probe.py $v 2>&1 | grep variant | tee -a gpurun_out/att_clock_probe.log; done
timeout -s KILL 600 python tools/bench_topk.py 2000000 768 > gpurun_out/topk_r02b.log 2>&1; echo "topk rc=$?"; grep ubinary gpurun_out/topk_r02b.log
