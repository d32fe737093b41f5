#!/bin/bash
# Round-2 re-capture after the epilogue rework of the persistent conv GEMM (run under gpurun on ONE GPU)
set -x
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r02b_launches_infer.csv python bench.py --steps 1 --warmup 3 --batches-per-step 2 --no-cpu-baseline --no-parity > gpurun_out/r02b_ncu_bench.log 2>&1
FULL="$NCU --set full --import-source on"
$FULL -k regex:tc_gemm2p -s 3 -c 1 -o gpurun_out/r02b_conv2 -f python scripts/profile_step.py tc 3 > /dev/null 2>&1
$FULL -k regex:tc_gemm2p -s 4 -c 1 -o gpurun_out/r02b_conv3 -f python scripts/profile_step.py tc 3 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
