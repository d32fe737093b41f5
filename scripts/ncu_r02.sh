#!/bin/bash
# Round-2 ncu captures (run under gpurun on ONE GPU; reports land in gpurun_out/, summaries are made here with scripts/ncu_summary.py)
set -x
NCU="ncu --clock-control none"
# every launch of two bench batches with its device time (cold-cache, serialised: compare SHARES)
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/r02_launches_infer.csv python bench.py --steps 1 --warmup 3 --batches-per-step 2 --no-cpu-baseline --no-parity > gpurun_out/r02_ncu_bench.log 2>&1
FULL="$NCU --set full --import-source on"
$FULL -k regex:tc_conv1_u8 -s 2 -c 1 -o gpurun_out/r02_conv1 -f python scripts/profile_step.py tc 4 > /dev/null 2>&1
$FULL -k regex:tc_gemm2p -s 3 -c 1 -o gpurun_out/r02_conv2 -f python scripts/profile_step.py tc 3 > /dev/null 2>&1
$FULL -k regex:tc_match2 -s 2 -c 1 -o gpurun_out/r02_match_b256 -f python scripts/profile_step.py tc 4 > /dev/null 2>&1
for B in 1 32 128; do
  $FULL -k regex:tc_match2 -s 3 -c 1 -o gpurun_out/r02_match_b$B -f python scripts/match_bench.py --batches $B --iters 6 > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep
