#!/bin/bash
# gpurun with retry on "no box / slot free" (exit 3, nothing charged).  usage: scripts/gpu.sh <timeout_s> '<command>' [--gpus N]
T=$1; shift; CMD=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
