#!/bin/bash
# Probe a list of variant libraries at 4096 and 32768 envs.  Usage (under gpurun): bash tools/gpu_probe.sh <tag> libs...
tag=$1; shift
mkdir -p gpurun_out
for f in "$@"; do
  for n in 4096 32768; do timeout 200 python tools/perf_probe.py $f $n 33 >> gpurun_out/${tag}_probe.log 2>&1; done
done
cat gpurun_out/${tag}_probe.log
