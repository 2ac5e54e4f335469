#!/bin/bash
# Under gpurun: time every build_variants/*.so (4096 and 32768 envs) -> gpurun_out/<tag>_probe.log
tag=${1:-probe}; shift
mkdir -p gpurun_out
for f in ${@:-build_variants/*.so}; do
  for n in 4096 32768; do timeout 200 python tools/perf_probe.py $f $n 12 >> gpurun_out/${tag}_probe.log 2>&1; done
done
cat gpurun_out/${tag}_probe.log
