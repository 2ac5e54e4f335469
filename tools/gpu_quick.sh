#!/bin/bash
# Short GPU visit: bench lines + stage probes of the listed variant libraries.  Usage (under gpurun): bash tools/gpu_quick.sh <tag> [libs...]
tag=${1:-q}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --no-cpu --envs 32768 --steps 10 > gpurun_out/${tag}_bench32k.json 2>> gpurun_out/${tag}_bench.err
for f in "$@"; do
  for n in 4096 32768; do timeout 200 python tools/perf_probe.py $f $n 33 >> gpurun_out/${tag}_probe.log 2>&1; done
done
cat gpurun_out/${tag}_probe.log; cat gpurun_out/${tag}_bench.json gpurun_out/${tag}_bench32k.json; tail -3 gpurun_out/${tag}_bench.err
