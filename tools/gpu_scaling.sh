#!/bin/bash
# SURVEY 8d config 3 (Exp1: 14 shapes, random hand orientation, 32k envs) and config 5 (env-count sweep 2^10..2^20 on the
# cone/vase shapes, random orientation): device time of the kernel chain per RL step over one 30-step episode cycle.
# Usage (under gpurun): bash tools/gpu_scaling.sh <tag>
tag=${1:-scal}; mkdir -p gpurun_out; out=gpurun_out/${tag}_scaling.log; : > $out
echo "# config 3: Exp1 shape set, random orientation (code-faithful variant A), 32768 envs" >> $out
KG_SHAPES=0,1,2,3,4,5,6,7,8,9,10,11,12,13 KG_ORIENT=3 KG_VARIANT=0 timeout 300 python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so 32768 33 >> $out 2>&1
echo "# config 5: cone/vase shapes, random orientation, env-count sweep" >> $out
for n in 1024 2048 4096 8192 16384 32768 65536 131072 262144 524288 1048576; do
  KG_SHAPES=6,7,8,9,10,11,12,13 KG_ORIENT=3 KG_VARIANT=0 timeout 600 python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so $n 33 >> $out 2>&1
done
cat $out
