#!/bin/bash
# ncu --set full capture of one launch deep inside an episode.  Usage (under gpurun): bash tools/gpu_ncu2.sh <tag> <lib> <kernel-regex> <skip> [envs] [steps]
tag=$1; lib=$2; k=$3; skip=$4; envs=${5:-4096}; steps=${6:-24}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/${tag} python tools/perf_probe.py $lib $envs $steps > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log; ls -la gpurun_out/${tag}.ncu-rep
