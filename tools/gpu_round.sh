#!/bin/bash
# One GPU visit: parity tests, smoke, bench (4096 and 32768 envs), per-launch-kind timing, ncu launch list of bench.py and
# `--set full` captures of the two hot kernels in a contact-free and a contact-rich RL step.
# Usage (under gpurun): bash tools/gpu_round.sh <tag> [nocapture]
tag=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --no-cpu --envs 32768 --steps 30 > gpurun_out/${tag}_bench32k.json 2>> gpurun_out/${tag}_bench.err
for n in 4096 32768; do
  KG_TIMING=1 timeout 200 python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so $n 33 >> gpurun_out/${tag}_probe.log 2>&1
  timeout 200 python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so $n 33 >> gpurun_out/${tag}_probe.log 2>&1
done
tail -3 gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_smoke.log gpurun_out/${tag}_bench.json gpurun_out/${tag}_bench32k.json gpurun_out/${tag}_probe.log; tail -3 gpurun_out/${tag}_bench.err
[ "$2" = "nocapture" ] && exit 0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_launch.log 2>&1
# perf_probe, 2 groups: kg_sub_kernel (both template instances) launches 30 times per RL step, kg_mpr_kernel 30 times; step 5 is contact-free, step 18 contact-rich
bash tools/gpu_ncu2.sh ${tag}_sub_light kinovagrasping_b200/libkg_b200.so kg_sub_kernel 154 4096 8
bash tools/gpu_ncu2.sh ${tag}_sub_heavy kinovagrasping_b200/libkg_b200.so kg_sub_kernel 544 4096 20
bash tools/gpu_ncu2.sh ${tag}_mpr_heavy kinovagrasping_b200/libkg_b200.so kg_mpr_kernel 545 4096 20
ls -la gpurun_out
