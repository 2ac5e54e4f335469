#!/bin/bash
# One GPU visit: parity tests, smoke, bench, stage profile, ncu launch list + full capture of kg_step_kernel.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${tag}_smoke.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?" >> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --no-cpu --envs 32768 --steps 10 > gpurun_out/${tag}_bench32k.json 2>> gpurun_out/${tag}_bench.err
for n in 4096 32768; do
  timeout 200 python tools/perf_probe.py build_variants/libkg_prof.so $n 12 >> gpurun_out/${tag}_probe.log 2>&1
  timeout 200 python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so $n 12 >> gpurun_out/${tag}_probe.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kg_step_kernel -s 6 -c 1 -f -o gpurun_out/${tag}_step python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so 4096 8 > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out
