#!/bin/bash
# GPU visit for the wavefront pipeline: parity tests, variant probes, per-kernel launch durations.
# Usage (under gpurun): bash tools/gpu_wave.sh <tag> [variant libs...]
tag=${1:-w}; shift
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
for f in kinovagrasping_b200/libkg_b200.so "$@"; do
  for n in 4096 32768; do timeout 200 python tools/perf_probe.py $f $n 33 >> gpurun_out/${tag}_probe.log 2>&1; done
done
cat gpurun_out/${tag}_probe.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so 4096 12 > gpurun_out/${tag}_ncu_launch.log 2>&1
tail -2 gpurun_out/${tag}_ncu_launch.log
