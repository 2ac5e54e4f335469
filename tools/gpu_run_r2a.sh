set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; tail -3 gpurun_out/r2a_smoke.log
for n in 4096 32768; do
 for gr in 0 1; do
  for G in 1 2 3 4; do
   KG_GRAPH=$gr KG_GROUPS=$G timeout 300 python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so $n 36 >> gpurun_out/r2a_probe.log 2>&1
  done
 done
done
cat gpurun_out/r2a_probe.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; tail -c 1500 gpurun_out/r2a_bench.json
