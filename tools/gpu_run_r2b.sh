set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -5 gpurun_out/r2b_pytest.log
for lib in kinovagrasping_b200/libkg_b200.so build_variants/libkg_d8x3.so build_variants/libkg_d7x4mpr4.so build_variants/libkg_d7x4mpr5.so; do
 for n in 4096 32768; do
  for G in 1 2; do
   KG_GROUPS=$G timeout 300 python tools/perf_probe.py $lib $n 36 >> gpurun_out/r2b_probe.log 2>&1
  done
 done
done
KG_TIMING=1 timeout 300 python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so 4096 36 >> gpurun_out/r2b_probe.log 2>&1
cat gpurun_out/r2b_probe.log
