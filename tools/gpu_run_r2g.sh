mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -4 gpurun_out/r2g_pytest.log
python tools/cmp_libs.py kinovagrasping_b200/libkg_b200.so build_variants/libkg_w14x2.so > gpurun_out/r2g_cmp.log 2>&1; cat gpurun_out/r2g_cmp.log
L=gpurun_out/r2g_probe.log
for G in 1 2 4 8; do KG_GROUPS=$G timeout 300 python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so 4096 36 >> $L 2>&1; done
for G in 2 4; do KG_GROUPS=$G timeout 300 python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so 32768 36 >> $L 2>&1; done
KG_TIMING=1 timeout 300 python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so 4096 36 >> $L 2>&1
KG_SHAPES=6,7,8,9,10,11,12,13 KG_ORIENT=3 KG_VARIANT=0 timeout 300 python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so 32768 36 >> $L 2>&1
KG_SHAPES=0,1,2,3,4,5,6,7,8,9,10,11,12,13 KG_ORIENT=3 KG_VARIANT=0 timeout 300 python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so 32768 36 >> $L 2>&1
cat $L
