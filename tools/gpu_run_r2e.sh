mkdir -p gpurun_out
L=gpurun_out/r2e_probe.log
for G in 1 2 4 8 16 32; do KG_GROUPS=$G timeout 300 python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so 4096 36 >> $L 2>&1; done
for G in 2 8 32; do KG_GROUPS=$G timeout 300 python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so 32768 36 >> $L 2>&1; done
for lib in w4x7 w14x2 half; do for G in 1 8 16; do KG_GROUPS=$G timeout 300 python tools/perf_probe.py build_variants/libkg_$lib.so 4096 36 >> $L 2>&1; done; done
KG_GROUPS=8 timeout 300 python tools/perf_probe.py build_variants/libkg_w4x7.so 32768 36 >> $L 2>&1
KG_GROUPS=8 timeout 300 python tools/perf_probe.py build_variants/libkg_half.so 32768 36 >> $L 2>&1
cat $L
python tools/cmp_libs.py kinovagrasping_b200/libkg_b200.so build_variants/libkg_w4x7.so
