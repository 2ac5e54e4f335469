mkdir -p gpurun_out
L=gpurun_out/r2h_probe.log
for v in teams2 teams4; do python tools/cmp_libs.py kinovagrasping_b200/libkg_b200.so build_variants/libkg_$v.so >> gpurun_out/r2h_cmp.log 2>&1; done
cat gpurun_out/r2h_cmp.log | cut -c1-200
for lib in kinovagrasping_b200/libkg_b200.so build_variants/libkg_teams2.so build_variants/libkg_teams4.so build_variants/libkg_teams2m4.so; do
 for n in 4096 32768; do KG_GROUPS=4 timeout 300 python tools/perf_probe.py $lib $n 36 >> $L 2>&1; done
 KG_TIMING=1 timeout 300 python tools/perf_probe.py $lib 4096 36 >> $L 2>&1
done
KG_SHAPES=6,7,8,9,10,11,12,13 KG_ORIENT=3 KG_VARIANT=0 timeout 300 python tools/perf_probe.py build_variants/libkg_teams2.so 32768 36 >> $L 2>&1
cat $L
