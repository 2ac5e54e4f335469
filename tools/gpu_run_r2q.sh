mkdir -p gpurun_out
B=kinovagrasping_b200/libkg_b200.so; P=build_variants/libkg_prev.so
timeout 300 python tools/cmp_libs.py $P $B 2048 12 2>&1 | tail -1 | cut -c1-260
KG_SHAPES=6,7,8,9,10,11,12,13 KG_ORIENT=3 KG_MODE=1 timeout 300 python tools/cmp_libs.py $P $B 2048 12 2>&1 | tail -1 | cut -c1-260
L=gpurun_out/r2q_probe.log; rm -f $L
for n in 4096 32768; do for lib in $P $B $P $B; do timeout 300 python tools/perf_probe.py $lib $n 36 >> $L 2>&1; done; done
cut -c1-200 $L
