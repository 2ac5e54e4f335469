mkdir -p gpurun_out
B=kinovagrasping_b200/libkg_b200.so; L=gpurun_out/r2af_probe.log; rm -f $L
for n in 4096; do for p in 0 1 0 1 0 1; do KG_PDL=$p timeout 300 python tools/perf_probe.py $B $n 36 >> $L 2>&1; done; done
for n in 32768; do for p in 0 1; do KG_PDL=$p timeout 300 python tools/perf_probe.py $B $n 36 >> $L 2>&1; done; done
cut -c1-250 $L
