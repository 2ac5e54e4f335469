mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity.py -m gpu -q -x -k "dependent_launch or graph_replay or pipeline_groups" 2>&1 | tail -3
B=kinovagrasping_b200/libkg_b200.so; L=gpurun_out/r2ae_probe.log; rm -f $L
for n in 4096 32768; do for p in 0 1 0 1; do KG_PDL=$p timeout 300 python tools/perf_probe.py $B $n 36 >> $L 2>&1; done; done
KG_PDL=0 KG_GRAPH=0 timeout 300 python tools/perf_probe.py $B 4096 36 >> $L 2>&1
KG_PDL=1 KG_GRAPH=0 timeout 300 python tools/perf_probe.py $B 4096 36 >> $L 2>&1
cut -c1-250 $L
