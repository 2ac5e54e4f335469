# success sweep on the seven medium shapes, grasp-reward overhead, 21-shape random-orientation batch
mkdir -p gpurun_out
timeout 600 python tools/success_sweep.py --shapes 14,15,16,17,18,19,20 --out gpurun_out/r2ab_success_sweep_medium.json > gpurun_out/r2ab_success_sweep_medium.log 2>&1; cat gpurun_out/r2ab_success_sweep_medium.log | cut -c1-160
B=kinovagrasping_b200/libkg_b200.so; L=gpurun_out/r2ab_probe.log; rm -f $L
for n in 4096 32768; do for g in "" 1 "" 1; do KG_GRASP=$g timeout 300 python tools/perf_probe.py $B $n 36 >> $L 2>&1; done; done
KG_SHAPES=0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20 KG_ORIENT=3 timeout 300 python tools/perf_probe.py $B 32768 36 >> $L 2>&1
cut -c1-230 $L
