mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2ak_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ak_pytest.log; grep -E "identical contacts|medium shapes|deeper than|same-set|passed|failed|rc=|Error" gpurun_out/r2ak_pytest.log | cut -c1-230
timeout 600 python tools/success_sweep.py --shapes 21,22,23 --out gpurun_out/r2ak_success_sweep_lemon.json > gpurun_out/r2ak_success_sweep_lemon.log 2>&1; cat gpurun_out/r2ak_success_sweep_lemon.log | cut -c1-160
B=kinovagrasping_b200/libkg_b200.so; L=gpurun_out/r2ak_probe.log; rm -f $L
timeout 300 python tools/perf_probe.py $B 4096 36 >> $L 2>&1
timeout 300 python tools/perf_probe.py $B 32768 36 >> $L 2>&1
KG_SHAPES=21,22,23 KG_ORIENT=3 timeout 300 python tools/perf_probe.py $B 4096 36 >> $L 2>&1
KG_SHAPES=0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23 KG_ORIENT=3 timeout 300 python tools/perf_probe.py $B 32768 36 >> $L 2>&1
cut -c1-230 $L
