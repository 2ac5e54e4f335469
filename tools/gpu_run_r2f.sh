mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -6 gpurun_out/r2f_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1; tail -2 gpurun_out/r2f_smoke.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 3000 gpurun_out/r2f_bench.json; tail -5 gpurun_out/r2f_bench.err
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum
KG_GROUPS=1 KG_GRAPH=0 timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2f_flops.csv python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so 4096 30 > gpurun_out/r2f_flops.log 2>&1
tail -2 gpurun_out/r2f_flops.log; wc -l gpurun_out/r2f_flops.csv
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra --no-learner --min-ms 1 > gpurun_out/r2f_launches.log 2>&1
tail -2 gpurun_out/r2f_launches.log
KG_GROUPS=1 KG_GRAPH=0 bash tools/gpu_ncu2.sh r2f_sub_light kinovagrasping_b200/libkg_b200.so "kg_sub_kernel<\(bool\)0>" 170 4096 24
KG_GROUPS=1 KG_GRAPH=0 bash tools/gpu_ncu2.sh r2f_sub_heavy kinovagrasping_b200/libkg_b200.so "kg_sub_kernel<\(bool\)0>" 330 4096 24
KG_GROUPS=1 KG_GRAPH=0 bash tools/gpu_ncu2.sh r2f_mpr_heavy kinovagrasping_b200/libkg_b200.so "kg_mpr_kernel" 350 4096 24
