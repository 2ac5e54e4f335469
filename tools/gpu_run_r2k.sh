mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2k_smi.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log; tail -3 gpurun_out/r2k_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; tail -1 gpurun_out/r2k_smoke.log | cut -c1-300
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; tail -c 400 gpurun_out/r2k_bench.json; tail -3 gpurun_out/r2k_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2k_bench_reference.json 2>/dev/null; tail -c 300 gpurun_out/r2k_bench_reference.json
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum
KG_GROUPS=1 KG_GRAPH=0 timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2k_flops.csv python tools/perf_probe.py kinovagrasping_b200/libkg_b200.so 4096 30 > gpurun_out/r2k_flops.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 400 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra --no-learner --min-ms 1 > gpurun_out/r2k_launches.log 2>&1
KG_GROUPS=1 KG_GRAPH=0 bash tools/gpu_ncu2.sh r2k_sub_light kinovagrasping_b200/libkg_b200.so kg_sub_kernel 170 4096 24
KG_GROUPS=1 KG_GRAPH=0 bash tools/gpu_ncu2.sh r2k_sub_heavy kinovagrasping_b200/libkg_b200.so kg_sub_kernel 270 4096 24
KG_GROUPS=1 KG_GRAPH=0 bash tools/gpu_ncu2.sh r2k_mpr_heavy kinovagrasping_b200/libkg_b200.so kg_mpr_kernel 290 4096 24
KG_GROUPS=1 KG_GRAPH=0 bash tools/gpu_ncu2.sh r2k_rays kinovagrasping_b200/libkg_b200.so kg_rays_kernel 18 4096 24
timeout 2400 python tools/success_sweep.py --starts 10000 --out gpurun_out/r2k_success_sweep.json > gpurun_out/r2k_success_sweep.log 2>&1; tail -3 gpurun_out/r2k_success_sweep.log
