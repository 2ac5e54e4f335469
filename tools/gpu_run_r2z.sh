# final evidence of round 2 at HEAD (second pass, after the solver restructuring)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2w_smi.txt
B=kinovagrasping_b200/libkg_b200.so
L=gpurun_out/r2w_probe_solver_passes.txt; rm -f $L
for n in 4096 32768; do for lib in build_variants/libkg_bc0f46b.so $B build_variants/libkg_bc0f46b.so $B; do timeout 300 python tools/perf_probe.py $lib $n 36 >> $L 2>&1; done; done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log; tail -3 gpurun_out/r2w_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.log 2>&1; tail -1 gpurun_out/r2w_smoke.log | cut -c1-300
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum
KG_GROUPS=1 KG_GRAPH=0 timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2w_flops.csv python tools/perf_probe.py $B 4096 30 > gpurun_out/r2w_flops.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 400 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra --no-learner --min-ms 1 > gpurun_out/r2w_launches.log 2>&1
KG_GROUPS=1 KG_GRAPH=0 bash tools/gpu_ncu2.sh r2w_sub_light $B kg_sub_kernel 170 4096 24
KG_GROUPS=1 KG_GRAPH=0 bash tools/gpu_ncu2.sh r2w_mpr_heavy $B kg_mpr_kernel 290 4096 24
timeout 2400 python tools/success_sweep.py --starts 10000 --out gpurun_out/r2w_success_sweep.json > gpurun_out/r2w_success_sweep.log 2>&1; tail -2 gpurun_out/r2w_success_sweep.log
KG_SHAPES=0,1,2,3,4,5,6,7,8,9,10,11,12,13 KG_ORIENT=3 KG_VARIANT=0 timeout 300 python tools/perf_probe.py $B 32768 33 > gpurun_out/r2w_config3.txt 2>&1; cut -c1-200 gpurun_out/r2w_config3.txt
timeout 900 python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2w_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2w_bench_ref.json 2>> gpurun_out/r2w_bench.err; cut -c1-160 gpurun_out/r2w_bench_ref.json
