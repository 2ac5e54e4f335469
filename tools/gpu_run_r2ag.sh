# final evidence at HEAD: gpu tests, smoke, launch list, bench, CPU arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2ag_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ag_pytest.log; grep -E "identical contacts|medium shapes|same-set|solver on identical|split vs|normals|passed|failed|rc=" gpurun_out/r2ag_pytest.log | cut -c1-230
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ag_smoke.log 2>&1; tail -1 gpurun_out/r2ag_smoke.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 400 --csv --log-file gpurun_out/r2ag_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra --no-learner --min-ms 1 > gpurun_out/r2ag_launches.log 2>&1; tail -2 gpurun_out/r2ag_launches.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/r2ag_bench.json 2> gpurun_out/r2ag_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r2ag_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2ag_bench_ref.json 2>> gpurun_out/r2ag_bench.err; cut -c1-200 gpurun_out/r2ag_bench_ref.json
