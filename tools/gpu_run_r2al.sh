# final evidence at HEAD: gpu tests, smoke, launch list, bench, CPU arm, a DDPGfD training run with the grasp-classifier reward and eval outputs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2al_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2al_pytest.log; grep -E "passed|failed|rc=" gpurun_out/r2al_pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2al_smoke.log 2>&1; tail -1 gpurun_out/r2al_smoke.log | cut -c1-120
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 400 --csv --log-file gpurun_out/r2al_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extra --no-learner --min-ms 1 > gpurun_out/r2al_launches.log 2>&1
timeout 900 python bench.py > gpurun_out/r2al_bench.json 2> gpurun_out/r2al_bench.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/r2al_bench.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2al_bench_ref.json 2>> gpurun_out/r2al_bench.err; cut -c1-160 gpurun_out/r2al_bench_ref.json
rm -rf gpurun_out/r2al_evalout
t0=$(date +%s); timeout 600 python -m kinovagrasping_b200.train --envs 4096 --sweeps 40 --tau 0.01 --with-grasp-reward --eval-out gpurun_out/r2al_evalout > gpurun_out/r2al_train.log 2>&1; echo "train rc=$? wall=$(( $(date +%s) - t0 ))s" >> gpurun_out/r2al_train.log
grep -E '"eval"|train rc' gpurun_out/r2al_train.log | cut -c1-200 | tail -6
(cd gpurun_out/r2al_evalout && find . -type f | sort | head -40) > gpurun_out/r2al_evalout_files.txt; wc -l gpurun_out/r2al_evalout_files.txt; rm -rf gpurun_out/r2al_evalout
