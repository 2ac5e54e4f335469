mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_gputests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2n_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2n_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2n_smoke.log
timeout 900 python bench.py > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2n_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2n_bench_ref.json 2>> gpurun_out/r2n_bench.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2n_bench_ref.json
