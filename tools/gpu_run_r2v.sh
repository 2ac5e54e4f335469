# gpu test suite, smoke, bench and the CPU arm at HEAD
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log; grep -E "identical contacts|same-set|solver on identical|split vs|normals|passed|failed|rc=" gpurun_out/r2v_pytest.log | cut -c1-230
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; tail -1 gpurun_out/r2v_smoke.log | cut -c1-330
timeout 900 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2v_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2v_bench_ref.json 2>> gpurun_out/r2v_bench.err; cut -c1-160 gpurun_out/r2v_bench_ref.json
