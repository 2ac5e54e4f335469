mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2p_gputests.log 2>&1; echo "pytest rc=$?"; grep -E "identical contacts|same-set|solver on identical|narrowphase normals|passed|failed" gpurun_out/r2p_gputests.log | cut -c1-220
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2p_smoke.log | cut -c1-400
timeout 2400 python tools/success_sweep.py --starts 10000 --out gpurun_out/r2p_success_sweep.json > gpurun_out/r2p_success_sweep.log 2>&1; tail -2 gpurun_out/r2p_success_sweep.log
timeout 900 python bench.py > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/r2p_bench.json
