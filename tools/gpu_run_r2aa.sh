# gpu test suite, smoke and bench at HEAD (grasp-classifier kernels, 21 shapes)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2aa_pytest.log; grep -E "identical contacts|medium shapes|same-set|solver on identical|split vs|normals|passed|failed|rc=|Error|error" gpurun_out/r2aa_pytest.log | cut -c1-230
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2aa_smoke.log 2>&1; tail -1 gpurun_out/r2aa_smoke.log | cut -c1-330
timeout 600 python bench.py > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2aa_bench.json
