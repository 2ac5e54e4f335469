mkdir -p gpurun_out
B=kinovagrasping_b200/libkg_b200.so; P=build_variants/libkg_prev.so
for st in 1 2 4; do timeout 300 python tools/cmp_libs.py $P $B 2048 $st 2>&1 | tail -1 | cut -c60-330; done
timeout 1200 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2r_gputests.log 2>&1; echo "pytest rc=$?"; grep -E "identical contacts|same-set|solver on identical|passed|failed" gpurun_out/r2r_gputests.log | cut -c1-220
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-420
