mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2y_bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2y_bench_ref.json 2>> gpurun_out/r2y_bench.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2y_bench_ref.json
