mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitizer_probe.py 700 6 1 > gpurun_out/r2x_memcheck_split.log 2>&1; tail -3 gpurun_out/r2x_memcheck_split.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitizer_probe.py 700 4 2 > gpurun_out/r2x_memcheck_full.log 2>&1; tail -3 gpurun_out/r2x_memcheck_full.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitizer_probe.py 64 4 1 > gpurun_out/r2x_racecheck.log 2>&1; tail -3 gpurun_out/r2x_racecheck.log
timeout 600 compute-sanitizer --tool initcheck --print-limit 20 python tools/sanitizer_probe.py 128 3 1 > gpurun_out/r2x_initcheck.log 2>&1; tail -3 gpurun_out/r2x_initcheck.log
