mkdir -p gpurun_out
timeout 300 python tools/e2e_probe.py 4096 300 2>&1 | tail -2 | tee gpurun_out/r2ai_e2e_probe.log
