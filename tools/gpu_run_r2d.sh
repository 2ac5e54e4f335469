mkdir -p gpurun_out
python tools/trace_probe.py build_variants/libkg_trace.so 4096 5 > gpurun_out/r2d_trace.log 2>&1
python tools/trace_probe.py build_variants/libkg_trace.so 4096 18 >> gpurun_out/r2d_trace.log 2>&1
python tools/trace_probe.py build_variants/libkg_trace.so 1024 18 >> gpurun_out/r2d_trace.log 2>&1
python tools/trace_probe.py build_variants/libkg_trace.so 148 18 >> gpurun_out/r2d_trace.log 2>&1
cat gpurun_out/r2d_trace.log
