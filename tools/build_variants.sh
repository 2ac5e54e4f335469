#!/bin/bash
# Build perf-experiment variants of the CUDA library into build_variants/ (git-ignored; they travel with gpurun).
# Usage: tools/build_variants.sh name:"flags" ...
mkdir -p build_variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -prec-div=false -prec-sqrt=false -ftz=true $flags -shared -Xcompiler -fPIC \
     -o build_variants/libkg_$name.so kinovagrasping_b200/csrc/kg_api.cu -Xptxas=-v > /tmp/bv_$name.log 2>&1 || { echo "BUILD FAILED: $name"; grep -i error /tmp/bv_$name.log | head; rm -f build_variants/libkg_$name.so; continue; }
  grep -A2 'entry function.*kg_sub_kernelILb0' /tmp/bv_$name.log | grep -E 'Used|spill' | tr '\n' ' '
  echo " <- $name"
done
