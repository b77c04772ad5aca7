#!/bin/bash
# Build variants of the library with extra -D flags (run in the build container): tools/build_variants.sh "name:flags" ...
# Output: build/variants/lib_<name>.so (git-ignored, travels with gpurun); time them with tools/time_variants.py
cd /root/repo/h1v2_isaac_b200/csrc
mkdir -p ../../build/variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 550,177 -use_fast_math -shared $flags \
    -o ../../build/variants/lib_$name.so h1v2_capi.cu h1v2_config.cpp &
  while [ $(jobs -r | wc -l) -ge 8 ]; do sleep 0.5; done
done
wait
ls -la ../../build/variants
