#!/bin/bash
# build variants of the library with different unroll factors (run in the build container), then time them on the GPU box
cd /root/repo/h1v2_isaac_b200/csrc
i=0
for v in "" "-DU_SWEEP2=2" "-DU_MPROD=2" "-DU_LSJ=2" "-DU_EVJ=2" "-DU_SWEEP2=2 -DU_MPROD=2" "-DU_SWEEP2=3 -DU_MPROD=3" "-DU_LSJ=3 -DU_EVJ=3" "-DU_SWEEP1=2" "-DU_SWEEP2=6 -DU_MPROD=6"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 550,177 -shared $v -o ../lib_var$i.so h1v2_capi.cu h1v2_config.cpp &
  echo "$i: $v" ; i=$((i+1))
done > /root/repo/gpurun_out/variants.txt
wait
