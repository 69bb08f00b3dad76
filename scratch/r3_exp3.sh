#!/bin/bash
mkdir -p gpurun_out/r3c; cd /root/repo
for cta in 74; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared --cudart static -DDGPMP2_TIMING=$cta -o scratch/libdgpmp2_timing.so dgpmp2_b200/csrc/c_abi.cu 2> gpurun_out/r3c/build.err || exit 1
  for cfg in "1 64" "1024 64" "1024 128"; do
    echo "== cta $cta B,T = $cfg" >> gpurun_out/r3c/phases.txt
    python scratch/phase_time.py $cfg >> gpurun_out/r3c/phases.txt 2>&1
  done
done
cat gpurun_out/r3c/phases.txt
