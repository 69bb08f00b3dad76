#!/bin/bash
# builds the library of the last commit into scratch/exp/lib_prev.so (A/B baseline for uncommitted kernel changes)
rm -rf /tmp/prev && mkdir -p /tmp/prev/dgpmp2_b200/csrc /tmp/prev/include scratch/exp
for f in c_abi.cu kernels.cuh bcr.cuh bcr_plan.cuh factors.cuh hd.cuh mp.cuh host_params.h; do git show HEAD:dgpmp2_b200/csrc/$f > /tmp/prev/dgpmp2_b200/csrc/$f; done
git show HEAD:include/dgpmp2_b200.h > /tmp/prev/include/dgpmp2_b200.h
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared --cudart static -o scratch/exp/lib_prev.so /tmp/prev/dgpmp2_b200/csrc/c_abi.cu
