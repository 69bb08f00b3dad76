#!/bin/bash
# Round-2 runbook for the experimental fused assembly -> level-1 elimination (DESIGN.md section 8, item 0;
# kernels.cuh: DGPMP2_EXPERIMENTAL_FUSE1).  Compiled and inspected in round 1, NOT yet run on a GPU.
#   here (no GPU):   bash scratch/r2_fuse1.sh build        -> scratch/lib_fuse1.so
#   on the GPU box:  gpurun -- 'bash scratch/r2_fuse1.sh run'
set -u
cd "$(dirname "$0")/.."
if [ "${1:-run}" = build ]; then
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared --cudart static \
       -DDGPMP2_EXPERIMENTAL_FUSE1 -o scratch/lib_fuse1.so dgpmp2_b200/csrc/c_abi.cu && ls -la scratch/lib_fuse1.so
  exit $?
fi
O=gpurun_out/fuse1; mkdir -p $O
# 1. the whole GPU suite on the experimental build (results are expected to be BIT-identical to the shipped build)
DGPMP2_LIB=$PWD/scratch/lib_fuse1.so timeout 200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee $O/tests.txt
# 2. bitwise comparison shipped vs experimental on a few shapes
for lib in "" "$PWD/scratch/lib_fuse1.so"; do
  DGPMP2_LIB=$lib timeout 120 python - "$O/dth_${lib:+fuse1}.pt" <<'PY'
import sys, torch
sys.path.insert(0, '.')
from dgpmp2_b200 import ops
from dgpmp2_b200.datasets.synthetic import make_problems
from tests.gpu_helpers import cparams
out = {}
for B, T in ((1, 64), (7, 64), (1024, 64), (300, 101), (512, 128), (5, 3), (9, 2)):
    pr = make_problems(B, T, im_size=64, seed=B + T, unique_envs=8)
    th, start, goal, sdf = (pr[k].cuda() for k in ('th_init', 'start', 'goal', 'sdf'))
    th = ops.gn_solve(cparams(T), th, start, goal, sdf, 3, 0.0)[0]
    out[(B, T)] = [t.cpu() for t in ops.gn_step(cparams(T), th, start, goal, sdf)[:3]]
torch.save(out, sys.argv[1])
PY
done
python - <<PY | tee -a $O/tests.txt
import torch
a, b = torch.load('$O/dth_.pt'), torch.load('$O/dth_fuse1.pt')
print('bitwise identical:', all(torch.equal(x, y) for k in a for x, y in zip(a[k], b[k])))
PY
# 3. A/B timing (L2-resident inputs, CUDA graph of 50 launches)
for i in 1 2 3; do
  for lib in "" "$PWD/scratch/lib_fuse1.so"; do
    echo -n "${lib:+fuse1}${lib:-shipped}: " | tee -a $O/ab.txt
    DGPMP2_LIB=$lib timeout 60 python scratch/graph_time.py 1024 64 2>&1 | tail -1 | tee -a $O/ab.txt
  done
done
for shape in "1 64" "1024 128"; do
  for lib in "" "$PWD/scratch/lib_fuse1.so"; do
    echo -n "${lib:+fuse1}${lib:-shipped}: " | tee -a $O/ab.txt
    DGPMP2_LIB=$lib timeout 60 python scratch/graph_time.py $shape 2>&1 | tail -1 | tee -a $O/ab.txt
  done
done
