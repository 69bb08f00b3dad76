# ncu target for the isolated-phase builds (see DGPMP2_REP in bcr.cuh): python scratch/rep_prof.py <lib.so> [B T]
import sys
sys.path.insert(0, '/root/repo')
import torch
from dgpmp2_b200 import _lib
_lib.LIB_PATH = sys.argv[1]
from dgpmp2_b200 import ops
from dgpmp2_b200.datasets.synthetic import make_problems
from tests.gpu_helpers import cparams
B, T = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1024, 64)
pr = make_problems(B, T, unique_envs=64, seed=0)
th, start, goal, sdf = (pr[k].cuda() for k in ('th_init', 'start', 'goal', 'sdf'))
cp = cparams(T)
for _ in range(3):
    ops.gn_step(cp, th, start, goal, sdf)
torch.cuda.synchronize()
