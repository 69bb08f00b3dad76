"""A few GN steps at the given size (for ncu captures)."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from dgpmp2_b200 import ops
from dgpmp2_b200.datasets.synthetic import make_problems
from tests.gpu_helpers import cparams
B, T = int(sys.argv[1]), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
pr = make_problems(B, T, unique_envs=min(B, 128), seed=0)
th, start, goal, sdf = (pr[k].cuda() for k in ('th_init', 'start', 'goal', 'sdf'))
cp = cparams(T)
th = ops.gn_solve(cp, th, start, goal, sdf, 5, 0.0)[0]
for _ in range(n):
    out = ops.gn_step(cp, th, start, goal, sdf)
torch.cuda.synchronize()
print('ok', float(out[0].abs().max()))
