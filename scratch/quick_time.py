import sys, time, os
sys.path.insert(0, '/root/repo')
import torch
from dgpmp2_b200 import ops, _lib
from dgpmp2_b200.datasets.synthetic import make_problems
from tests.gpu_helpers import cparams
B, T = int(sys.argv[1]), int(sys.argv[2])
pr = make_problems(B, T, unique_envs=64, seed=0)
dev = 'cuda'
th, start, goal, sdf = (pr[k].to(dev) for k in ('th_init', 'start', 'goal', 'sdf'))
cp = cparams(T)
print(ops.launch_shape(cparams(T, B=B), torch.float32))
for _ in range(5):
    out = ops.gn_step(cp, th, start, goal, sdf)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 50
e0.record()
for _ in range(N):
    out = ops.gn_step(cp, th, start, goal, sdf)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / N
print('B=%d T=%d  %.2f us/step  %.3g problem-iters/s  alg GB/s %.1f' % (B, T, ms * 1e3, B / ms * 1e3, B * (2*T*16+32+16*T+8) / ms / 1e6))
