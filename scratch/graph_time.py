# CUDA-graph timing of gn_step for one (B, T): python scratch/graph_time.py B T [dof]
import sys, os
sys.path.insert(0, '/root/repo')
import torch
from dgpmp2_b200 import _lib
if os.environ.get('DGPMP2_LIB'): _lib.LIB_PATH = os.environ['DGPMP2_LIB']
from dgpmp2_b200 import ops
from dgpmp2_b200.datasets.synthetic import make_problems
from tests.gpu_helpers import cparams
B, T = int(sys.argv[1]), int(sys.argv[2])
pr = make_problems(B, T, unique_envs=64, seed=0)
th, start, goal, sdf = (pr[k].cuda() for k in ('th_init', 'start', 'goal', 'sdf'))
cp = cparams(T)
shape = ops.launch_shape(cparams(T, B=B), torch.float32)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3):
        ops.gn_step(cp, th, start, goal, sdf)
    s.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        for _ in range(50):
            out = ops.gn_step(cp, th, start, goal, sdf)
    g.replay(); s.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    for _ in range(10):
        g.replay()
    e1.record(s); s.synchronize()
us = e0.elapsed_time(e1) / 500 * 1e3
print('B=%d T=%d %s  %.2f us/step  %.3g problem-iters/s' % (B, T, shape, us, B / us * 1e6))
