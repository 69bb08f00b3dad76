"""Backward of the GN step: device time with g_th only and with all seven gradients (incl. the SDF atomics), vs forward."""
import sys, json
sys.path.insert(0, '/root/repo')
import torch
from dgpmp2_b200 import ops
from dgpmp2_b200.datasets.synthetic import make_problems
from tests.gpu_helpers import cparams
B, T = 1024, 64
pr = make_problems(B, T, unique_envs=128, seed=3)
th, start, goal, sdf = (pr[k].cuda().contiguous() for k in ('th_init', 'start', 'goal', 'sdf'))
cp = cparams(T)
th = ops.gn_solve(cp, th, start, goal, sdf, 5, 0.0)[0]
qc = torch.eye(2, device='cuda').reshape(1, 1, 2, 2).expand(B, T - 1, 2, 2).contiguous()
w = torch.full((B, T, 1, 1), 1e4, device='cuda')
eps = torch.full((B, T, 1, 1), 0.4, device='cuda')
dth = ops.gn_step(cp, th, start, goal, sdf)[0]
g = torch.randn_like(dth)


def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g_ = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_):
        for _ in range(n): fn()
    g_.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g_.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

out = {'forward static (us)': timed(lambda: ops.gn_step(cp, th, start, goal, sdf)),
       'forward learned weights (us)': timed(lambda: ops.gn_step(cp, th, start, goal, sdf, qc_inv=qc, w_obs=w, eps=eps)),
       'backward g_th only (us)': timed(lambda: ops.gn_step_backward(cp, th, start, goal, sdf, dth, g, None, need_th=True)),
       'backward all seven gradients (us)': timed(lambda: ops.gn_step_backward(
           cp, th, start, goal, sdf, dth, g, None, qc_inv=qc, w_obs=w, eps=eps, need_th=True, need_start=True, need_goal=True,
           need_qc=True, need_w=True, need_eps=True, need_sdf=True)),
       'errors (us)': timed(lambda: ops.errors(cp, th, start, goal, sdf)),
       'errors backward (us)': timed(lambda: ops.errors_backward(cp, th, start, goal, sdf, torch.ones(B, device='cuda'), None, torch.ones(B, device='cuda'), torch.ones(B, device='cuda')))}
print(json.dumps(out, indent=1))
