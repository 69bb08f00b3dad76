"""Device-timed numbers for the other BASELINE configs (not bench lines): per-GPU shard sizes."""
import sys, json
sys.path.insert(0, '/root/repo')
import torch
from dgpmp2_b200 import ops
from dgpmp2_b200.datasets.synthetic import make_problems
from tests.gpu_helpers import cparams
from tests.helpers import XYH, YAML

def timeit(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

out = {}
cases = [('config2 B=1024 T=64 d=4', 1024, 64, 2, {}), ('config3 shard B=1024 T=128 d=4', 1024, 128, 2, {}),
         ('config4 nonholonomic B=512 T=96 d=6', 512, 96, 3, dict(non_holonomic=True)),
         ('config5 shard vel-limits B=1024 T=64 d=4', 1024, 64, 2, dict(use_vel_limits=True)),
         ('B=8192 T=64 d=4', 8192, 64, 2, {}), ('B=148 T=64 d=4', 148, 64, 2, {}), ('B=1 T=64 d=4', 1, 64, 2, {})]
for name, B, T, dof, flags in cases:
    base = XYH if dof == 3 else dict(YAML, K_v=0.01, v_x=1.0, v_y=1.0)
    pr = make_problems(B, T, dof=dof, unique_envs=128, seed=1)
    th, start, goal, sdf = (pr[k].cuda().contiguous() for k in ('th_init', 'start', 'goal', 'sdf'))
    cp = cparams(T, base=base, dof=dof, **flags)
    th = ops.gn_solve(cp, th, start, goal, sdf, 3, 0.0)[0]
    d = 2 * dof
    dth = torch.empty_like(th); err = torch.empty(B, device='cuda'); ee = torch.empty(B, device='cuda'); st = torch.zeros(B, dtype=torch.int32, device='cuda')
    import ctypes
    from dgpmp2_b200 import _lib
    lib = _lib.load(); cp.B = B; _lib.set_sdf_shape(cp, 128, 128, 128 * 128)
    vp = ctypes.c_void_p
    s2, g2, sd = start.reshape(B, d).contiguous(), goal.reshape(B, d).contiguous(), sdf[:, 0].contiguous()
    def step():
        rc = lib.dgpmp2_gn_step_f32(ctypes.byref(cp), vp(th.data_ptr()), vp(s2.data_ptr()), vp(g2.data_ptr()), vp(sd.data_ptr()), None,
                                    vp(dth.data_ptr()), vp(err.data_ptr()), vp(ee.data_ptr()), vp(st.data_ptr()), vp(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
    us = timeit(step, 50)
    alg = B * (2 * T * d * 4 + 2 * d * 4 + 16 * T + 8)
    out[name] = {'us_per_step': us, 'problem_iters_per_s': B / us * 1e6, 'alg_GBs': alg / us / 1e3, 'launch': ops.launch_shape(cp)}
    print(name, json.dumps(out[name]), flush=True)
# persistent solver: 100 iterations, B=1024, T=64
pr = make_problems(1024, 64, unique_envs=128, seed=2)
th, start, goal, sdf = (pr[k].cuda().contiguous() for k in ('th_init', 'start', 'goal', 'sdf'))
cp = cparams(64)
for _ in range(2): r = ops.gn_solve(cp, th, start, goal, sdf, 100, 1e-4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); r = ops.gn_solve(cp, th, start, goal, sdf, 100, 1e-4); e1.record(); torch.cuda.synchronize()
its = r[1].float()
out['gn_solve B=1024 T=64 max_iters=100'] = {'ms': e0.elapsed_time(e1), 'mean_iters': float(its.mean()), 'problem_iters_per_s': float(its.sum()) / e0.elapsed_time(e1) * 1e3}
print(json.dumps(out['gn_solve B=1024 T=64 max_iters=100']))
# backward
from dgpmp2_b200.ops import gn_step_backward
pr = make_problems(1024, 64, unique_envs=128, seed=3)
th, start, goal, sdf = (pr[k].cuda().contiguous() for k in ('th_init', 'start', 'goal', 'sdf'))
dth = ops.gn_step(cp, th, start, goal, sdf)[0]
g = torch.randn_like(dth)
def bwd(): gn_step_backward(cp, th, start, goal, sdf, dth, g, None, need_th=True)
for _ in range(3): bwd()
torch.cuda.synchronize(); e0.record()
for _ in range(20): bwd()
e1.record(); torch.cuda.synchronize()
out['gn_step_backward B=1024 T=64 (g_th only)'] = {'us': e0.elapsed_time(e1) / 20 * 1e3}
print(json.dumps(out['gn_step_backward B=1024 T=64 (g_th only)']))
json.dump(out, open('/root/repo/gpurun_out/r01_extra_timings.json', 'w'), indent=1)
