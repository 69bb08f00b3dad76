"""Device time of the backward of the GN step through the C ABI with preallocated outputs (CUDA graph of 50 launches), next to the forward."""
import ctypes, sys, json, os
sys.path.insert(0, '/root/repo')
import torch
from dgpmp2_b200 import _lib, ops
import bench
B, T, d = 1024, 64, 4
dev = torch.device('cuda', 0)
cp = bench.make_cparams()
pr = bench.make_inputs(0, 1, B)[0]
th, start, goal, sdf = (pr[k].to(dev).contiguous() for k in ('th_init', 'start', 'goal', 'sdf'))
th = ops.gn_solve(cp, th, start, goal, sdf, 5, 0.0)[0].contiguous()
dth, err, ee, st = ops.gn_step(cp, th, start, goal, sdf)
_lib.set_sdf_shape(cp, 128, 128, 128 * 128); cp.B = B
g = torch.randn_like(dth); g_th = torch.empty_like(th); g_ee = torch.ones(B, device=dev)
g_start = torch.empty(B, d, device=dev); g_goal = torch.empty(B, d, device=dev)
g_w = torch.empty(B, T, device=dev); g_eps = torch.empty(B, T, device=dev); g_qc = torch.empty(B, T - 1, 2, 2, device=dev)
g_sdf = torch.zeros(B, 128, 128, device=dev)
lib = _lib.load(); vp = ctypes.c_void_p
s2, g2, sd2 = start.reshape(B, d).contiguous(), goal.reshape(B, d).contiguous(), sdf[:, 0].contiguous()
def P(t): return vp(t.data_ptr()) if t is not None else None
def timed(fn, n=50):
    for _ in range(3): fn(vp(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        s = vp(torch.cuda.current_stream().cuda_stream)
        for _ in range(n): fn(s)
    gr.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return round(best, 2)
def fwd(s): assert lib.dgpmp2_gn_step_f32(ctypes.byref(cp), P(th), P(s2), P(g2), P(sd2), None, P(dth), P(err), P(ee), P(st), s) == 0
def bwd(need, s):
    a = dict(g_ee=None, g_th=g_th, g_start=None, g_goal=None, g_qc=None, g_w=None, g_eps=None, g_sdf=None); a.update(need)
    assert lib.dgpmp2_gn_step_backward_f32(ctypes.byref(cp), P(th), P(s2), P(g2), P(sd2), None, P(dth), P(g), P(a['g_ee']), P(a['g_th']), P(a['g_start']),
                                           P(a['g_goal']), P(a['g_qc']), P(a['g_w']), P(a['g_eps']), P(a['g_sdf']), s) == 0
out = {'forward': timed(fwd), 'backward g_th': timed(lambda s: bwd({}, s)),
       'backward g_th + g_err_ext': timed(lambda s: bwd(dict(g_ee=g_ee), s)),
       'backward g_th,start,goal,w,eps,qc': timed(lambda s: bwd(dict(g_ee=g_ee, g_start=g_start, g_goal=g_goal, g_w=g_w, g_eps=g_eps, g_qc=g_qc), s)),
       'backward all seven (sdf atomics)': timed(lambda s: bwd(dict(g_ee=g_ee, g_start=g_start, g_goal=g_goal, g_w=g_w, g_eps=g_eps, g_qc=g_qc, g_sdf=g_sdf), s))}
print(json.dumps(out))
