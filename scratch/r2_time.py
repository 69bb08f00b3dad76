"""Round-2 A/B timing of the GN step (device-resident, CUDA-graph replay): mixed-precision kernel vs the
all-double kernel, per BASELINE config, with the refinement statistics of the mixed-precision path."""
import ctypes, json, os, sys
sys.path.insert(0, '/root/repo')
import torch
from dgpmp2_b200 import _lib, ops
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
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


cases = [('config2 B=1024 T=64 d=4', 1024, 64, 2, {}), ('config3 shard B=1024 T=128 d=4', 1024, 128, 2, {}),
         ('config4 nonholonomic B=512 T=96 d=6', 512, 96, 3, dict(non_holonomic=True)),
         ('config5 shard vel-limits B=1024 T=64 d=4', 1024, 64, 2, dict(use_vel_limits=True)),
         ('B=8192 T=64 d=4', 8192, 64, 2, {}), ('B=1 T=64 d=4', 1, 64, 2, {})]
variants = [('fp64', dict(DGPMP2_PRECISION='64')), ('mp', dict()), ('mp accept=2^-12', dict(DGPMP2_MP_ACCEPT_LOG2='12')),
            ('mp maxt512', dict(DGPMP2_MP_MAXT='512'))]
if len(sys.argv) > 1:
    cases = [c for c in cases if any(k in c[0] for k in sys.argv[1].split(','))]
out = {}
vp = ctypes.c_void_p
lib = _lib.load()
for name, B, T, dof, flags in cases:
    base = XYH if dof == 3 else dict(YAML, K_v=0.01, v_x=1.0, v_y=1.0)
    pr = make_problems(B, T, dof=dof, unique_envs=128, seed=1)
    th, start, goal, sdf = (pr[k].cuda().contiguous() for k in ('th_init', 'start', 'goal', 'sdf'))
    cp = cparams(T, base=base, dof=dof, **flags)
    th = ops.gn_solve(cp, th, start, goal, sdf, 5, 0.0)[0]
    d = 2 * dof
    ref64 = ops.gn_step(cp, th.double(), start.double(), goal.double(), sdf.double())[0]
    dth = torch.empty_like(th); err = torch.empty(B, device='cuda'); ee = torch.empty(B, device='cuda')
    st = torch.zeros(B, dtype=torch.int32, device='cuda')
    cp.B = B; _lib.set_sdf_shape(cp, 128, 128, 128 * 128)
    s2, g2, sd = start.reshape(B, d).contiguous(), goal.reshape(B, d).contiguous(), sdf[:, 0].contiguous()

    def step():
        rc = lib.dgpmp2_gn_step_f32(ctypes.byref(cp), vp(th.data_ptr()), vp(s2.data_ptr()), vp(g2.data_ptr()), vp(sd.data_ptr()),
                                    None, vp(dth.data_ptr()), vp(err.data_ptr()), vp(ee.data_ptr()), vp(st.data_ptr()),
                                    vp(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
    out[name] = {}
    for vname, env in variants:
        if vname == 'mp maxt512' and T * ((B + 147) // 148) <= 512:
            continue
        saved = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            us = timeit(step, 50)
            rel = (torch.linalg.norm((dth.double() - ref64).reshape(B, -1), dim=1) / torch.linalg.norm(ref64.reshape(B, -1), dim=1))
            rec = {'us_per_step': round(us, 3), 'problem_iters_per_s': B / us * 1e6, 'max_rel_vs_f64io': float(rel.max()),
                   'status_max': int(st.abs().max()), 'launch': ops.launch_shape(cp)}
            if vname != 'fp64':
                refine = ops.gn_step_diag(cp, th, start, goal, sdf)[4]
                vals, cnt = torch.unique(refine, return_counts=True)
                rec['refine_hist'] = {int(v): int(c) for v, c in zip(vals.cpu(), cnt.cpu())}
        finally:
            for k, v in saved.items():
                if v is None: os.environ.pop(k, None)
                else: os.environ[k] = v
        out[name][vname] = rec
        print(name, '|', vname, json.dumps(rec), flush=True)
os.makedirs('/root/repo/gpurun_out', exist_ok=True)
json.dump(out, open('/root/repo/gpurun_out/r02_step_ab.json', 'w'), indent=1)
