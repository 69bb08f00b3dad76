import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/examples')
import torch
import importlib.util
spec = importlib.util.spec_from_file_location('ex', '/root/repo/examples/learn_covariances_headless.py'); ex = importlib.util.module_from_spec(spec); spec.loader.exec_module(ex)
from dgpmp2_b200.datasets.synthetic import make_problems
T = 64; dtype = torch.float32
for B in (64, 512, 1024):
    pr = make_problems(B, T, im_size=64, seed=0, dtype=dtype)
    th0, start, goal, sdf, im = (pr[k].cuda().contiguous() for k in ('th_init', 'start', 'goal', 'sdf', 'im'))
    planner = ex.make_planner(T, dtype)
    planner.plan_layer.strict = False
    ref = planner.step(th0, start, goal, im, sdf)[0]
    print(B, 'static status nonzero', int((planner.plan_layer.last_status != 0).sum()))
    head = ex.CovarianceHead(T, 4, 0.01, dtype).cuda()
    planner.set_learn_module(head, 'diag_identity')
    with torch.no_grad():
        d = planner.step(th0, start, goal, im, sdf)[0]
    st = planner.plan_layer.last_status
    bad = torch.nonzero(st).reshape(-1)
    print(B, 'head status nonzero', int((st != 0).sum()), bad[:5].tolist(), st[bad[:5]].tolist(), 'max diff vs static', float((d - ref).abs().max()), 'nan', bool(torch.isnan(d).any()))
    out = head(th0, im, sdf)
    print('out finite', bool(torch.isfinite(out).all()), out.min().item(), out.max().item())
print('--- unrolled steps, B=1024')
B = 1024
pr = make_problems(B, T, im_size=64, seed=0, dtype=dtype)
th0, start, goal, sdf, im = (pr[k].cuda().contiguous() for k in ('th_init', 'start', 'goal', 'sdf', 'im'))
for use_head in (False, True):
    planner = ex.make_planner(T, dtype); planner.plan_layer.strict = False
    if use_head:
        planner.set_learn_module(ex.CovarianceHead(T, 4, 0.01, dtype).cuda(), 'diag_identity')
    th = th0
    for k in range(4):
        dth = planner.step(th, start, goal, im, sdf)[0]
        st = planner.plan_layer.last_status
        bad = torch.nonzero(st).reshape(-1)
        print('head' if use_head else 'static', 'step', k, 'bad', bad.numel(), bad[:4].tolist(), 'finite', bool(torch.isfinite(dth).all()), '|dth|max', float(dth.detach().abs().max()), '|th|max', float(th.detach().abs().max()))
        th = th + dth
