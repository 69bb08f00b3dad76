"""Shape sweep of gn_step on the GPU: normal-equation residual of every problem (band kernel) for many (B, T, dof),
default schedule and forced one-lane / multi-problem schedules.  python scratch/stress.py"""
import sys, os, itertools
sys.path.insert(0, '/root/repo')
import torch
from dgpmp2_b200 import ops
from dgpmp2_b200.datasets.synthetic import make_problems
from tests.gpu_helpers import cparams
from tests.helpers import XYH, YAML

def band_matvec(D, U, x):
    y = torch.einsum('btij,btj->bti', D, x)
    y[:, :-1] += torch.einsum('btij,btj->bti', U, x[:, 1:])
    y[:, 1:] += torch.einsum('btji,btj->bti', U, x[:, :-1])
    return y

worst = 0.0; n = 0
Ts = list(range(2, 41)) + [47, 63, 64, 65, 95, 96, 97, 100, 127, 128, 129, 200, 233]
for env in ({}, {'DGPMP2_WIDE': '4'}, {'DGPMP2_WIDE': '4', 'DGPMP2_NP': '3'}, {'DGPMP2_TAIL': '1'}, {'DGPMP2_TAIL': '7', 'DGPMP2_NP': '5'}):
    for k in ('DGPMP2_WIDE', 'DGPMP2_NP', 'DGPMP2_TAIL'):
        os.environ.pop(k, None)
    os.environ.update(env)
    for dof in (2, 3):
        base = XYH if dof == 3 else YAML
        for T in Ts:
            if dof == 3 and T > 200: continue
            for B in (1, 7, 301) if T <= 129 else (2,):
                pr = make_problems(B, T, dof=dof, unique_envs=4, seed=T * 7 + B, im_size=64)
                th, start, goal, sdf = (pr[k].cuda().double() for k in ('th_init', 'start', 'goal', 'sdf'))
                th = th + 0.05 * torch.randn_like(th)
                cp = cparams(T, base=base, dof=dof, non_holonomic=(dof == 3))
                dth, err, err_ext, status = ops.gn_step(cp, th, start, goal, sdf)
                assert int(status.abs().max()) == 0 and bool(torch.isfinite(dth).all()), (env, dof, T, B)
                D, U, r = ops.band(cp, th, start, goal, sdf)
                res = band_matvec(D, U, dth) - r
                lam = (D.reshape(B, -1).norm(dim=1) ** 2 + 2 * U.reshape(B, -1).norm(dim=1) ** 2).sqrt()
                eta = (res.reshape(B, -1).norm(dim=1) / (lam * dth.reshape(B, -1).norm(dim=1) + r.reshape(B, -1).norm(dim=1))).max().item()
                assert eta < 1e-13, (env, dof, T, B, eta)
                worst = max(worst, eta); n += 1
    print('schedule', env or 'default', 'ok; worst backward error so far %.2e over %d cases' % (worst, n), flush=True)
print('ALL OK')
