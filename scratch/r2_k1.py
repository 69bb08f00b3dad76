"""Roofline probe of K1 = the fused SDF lookup + hinge + gradient in isolation (36 B per state):
dgpmp2_hinge_batch_f32 (positions in, cost + 2-wide gradient out) and, for comparison, the trajectory-in / d-wide-row-out
form (dgpmp2_factors_f32, obstacle outputs only)."""
import sys, json
sys.path.insert(0, '/root/repo')
import torch
from dgpmp2_b200 import ops
from dgpmp2_b200.datasets.synthetic import make_problems
from tests.gpu_helpers import cparams
B, T = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
pool = make_problems(256, T, unique_envs=256, seed=0)
dev = 'cuda'
idx = torch.arange(B, device=dev) % 256
sdf = pool['sdf'].to(dev)[idx].contiguous()            # (B,1,H,W): every problem has its own copy in HBM
start = (torch.rand(B, 1, 2, device=dev) * 8 - 4)
goal = (torch.rand(B, 1, 2, device=dev) * 8 - 4)
w = torch.linspace(0, 1, T, device=dev).reshape(1, T, 1)
pos = (start * (1 - w) + goal * w).contiguous()
th = torch.cat((pos, torch.zeros(B, T, 2, device=dev)), dim=2).contiguous()
cp = cparams(T)
peak = json.load(open('/root/repo/MEASURED_PEAKS.json'))['hbm_gbs']


def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

res = 10.0 / 128
ms_h = timed(lambda: ops.hinge_batch(sdf, pos, res, -5.0, -5.0, 0.4, eps_const=0.4))
ms_f = timed(lambda: ops.factors(cp, th, sdf, want_gp=False))
c1, h1 = ops.hinge_batch(sdf, pos, res, -5.0, -5.0, 0.4, eps_const=0.4)
_, c2, h2, _, _ = ops.factors(cp, th, sdf, want_gp=False)
assert torch.equal(c1, c2) and torch.equal(h1, h2[..., :2])
alg = B * T * 36
print(json.dumps({'B': B, 'T': T, 'states': B * T, 'sdf_GB': B * 128 * 128 * 4 / 1e9,
                  'hinge_batch': {'us': ms_h * 1e3, 'alg_GBs': alg / ms_h / 1e6, 'frac_alg': alg / ms_h / 1e6 / peak},
                  'factors(obstacle only)': {'us': ms_f * 1e3, 'alg_GBs': alg / ms_f / 1e6, 'frac_alg': alg / ms_f / 1e6 / peak}}))
