"""Roofline probe of the stand-alone fused SDF-lookup + hinge + obstacle-Jacobian kernel
(dgpmp2_factors_f32 with only the obstacle outputs) at streaming sizes."""
import sys, json
sys.path.insert(0, '/root/repo')
import torch
from dgpmp2_b200 import ops, _lib
from dgpmp2_b200.datasets.synthetic import make_problems
from tests.gpu_helpers import cparams
B, T = int(sys.argv[1]), int(sys.argv[2])
pool = make_problems(256, T, unique_envs=256, seed=0)
dev = 'cuda'
idx = torch.arange(B, device=dev) % 256
sdf = pool['sdf'].to(dev)[idx].contiguous()            # (B,1,H,W): every problem has its own copy in HBM
start = (torch.rand(B, 1, 2, device=dev) * 8 - 4)
goal = (torch.rand(B, 1, 2, device=dev) * 8 - 4)
w = torch.linspace(0, 1, T, device=dev).reshape(1, T, 1)
pos = start * (1 - w) + goal * w
th = torch.cat((pos, torch.zeros(B, T, 2, device=dev)), dim=2).contiguous()
cp = cparams(T)
for _ in range(3):
    out = ops.factors(cp, th, sdf, want_gp=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 20
e0.record()
for _ in range(N):
    out = ops.factors(cp, th, sdf, want_gp=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / N
alg = B * T * 36            # positions in (8) + 4 taps (16) + cost and gradient out (12)
moved = B * T * (16 + 16 + 4 + 16)   # what this kernel actually reads/writes per state: th (16) + taps (16) + cost (4) + H (16)
peak = json.load(open('/root/repo/MEASURED_PEAKS.json'))['hbm_gbs']
print(json.dumps({'B': B, 'T': T, 'states': B * T, 'us': ms * 1e3, 'alg_GBs': alg / ms / 1e6, 'frac_alg': alg / ms / 1e6 / peak,
                  'moved_GBs': moved / ms / 1e6, 'frac_moved': moved / ms / 1e6 / peak, 'sdf_GB': B * 128 * 128 * 4 / 1e9}))
