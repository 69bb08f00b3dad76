import sys, os
sys.path.insert(0, '/root/repo')
import torch
from dgpmp2_b200 import ops
from dgpmp2_b200.datasets.synthetic import make_problems
from tests.gpu_helpers import cparams
B, T = int(sys.argv[1]), int(sys.argv[2])
pr = make_problems(B, T, unique_envs=64, seed=0)
th, start, goal, sdf = (pr[k].cuda().double() for k in ('th_init', 'start', 'goal', 'sdf'))
cp = cparams(T)
res = {}
for w in (1, 64, 100000):
    os.environ['DGPMP2_WIDE'] = str(w)
    res[w] = ops.gn_step(cp, th, start, goal, sdf)
    r2 = ops.gn_step(cp, th, start, goal, sdf)
    print('WIDE', w, 'repeatable', torch.equal(res[w][0], r2[0]), 'status', int(res[w][3].abs().max()))
for w in (1, 64):
    d = (res[w][0] - res[100000][0]).abs()
    print('WIDE', w, 'vs LPN4: max abs diff', d.max().item(), 'rel', (d.max() / res[100000][0].abs().max()).item(), 'n diff', int((d > 0).sum()), 'err equal', torch.equal(res[w][1], res[100000][1]))
    bt = (d.reshape(B, -1).max(dim=1).values > 0).nonzero().flatten()
    print('  problems differing', bt.numel(), bt[:10].tolist())
    if bt.numel():
        b = int(bt[0]); print('  first: t with diff', (d[b].max(dim=1).values > 0).nonzero().flatten().tolist()[:70])
