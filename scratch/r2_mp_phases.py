"""Per-phase clock64() stamps of the mixed-precision step (timing build: -DDGPMP2_MP_TIMING=cta+1)."""
import sys, ctypes, os
sys.path.insert(0, '/root/repo')
import torch
from dgpmp2_b200 import _lib
_lib.LIB_PATH = os.environ.get('DGPMP2_LIB', '/root/repo/scratch/exp/libtiming.so')
from dgpmp2_b200 import ops
from dgpmp2_b200.datasets.synthetic import make_problems
from tests.gpu_helpers import cparams
B, T = int(sys.argv[1]), int(sys.argv[2])
pr = make_problems(B, T, unique_envs=64, seed=0)
th, start, goal, sdf = (pr[k].cuda() for k in ('th_init', 'start', 'goal', 'sdf'))
cp = cparams(T)
th = ops.gn_solve(cp, th, start, goal, sdf, 5, 0.0)[0]
for _ in range(5):
    ops.gn_step(cp, th, start, goal, sdf)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * 64)()
lib.dgpmp2_debug_mp_clocks.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
lib.dgpmp2_debug_mp_clocks(buf)
c = list(buf)
names = {0: 'start', 1: 'assembled', 2: 'err reduced', 12: 'factor loop done', 13: 'root done', 14: 'backsub done', 40: 'end'}
for l in range(1, 10): names[2 + l] = 'level %d eliminated+synced' % l
for it in range(1, 6):
    names[11 + 4 * it] = 'it%d residual+sync' % it
    names[12 + 4 * it] = 'it%d fwd sweep+root' % it
    names[13 + 4 * it] = 'it%d backsub' % it
    names[14 + 4 * it] = 'it%d norms' % it
ev = sorted((v, k) for k, v in enumerate(c) if v > 0)
t0 = ev[0][0]; prev = t0
print('B=%d T=%d' % (B, T))
for v, k in ev:
    print('%6d  +%5d  %s' % (v - t0, v - prev, names.get(k, str(k)))); prev = v
