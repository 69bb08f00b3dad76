import sys, ctypes, os
sys.path.insert(0, '/root/repo')
import torch
from dgpmp2_b200 import _lib
_lib.LIB_PATH = os.environ.get('DGPMP2_LIB', '/root/repo/scratch/libdgpmp2_timing.so')
from dgpmp2_b200 import ops
from dgpmp2_b200.datasets.synthetic import make_problems
from tests.gpu_helpers import cparams
B, T = int(sys.argv[1]), int(sys.argv[2])
pr = make_problems(B, T, unique_envs=64, seed=0)
th, start, goal, sdf = (pr[k].cuda() for k in ('th_init', 'start', 'goal', 'sdf'))
cp = cparams(T)
for _ in range(5):
    ops.gn_step(cp, th, start, goal, sdf)
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_longlong * 64)()
lib.dgpmp2_debug_phase_clocks.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
lib.dgpmp2_debug_phase_clocks(buf)
c = list(buf)
t0 = c[0]
print('prologue', c[1]-c[0], 'assembly', c[2]-c[1], 'bcr', c[3]-c[2], 'epilogue', c[4]-c[3], 'total', c[4]-c[0])
prev = c[2]
for l in range(1, 8):
    a, b = c[8+2*l], c[9+2*l]
    if a == 0: break
    print('level', l, 'elim', a-prev, 'kept', b-a)
    prev = b
print('root', c[5]-prev, '(tail fwd', c[6]-prev, 'back', c[5]-c[6], ')')
prev = c[5]
for l in range(7, 0, -1):
    if c[40+l] == 0: continue
    print('back level', l, c[40+l]-prev); prev = c[40+l]
