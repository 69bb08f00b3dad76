# GPU SDF generation timing: python scratch/edt_time.py [B] [size]   (CUDA events, 20 launches after warm-up)
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
import torch
import os
from dgpmp2_b200 import _lib
if os.environ.get('DGPMP2_LIB'): _lib.LIB_PATH = os.environ['DGPMP2_LIB']
from dgpmp2_b200 import ops
from dgpmp2_b200.datasets.synthetic import random_obstacle_map
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 128
rng = np.random.default_rng(0)
uniq = [random_obstacle_map(rng, N, 'forest' if i % 2 == 0 else 'multi_obs') for i in range(64)]
ims = torch.from_numpy(np.stack([uniq[i % 64] for i in range(B)])).float().cuda()
for pad in (0, 1):
    for _ in range(3):
        out = ops.sdf_from_occupancy(ims, padlen=pad, res=10.0 / N)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        out = ops.sdf_from_occupancy(ims, padlen=pad, res=10.0 / N)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print('sdf_from_occupancy B=%d %dx%d pad=%d: %.3f ms per batch, %.2f us per map, %.1f GB/s (im in + sdf out, fp32)' % (
        B, N, N, pad, ms, ms * 1e3 / B, (ims.numel() * 4 + out.numel() * 4) / ms / 1e6))
