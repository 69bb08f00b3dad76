"""EDT from bit-packed maps (the e2e-from-occupancy path): device time per 1024 maps of 128^2."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from dgpmp2_b200 import ops
from dgpmp2_b200.datasets.synthetic import random_obstacle_map
B, N = 1024, 128
rng = np.random.default_rng(0)
uniq = [random_obstacle_map(rng, N, 'forest' if i % 2 == 0 else 'multi_obs') for i in range(64)]
ims = torch.from_numpy(np.stack([uniq[i % 64] for i in range(B)])).float()
bits = ops.pack_occupancy_bits(ims).cuda()
for _ in range(3): out = ops.sdf_from_occupancy_bits(bits, N, res=10.0 / N)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): out = ops.sdf_from_occupancy_bits(bits, N, res=10.0 / N)
e1.record(); torch.cuda.synchronize()
print('edt bits: %.3f ms per 1024 maps' % (e0.elapsed_time(e1) / 20))
