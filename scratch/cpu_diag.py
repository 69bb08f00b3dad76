import os, sys, time
sys.path.insert(0, '/root/repo')
import torch
print('cpu_count', os.cpu_count(), 'torch threads default', torch.get_num_threads(), 'interop', torch.get_num_interop_threads())
try:
    print('affinity', len(os.sched_getaffinity(0)))
except Exception as e:
    print('aff err', e)
import bench
pr = bench.make_inputs(0, 1, 128)[0]
for nt in [int(a) for a in sys.argv[1:]]:
    torch.set_num_threads(nt)
    step = bench.cpu_oracle_setup(128, pr)
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); step(); t2 = time.perf_counter()
    print('threads', nt, 'first %.2fs second %.2fs' % (t1 - t0, t2 - t1), flush=True)
