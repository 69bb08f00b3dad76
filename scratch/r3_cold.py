"""DRAM-cold A/B of env-selected variants of the GN step (16 rotating input sets as bench.py, CUDA graph of 96 launches).
usage: r3_cold.py T B [gran=32|64|128] [VAR=val,VAR=val ...]   -- gran sets cudaLimitMaxL2FetchGranularity first."""
import ctypes, os, sys, json
sys.path.insert(0, '/root/repo')
import torch
from dgpmp2_b200 import _lib, ops
import bench
T, B, d = int(sys.argv[1]), int(sys.argv[2]), 4
args = sys.argv[3:]
dev = torch.device('cuda', 0)
torch.cuda.init(); torch.zeros(1, device=dev)
info = {}
if args and args[0].startswith('gran='):
    rt = ctypes.CDLL([l.split()[-1] for l in open('/proc/self/maps') if 'libcudart' in l][0])
    v = ctypes.c_size_t()
    rt.cudaDeviceGetLimit(ctypes.byref(v), 5); info['gran_before'] = v.value
    rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(args[0][5:])))
    rt.cudaDeviceGetLimit(ctypes.byref(v), 5); info['gran_after'] = v.value; info['rc'] = rc
    args = args[1:]
cp = bench.make_cparams(B=B, T=T)
sets = bench.make_inputs(0, 4, B, T=T)
dsets = []
for pr in sets:
    th, start, goal, sdf = (pr[k].to(dev).contiguous() for k in ('th_init', 'start', 'goal', 'sdf'))
    dsets.append([th, start.reshape(B, d).contiguous(), goal.reshape(B, d).contiguous(), sdf[:, 0].contiguous()])
gen = torch.Generator(device='cpu').manual_seed(1234)
for s in range(4, 16):
    a, b = dsets[s % 4], dsets[(s + 1 + s // 4) % 4]
    perm = torch.randperm(B, generator=gen).to(dev)
    dsets.append([a[0], a[1], a[2], b[3][perm].contiguous()])
for ds in dsets:
    ds[0] = ops.gn_solve(cp, ds[0], ds[1].reshape(B, 1, d), ds[2].reshape(B, 1, d), ds[3].unsqueeze(1), 5, 0.0)[0].contiguous()
_lib.set_sdf_shape(cp, 128, 128, 128 * 128); cp.B = B
dth = torch.empty(B, T, d, device=dev); err = torch.empty(B, device=dev); ee = torch.empty(B, device=dev)
st = torch.zeros(B, dtype=torch.int32, device=dev)
lib = _lib.load(); vp = ctypes.c_void_p
variants = [('default', {})] + [(a, dict(kv.split('=') for kv in a.split(','))) for a in args]
res = {}
for rep in range(3):
    for name, env in variants:
        saved = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        for nsets, tag in ((16, 'cold'), (1, 'warm')):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                gs = vp(torch.cuda.current_stream().cuda_stream)
                for i in range(96):
                    a = dsets[i % nsets]
                    assert lib.dgpmp2_gn_step_f32(ctypes.byref(cp), vp(a[0].data_ptr()), vp(a[1].data_ptr()), vp(a[2].data_ptr()), vp(a[3].data_ptr()), None,
                                                  vp(dth.data_ptr()), vp(err.data_ptr()), vp(ee.data_ptr()), vp(st.data_ptr()), gs) == 0
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): g.replay()
            e1.record(); torch.cuda.synchronize()
            res.setdefault(name + ' ' + tag, []).append(round(e0.elapsed_time(e1) / (20 * 96) * 1e3, 3))
        for k, v in saved.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v
assert int(st.abs().max()) == 0
print(json.dumps({'T': T, 'B': B, 'info': info, 'us_per_step': res}))
