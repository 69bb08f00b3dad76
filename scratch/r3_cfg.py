"""L2-warm A/B of env-selected variants on one of bench.py's extra configs.  usage: r3_cfg.py <config key> [VAR=val,VAR=val ...]"""
import ctypes, os, sys, json
sys.path.insert(0, '/root/repo')
import torch
from dgpmp2_b200 import _lib, ops
import bench
name = sys.argv[1]
cfg = bench.EXTRA_CONFIGS[name]
dev = torch.device('cuda', 0)
Bc, Tc, dof = cfg['B'], cfg['T'], cfg['dof']
if len(sys.argv) > 2 and sys.argv[2].startswith('B='):
    Bc = int(sys.argv[2][2:]); del sys.argv[2]
dc = 2 * dof
prc = bench.make_inputs(100, 1, Bc, Tc, dof)[0]
cpc = bench.make_cparams(Bc, Tc, dof, cfg['base'], **cfg['flags'])
thc, stc, goc, sdfc = (prc[k].to(dev).contiguous() for k in ('th_init', 'start', 'goal', 'sdf'))
thc = ops.gn_solve(cpc, thc, stc, goc, sdfc, 5, 0.0)[0].contiguous()
_lib.set_sdf_shape(cpc, 128, 128, 128 * 128); cpc.B = Bc
dthc = torch.empty(Bc, Tc, dc, device=dev); errc = torch.empty(Bc, device=dev); eec = torch.empty(Bc, device=dev)
stc2, goc2, sdc2 = stc.reshape(Bc, dc).contiguous(), goc.reshape(Bc, dc).contiguous(), sdfc[:, 0].contiguous()
stat = torch.zeros(Bc, dtype=torch.int32, device=dev)
lib = _lib.load(); vp = ctypes.c_void_p
variants = [('default', {})] + [(a, dict(kv.split('=') for kv in a.split(','))) for a in sys.argv[2:]]
res, shapes, ref = {}, {}, None
for rep in range(3):
    for vname, env in variants:
        saved = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            gs = vp(torch.cuda.current_stream().cuda_stream)
            for i in range(50):
                assert lib.dgpmp2_gn_step_f32(ctypes.byref(cpc), vp(thc.data_ptr()), vp(stc2.data_ptr()), vp(goc2.data_ptr()), vp(sdc2.data_ptr()), None,
                                              vp(dthc.data_ptr()), vp(errc.data_ptr()), vp(eec.data_ptr()), vp(stat.data_ptr()), gs) == 0
        shapes[vname] = ops.launch_shape(cpc)
        for k, v in saved.items():
            if v is None: os.environ.pop(k, None)
            else: os.environ[k] = v
        g.replay(); torch.cuda.synchronize()
        if ref is None: ref = dthc.clone()
        same = bool(torch.equal(ref, dthc))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): g.replay()
        e1.record(); torch.cuda.synchronize()
        res.setdefault(vname, []).append(round(e0.elapsed_time(e1) / (20 * 50) * 1e3, 3))
        res[vname + ' same_bits'] = same
assert int(stat.abs().max()) == 0
print(json.dumps({'config': name, 'B': Bc, 'us_per_step': res, 'shapes': shapes}))
