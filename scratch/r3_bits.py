"""Bitwise A/B of two builds of the library on the bench workload and on the extra configs: run once per library
(DGPMP2_LIB), dump dth / err / err_ext; then compare the dumps.  usage: r3_bits.py dump <file> | r3_bits.py cmp <a> <b>"""
import sys, json
sys.path.insert(0, '/root/repo')
import torch
if sys.argv[1] == 'cmp':
    a, b = torch.load(sys.argv[2]), torch.load(sys.argv[3])
    ok = True
    for k in a:
        for i, (x, y) in enumerate(zip(a[k], b[k])):
            same = bool(torch.equal(x, y))
            ok &= same
            if not same:
                print(k, i, 'DIFF max abs', float((x.double() - y.double()).abs().max()), 'rel', float(((x.double() - y.double()).abs().max()) / x.double().abs().max()))
    print('bitwise identical' if ok else 'NOT identical')
    sys.exit(0)
from dgpmp2_b200 import _lib, ops
import bench
out = {}
dev = torch.device('cuda', 0)
for dt in (torch.float32, torch.float64):
    pr = bench.make_inputs(0, 1, 1024)[0]
    cp = bench.make_cparams()
    a = [pr[k].to(dev).to(dt).contiguous() for k in ('th_init', 'start', 'goal', 'sdf')]
    th = ops.gn_solve(cp, *a, 5, 0.0)[0].contiguous()
    out['cfg2 %s' % dt] = [t.cpu() for t in ops.gn_step(cp, th, a[1], a[2], a[3])[:3]] + [th.cpu()]
    for name, cfg in bench.EXTRA_CONFIGS.items():
        prc = bench.make_inputs(100, 1, 300, cfg['T'], cfg['dof'])[0]
        cpc = bench.make_cparams(300, cfg['T'], cfg['dof'], cfg['base'], **cfg['flags'])
        a = [prc[k].to(dev).to(dt).contiguous() for k in ('th_init', 'start', 'goal', 'sdf')]
        th = ops.gn_solve(cpc, *a, 3, 0.0)[0].contiguous()
        out['%s %s' % (name, dt)] = [t.cpu() for t in ops.gn_step(cpc, th, a[1], a[2], a[3])[:3]] + [th.cpu()]
        g = ops.gn_step_backward(cpc, th, a[1], a[2], a[3], out['%s %s' % (name, dt)][0].to(dev), torch.ones_like(th), None) if hasattr(ops, 'gn_step_backward') else None
torch.save(out, sys.argv[2])
print('dumped', len(out))
