"""-m gpu: gn_step can be launched with programmatic dependent launch (DGPMP2_PDL=2; c_abi.cu launch_step; opt-in,
see DESIGN.md 4.1): the next step may be scheduled before the previous grid has drained, but every global access
of the kernel comes after griddepcontrol.wait.  Dependent chains of launches (th <- th + dth with a plain torch kernel in between, and
launches that overwrite one output buffer) must therefore give the bits of plain launches (DGPMP2_PDL=1), eagerly
and when the chain is captured in a CUDA graph (how bench.py times the step)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _chain(B, T, n, graph):
    from dgpmp2_b200 import ops
    from dgpmp2_b200.datasets.synthetic import make_problems
    from tests.gpu_helpers import cparams
    pr = make_problems(B, T, im_size=64, seed=17, unique_envs=8)
    th, start, goal, sdf = (pr[k].cuda().contiguous() for k in ('th_init', 'start', 'goal', 'sdf'))
    cp = cparams(T)
    dth = torch.empty_like(th)
    errs = []

    def body():
        for _ in range(n):
            _, err, _, _ = ops.gn_step(cp, th, start, goal, sdf, out=dth, want_status=False)   # same output buffer every step
            th.add_(dth)                                                                      # plain kernel in between
            errs.append(err)
    if graph:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            ops.gn_step(cp, th.clone(), start, goal, sdf)          # shared-memory opt-in happens outside the capture
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                body()
            g.replay()
            s.synchronize()
    else:
        body()
        torch.cuda.synchronize()
    return th.clone(), torch.stack(errs).clone()


@pytest.mark.parametrize('graph', [False, True], ids=['eager', 'cuda-graph'])
@pytest.mark.parametrize('B,T', [(5, 16), (300, 64)])
def test_programmatic_dependent_launch_preserves_stream_order(B, T, graph):
    saved = os.environ.pop('DGPMP2_PDL', None)
    try:
        os.environ['DGPMP2_PDL'] = '1'
        th_plain, err_plain = _chain(B, T, 12, graph)
        os.environ['DGPMP2_PDL'] = '2'
        th_pdl, err_pdl = _chain(B, T, 12, graph)
        os.environ.pop('DGPMP2_PDL')
        th_dflt, err_dflt = _chain(B, T, 12, graph)
    finally:
        os.environ.pop('DGPMP2_PDL', None)
        if saved is not None:
            os.environ['DGPMP2_PDL'] = saved
    assert torch.equal(th_plain, th_pdl) and torch.equal(err_plain, err_pdl)
    assert torch.equal(th_plain, th_dflt) and torch.equal(err_plain, err_dflt)
    assert bool(torch.isfinite(th_plain).all()) and float(err_plain[-1].mean()) < float(err_plain[0].mean())
