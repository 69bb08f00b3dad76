"""-m gpu: the opt-in mixed-precision GN step (DGPMP2_PRECISION=32; dgpmp2_b200/csrc/mp.cuh: fp32 node-owner block
cyclic reduction + fp64 residual refinement + in-launch fp64 fallback) against the live reference's goldens, the
all-double kernel and its own guard.  The CPU suite checks the same source through tests/host_emu."""
import os

import numpy as np
import pytest
import torch

from tests.helpers import XYH, YAML, load_golden, rel_err, step_cases

pytestmark = pytest.mark.gpu


@pytest.fixture
def mp_env():
    saved = {k: os.environ.pop(k, None) for k in ('DGPMP2_PRECISION', 'DGPMP2_MP_FORCE64', 'DGPMP2_MP_ACCEPT_LOG2')}
    os.environ['DGPMP2_PRECISION'] = '32'
    yield os.environ
    for k, v in saved.items():
        os.environ.pop(k, None)
        if v is not None:
            os.environ[k] = v


@pytest.mark.parametrize('name', step_cases())
def test_mp_step_vs_reference_golden(mp_env, name):
    from dgpmp2_b200 import ops
    from tests.gpu_helpers import cparams, dev
    g = load_golden(name)
    cp = cparams(g['T'], x_lims=g['x_lims'], y_lims=g['y_lims'], q_full=bool(g['q_full']))
    f32 = torch.float32
    th, start, goal, sdf = (dev(g[k], f32) for k in ('th', 'start', 'goal', 'sdf'))
    kw = {}
    if not bool(g['static']):
        kw = dict(qc_inv=dev(g['qc'], f32), w_obs=dev(g['w'], f32), eps=dev(g['eps'], f32))
    dth, err, err_ext, status, refine = ops.gn_step_diag(cp, th, start, goal, sdf, **kw)
    assert int(status.abs().max()) == 0
    assert int(refine.min()) >= 1 and int(refine.max()) <= 3, refine       # fp32 + refinement accepted, no fp64 fallback
    assert rel_err(dth.cpu(), g['dth']) < 1e-5
    np.testing.assert_allclose(err.cpu().double().numpy(), g['err'].reshape(-1), rtol=1e-6)
    np.testing.assert_allclose(err_ext.cpu().double().numpy(), g['err_ext'].reshape(-1), rtol=1e-6)


@pytest.mark.parametrize('dof,B,T', [(2, 1024, 64), (2, 300, 101), (2, 64, 128), (3, 64, 96), (2, 33, 7), (2, 40, 33)])
def test_mp_step_vs_all_double_kernel_and_forced_fallback(mp_env, dof, B, T):
    from dgpmp2_b200 import ops
    from dgpmp2_b200.datasets.synthetic import make_problems
    from tests.gpu_helpers import cparams
    base = XYH if dof == 3 else YAML
    pr = make_problems(B, T, dof=dof, im_size=64, seed=B + T, unique_envs=8)
    th, start, goal, sdf = (pr[k].cuda() for k in ('th_init', 'start', 'goal', 'sdf'))
    cp = cparams(T, base=base, dof=dof, non_holonomic=(dof == 3))
    th = ops.gn_solve(cp, th, start, goal, sdf, 3, 0.0)[0]
    mp = ops.gn_step_diag(cp, th, start, goal, sdf)
    assert int(mp[3].abs().max()) == 0 and int(mp[4].min()) >= 1
    ref64 = ops.gn_step(cp, th.double(), start.double(), goal.double(), sdf.double())        # float64 I/O: all-double
    assert rel_err(mp[0].cpu(), ref64[0].cpu()) < 1e-5
    torch.testing.assert_close(mp[1].double(), ref64[1], rtol=1e-6, atol=0)
    # position independence: a permuted batch gives the same bits per problem
    perm = torch.randperm(B, device='cuda', generator=torch.Generator(device='cuda').manual_seed(2))
    mp_p = ops.gn_step(cp, th[perm].contiguous(), start[perm].contiguous(), goal[perm].contiguous(), sdf[perm].contiguous())
    assert torch.equal(mp_p[0], mp[0][perm]) and torch.equal(mp_p[1], mp[1][perm])
    # every problem forced through the in-launch fp64 path == the all-double kernel, bit for bit
    mp_env['DGPMP2_MP_FORCE64'] = '1'
    forced = ops.gn_step_diag(cp, th, start, goal, sdf)
    mp_env.pop('DGPMP2_MP_FORCE64')
    mp_env['DGPMP2_PRECISION'] = '64'
    legacy = ops.gn_step_diag(cp, th, start, goal, sdf)
    assert int(legacy[4].abs().max()) == 0                                                     # 0 = all-double kernel
    for x, y in zip(forced[:4], legacy[:4]):
        assert torch.equal(x, y)


def test_mp_guard_hands_ill_conditioned_problems_to_fp64(mp_env):
    """reg = 0, weak priors, weak obstacle weight: cond(Lambda) ~ 1e12.  The fp32 factorisation cannot contract; the
    guard must route those problems to the fp64 path inside the launch and the results must equal the all-double kernel."""
    from dgpmp2_b200 import ops
    from dgpmp2_b200.datasets.synthetic import make_problems
    from tests.gpu_helpers import cparams
    T, B = 16, 12
    pr = make_problems(B, T, im_size=32, seed=5, unique_envs=2)
    base = dict(YAML, K_s=1e3, K_g=1e3, reg=0.0, cost_sigma=10.0)
    cp = cparams(T, base=base)
    th, start, goal, sdf = (pr[k].cuda() for k in ('th_init', 'start', 'goal', 'sdf'))
    mp = ops.gn_step_diag(cp, th, start, goal, sdf)
    mp_env['DGPMP2_PRECISION'] = '64'
    legacy = ops.gn_step(cp, th, start, goal, sdf)
    flagged = mp[4] < 0
    assert bool(flagged.any())
    assert torch.equal(mp[0][flagged], legacy[0][flagged]) and torch.equal(mp[3][flagged], legacy[3][flagged])
    if bool((~flagged).any()):
        assert rel_err(mp[0][~flagged].cpu(), legacy[0][~flagged].cpu()) < 1e-4
