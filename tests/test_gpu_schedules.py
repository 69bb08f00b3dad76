"""-m gpu: the solver across trajectory lengths, batch sizes and elimination schedules.

Every problem's dtheta must satisfy its own normal equations (band kernel, fp64) to rounding, for ragged T
(non powers of two, shorter than the tail, longer than a level), 1 / several / many problems per CTA, and with
the schedule knobs forced so that the one-lane levels, the 4-lane levels, multi-problem item packing and tails of
length 1..7 are all exercised on small problems.  The knobs are the library's documented debug overrides
(DGPMP2_WIDE / DGPMP2_NP / DGPMP2_TAIL, read per call in c_abi.cu)."""
import os

import pytest
import torch

from tests.helpers import XYH, YAML

pytestmark = pytest.mark.gpu

SCHEDULES = [{}, {'DGPMP2_WIDE': '4'}, {'DGPMP2_WIDE': '4', 'DGPMP2_NP': '3'}, {'DGPMP2_TAIL': '1'},
             {'DGPMP2_TAIL': '7', 'DGPMP2_NP': '5'}]
KNOBS = ('DGPMP2_WIDE', 'DGPMP2_NP', 'DGPMP2_TAIL')


def _band_matvec(D, U, x):
    y = torch.einsum('btij,btj->bti', D, x)
    y[:, :-1] += torch.einsum('btij,btj->bti', U, x[:, 1:])
    y[:, 1:] += torch.einsum('btji,btj->bti', U, x[:, :-1])
    return y


@pytest.mark.parametrize('dof', [2, 3])
@pytest.mark.parametrize('sched', SCHEDULES, ids=lambda s: '-'.join('%s%s' % (k[7:], v) for k, v in s.items()) or 'default')
def test_every_problem_solves_its_normal_equations(sched, dof):
    from dgpmp2_b200 import ops
    from dgpmp2_b200.datasets.synthetic import make_problems
    from tests.gpu_helpers import cparams
    saved = {k: os.environ.pop(k, None) for k in KNOBS}
    os.environ.update(sched)
    try:
        base = XYH if dof == 3 else YAML
        ref = {}
        for T in (2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 33, 64, 65, 101, 128):
            for B in (1, 7, 150):
                pr = make_problems(B, T, dof=dof, unique_envs=3, seed=T * 7 + B, im_size=48)
                th, start, goal, sdf = (pr[k].cuda().double() for k in ('th_init', 'start', 'goal', 'sdf'))
                th = th + 0.05 * torch.randn(th.shape, generator=torch.Generator().manual_seed(T)).cuda().double()
                cp = cparams(T, base=base, dof=dof, non_holonomic=(dof == 3))
                dth, err, err_ext, status = ops.gn_step(cp, th, start, goal, sdf)
                assert int(status.abs().max()) == 0 and bool(torch.isfinite(dth).all()), (T, B)
                D, U, r = ops.band(cp, th, start, goal, sdf)
                res = _band_matvec(D, U, dth) - r
                lam = (D.reshape(B, -1).norm(dim=1) ** 2 + 2 * U.reshape(B, -1).norm(dim=1) ** 2).sqrt()
                eta = (res.reshape(B, -1).norm(dim=1) /
                       (lam * dth.reshape(B, -1).norm(dim=1) + r.reshape(B, -1).norm(dim=1))).max().item()
                assert eta < 1e-13, (T, B, eta)      # normwise backward error: independent of cond(Lambda)
                ref[(T, B)] = dth
        if sched.get('DGPMP2_WIDE') == '4' and 'DGPMP2_NP' not in sched:
            # one lane per item vs four lanes per item: same bits (explicit-fma arithmetic)
            os.environ['DGPMP2_WIDE'] = '100000'
            for (T, B), d0 in ref.items():
                if B != 7 or T not in (16, 64, 101):
                    continue
                pr = make_problems(B, T, dof=dof, unique_envs=3, seed=T * 7 + B, im_size=48)
                th, start, goal, sdf = (pr[k].cuda().double() for k in ('th_init', 'start', 'goal', 'sdf'))
                th = th + 0.05 * torch.randn(th.shape, generator=torch.Generator().manual_seed(T)).cuda().double()
                cp = cparams(T, base=base, dof=dof, non_holonomic=(dof == 3))
                assert torch.equal(ops.gn_step(cp, th, start, goal, sdf)[0], d0), (T, B)
    finally:
        for k in KNOBS:
            os.environ.pop(k, None)
            if saved[k] is not None:
                os.environ[k] = saved[k]


@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
def test_level1_elimination_fused_into_the_assembly_is_bit_identical(dtype):
    """Static-GP launches eliminate the level-1 nodes inside the assembly (kernels.cuh: assemble_cta, fuse1): the same
    arithmetic on the same doubles as the separate elimination phase (DGPMP2_FUSE1=2) -> identical bits."""
    from dgpmp2_b200 import ops
    from dgpmp2_b200.datasets.synthetic import make_problems
    from tests.gpu_helpers import cparams
    saved = os.environ.pop('DGPMP2_FUSE1', None)
    try:
        for dof, B, T in ((2, 1, 64), (2, 7, 64), (2, 1024, 64), (2, 300, 101), (2, 512, 128), (2, 5, 3), (2, 9, 2),
                          (3, 64, 96), (3, 3, 12)):
            base = XYH if dof == 3 else YAML
            pr = make_problems(B, T, dof=dof, im_size=64, seed=B + T, unique_envs=8, dtype=dtype)
            th, start, goal, sdf = (pr[k].cuda() for k in ('th_init', 'start', 'goal', 'sdf'))
            cp = cparams(T, base=base, dof=dof, non_holonomic=(dof == 3))
            th = ops.gn_solve(cp, th, start, goal, sdf, 3, 0.0)[0]
            os.environ.pop('DGPMP2_FUSE1', None)
            a = ops.gn_step(cp, th, start, goal, sdf)
            os.environ['DGPMP2_FUSE1'] = '2'
            b = ops.gn_step(cp, th, start, goal, sdf)
            for x, y in zip(a, b):
                assert torch.equal(x, y), (dof, B, T)
    finally:
        os.environ.pop('DGPMP2_FUSE1', None)
        if saved is not None:
            os.environ['DGPMP2_FUSE1'] = saved


@pytest.mark.parametrize('dof,T,B', [(2, 128, 1000), (2, 128, 593), (3, 96, 300), (2, 64, 1333)])
def test_balanced_waves_give_the_same_bits_and_cover_every_problem(dof, T, B):
    """When an SM's share of the batch does not fit one CTA, gn_step_kernel runs CTAs of np and np - 1 problems
    (c_abi.cu: choose_shape, balanced).  A problem's result does not depend on the CTA it lands in: identical bits to
    the uniform grid (DGPMP2_BALANCED=2), every problem written exactly once (NaN-filled outputs, ragged last CTA)."""
    from dgpmp2_b200 import ops
    from dgpmp2_b200.datasets.synthetic import make_problems
    from tests.gpu_helpers import cparams
    base = XYH if dof == 3 else YAML
    pr = make_problems(B, T, dof=dof, unique_envs=5, seed=B + T, im_size=48)
    th, start, goal, sdf = (pr[k].cuda().float() for k in ('th_init', 'start', 'goal', 'sdf'))
    cp = cparams(T, base=base, dof=dof, non_holonomic=(dof == 3))
    saved = os.environ.pop('DGPMP2_BALANCED', None)
    try:
        shape_bal = ops.launch_shape(cparams(T, B=B, base=base, dof=dof, non_holonomic=(dof == 3)))
        got = ops.gn_step(cp, th, start, goal, sdf)
        os.environ['DGPMP2_BALANCED'] = '2'
        shape_uni = ops.launch_shape(cparams(T, B=B, base=base, dof=dof, non_holonomic=(dof == 3)))
        ref = ops.gn_step(cp, th, start, goal, sdf)
    finally:
        os.environ.pop('DGPMP2_BALANCED', None)
        if saved is not None:
            os.environ['DGPMP2_BALANCED'] = saved
    assert int(ref[3].abs().max()) == 0 and bool(torch.isfinite(ref[0]).all())
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    print(shape_bal, shape_uni)
