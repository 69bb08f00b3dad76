"""Backward of the GN step: the oracle's autograd is pinned against gradients taken through the LIVE
reference's own autograd graph (tests/golden/grad_B2_T16.npz); the CUDA backward kernel is checked against
both (fp64 I/O, rel 1e-7; tolerance reflects two chained solves with cond(Lambda) up to ~1e5)."""
import numpy as np
import pytest
import torch

from oracle import gn_oracle
from tests.helpers import load_golden, oracle_params, t64, XYH, YAML

NAMES = ['th', 'start', 'goal', 'sdf', 'qc', 'w', 'eps']
OUT_ORDER = ['th', 'start', 'goal', 'qc', 'w', 'eps', 'sdf']      # ops.gn_step_backward return order


def _oracle_grads(g, p, q_full=False, leaves_in=None):
    leaves = leaves_in or [t64(g[n]).requires_grad_(True) for n in NAMES]
    dth, err, err_ext = gn_oracle.gn_step(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], leaves[5], leaves[6], p, q_full)
    loss = (dth * t64(g['G'])).sum() + (err_ext * t64(g['g_err_ext'])).sum()
    return torch.autograd.grad(loss, leaves, allow_unused=True)


def _close(a, b, rtol, name):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(np.abs(b).max(), 1e-300)
    assert np.abs(a - b).max() <= rtol * scale, '%s: max abs diff %.3e vs scale %.3e' % (name, np.abs(a - b).max(), scale)


def test_oracle_autograd_matches_reference_autograd():
    g = load_golden('grad_B2_T16')
    assert not bool(g['err_requires_grad'])          # err is computed under no_grad in the reference
    p = oracle_params(g['T'])
    grads = _oracle_grads(g, p)
    for n, gr in zip(NAMES, grads):
        _close(gr.numpy(), g['g_' + n], 1e-8, n)


@pytest.mark.gpu
def test_cuda_backward_matches_reference_autograd():
    from dgpmp2_b200 import ops
    from tests.gpu_helpers import cparams, dev
    g = load_golden('grad_B2_T16')
    dt = torch.float64
    cp = cparams(g['T'])
    th, start, goal, sdf, qc, w, eps = (dev(g[n], dt) for n in NAMES)
    dth, err, err_ext, status = ops.gn_step(cp, th, start, goal, sdf, qc_inv=qc, w_obs=w, eps=eps)
    _close(dth.cpu().numpy(), g['dth'], 1e-9, 'dth')
    outs = ops.gn_step_backward(cp, th, start, goal, sdf, dth, dev(g['G'], dt), dev(g['g_err_ext'], dt).reshape(-1),
                                qc_inv=qc, w_obs=w, eps=eps, need_th=True, need_start=True, need_goal=True,
                                need_qc=True, need_w=True, need_eps=True, need_sdf=True)
    outs = dict(zip(OUT_ORDER, outs))
    for n in NAMES:
        _close(outs[n].cpu().numpy().reshape(g['g_' + n].shape), g['g_' + n], 1e-7, n)


@pytest.mark.gpu
def test_autograd_function_through_the_planner_api():
    """planner.plan_layer(...) is differentiable end to end (CPU leaves, CUDA compute), like the reference."""
    from tests.test_gpu_api import _planner
    g = load_golden('grad_B2_T16')
    planner = _planner(int(g['T']), 2)
    leaves = [t64(g[n]).requires_grad_(True) for n in NAMES]
    dth, err, err_ext = planner.plan_layer(leaves[0], leaves[1], leaves[2], None, leaves[3], leaves[4], leaves[5], leaves[6])
    assert dth.requires_grad and err_ext.requires_grad and not err.requires_grad
    loss = (dth * t64(g['G'])).sum() + (err_ext * t64(g['g_err_ext'])).sum()
    loss.backward()
    for n, l in zip(NAMES, leaves):
        _close(l.grad.numpy(), g['g_' + n], 1e-7, n)


@pytest.mark.gpu
@pytest.mark.parametrize('case', ['static', 'q_full', 'nonholonomic', 'vel_limits'])
def test_cuda_backward_vs_oracle_autograd_seeded(case):
    from dgpmp2_b200 import ops
    from tests.gpu_helpers import cparams, dev
    rng = np.random.default_rng(5)
    B, T, H, W = 3, 12, 24, 24
    dof = 3 if case == 'nonholonomic' else 2
    d = 2 * dof
    f32 = lambda a: torch.as_tensor(a).float().double()
    th = f32(rng.uniform(-4.5, 4.5, (B, T, d)))
    if case == 'vel_limits':
        th[:, :, 2:] = f32(rng.uniform(-2.5, 2.5, (B, T, 2)))
    start, goal = f32(rng.uniform(-4, 4, (B, 1, d))), f32(rng.uniform(-4, 4, (B, 1, d)))
    sdf = f32(rng.uniform(-1.0, 3.0, (B, 1, H, W)))
    blk = d if case == 'q_full' else dof
    q = rng.standard_normal((B, T - 1, blk, 1))
    qc = f32(q @ q.transpose(0, 1, 3, 2) + np.eye(blk) * rng.uniform(0.3, 2.0, (B, T - 1, 1, 1)))
    w = f32(rng.uniform(10.0, 2e4, (B, T, 1, 1)))
    eps = f32(rng.uniform(0.0, 1.0, (B, T, 1, 1)))
    base = XYH if dof == 3 else dict(YAML, K_v=0.01, v_x=1.0, v_y=1.0)
    over = dict(non_holonomic=(case == 'nonholonomic'), use_vel_limits=(case == 'vel_limits'))
    p = oracle_params(T, base=base, dof=dof, **({'K_v': 0.01, 'v_x': 1.0, 'v_y': 1.0} if case == 'vel_limits' else {}), **over)
    g = {'G': rng.standard_normal((B, T, d)), 'g_err_ext': rng.standard_normal((B, 1, 1))}
    leaves = [x.clone().requires_grad_(True) for x in (th, start, goal, sdf, qc, w, eps)]
    ref = _oracle_grads(g, p, q_full=(case == 'q_full'), leaves_in=leaves)
    cp = cparams(T, base=base, dof=dof, q_full=(case == 'q_full'), **over)
    dt = torch.float64
    kw = dict(qc_inv=dev(qc, dt), w_obs=dev(w, dt), eps=dev(eps, dt))
    args = [dev(x, dt) for x in (th, start, goal, sdf)]
    dth = ops.gn_step(cp, *args, **kw)[0]
    outs = ops.gn_step_backward(cp, *args, dth, dev(g['G'], dt), dev(g['g_err_ext'], dt).reshape(-1), need_th=True,
                                need_start=True, need_goal=True, need_qc=True, need_w=True, need_eps=True, need_sdf=True, **kw)
    outs = dict(zip(OUT_ORDER, outs))
    for n, r in zip(NAMES, ref):
        _close(outs[n].cpu().numpy().reshape(tuple(r.shape)), r.numpy(), 1e-7, case + ':' + n)


@pytest.mark.gpu
def test_forward_is_differentiable_like_the_reference_example():
    """examples/diff_gpmp2_2d_example.py:75-78: th_final.backward(...) through the unrolled iterations."""
    from tests.test_gpu_api import _planner
    g = load_golden('forward_B3_T32')
    planner = _planner(32, 3, max_iters=4, tol_delta=1e-9)
    th0 = t64(g['th']).requires_grad_(True)
    start, goal, sdf = t64(g['start']), t64(g['goal']), t64(g['sdf'])
    th_final, _, _, _, _, _, k, _ = planner.forward(th0, start, goal, None, sdf)
    assert k == [4, 4, 4] and th_final.requires_grad
    R = torch.randn(th_final.shape, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    th_final.backward(R)
    # oracle: the same 4 unrolled GN iterations under autograd
    p = oracle_params(32)
    th = t64(g['th']).requires_grad_(True)
    qc = torch.tensor(p.Q_c_inv, dtype=torch.float64).expand(3, 31, 2, 2)
    w = torch.full((3, 32, 1, 1), 1.0 / p.cost_sigma ** 2, dtype=torch.float64)
    eps = torch.full((3, 32, 1, 1), p.epsilon_dist, dtype=torch.float64)
    x = th
    for _ in range(4):
        x = x + gn_oracle.gn_step(x, start, goal, sdf, qc, w, eps, p)[0]
    (x * R).sum().backward()
    _close(th0.grad.numpy(), th.grad.numpy(), 1e-6, 'd th_final / d th_init')


def test_oracle_error_gradients_match_reference_autograd():
    """d(err_ext, err_sg, err_gp, err_obs)/d th of the oracle's restatement vs the LIVE reference's autograd through
    error_ext_batch / unweighted_errors_batch (tests/golden/errgrad_B2_T16.npz, oracle/make_golden_r2.py)."""
    g = load_golden('errgrad_B2_T16')
    assert not bool(g['err_requires_grad'])
    p = oracle_params(g['T'])
    th = t64(g['th']).requires_grad_(True)
    start, goal, sdf, eps = (t64(g[k]) for k in ('start', 'goal', 'sdf', 'eps'))
    B, T = th.shape[0], int(g['T'])
    q_fix, w_fix = gn_oracle.fixed_covariances(p, B)
    e_ext = gn_oracle.weighted_error(th, start, goal, sdf, q_fix, w_fix, eps, p)
    e_sg, e_gp, e_obs = gn_oracle.unweighted_errors(th, start, goal, sdf, eps, p)
    for k, e in dict(ext=e_ext, sg=e_sg, gp=e_gp, obs=e_obs).items():
        np.testing.assert_allclose(e.detach().numpy().reshape(g['err_' + k].shape), g['err_' + k], rtol=1e-12)
        gr, = torch.autograd.grad((e.reshape(g['c_' + k].shape) * t64(g['c_' + k])).sum(), th, retain_graph=True)
        _close(gr.numpy(), g['g_th_' + k], 1e-10, k)


@pytest.mark.gpu
@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
def test_cuda_error_gradients_match_reference_autograd(dtype):
    """error_ext_batch / gp_error / obs_error / start_goal_error / unweighted_errors_batch are differentiable w.r.t. the
    trajectory, as in the reference (its training loss is built from them: learning/train_planner.py:327-346)."""
    from tests.test_gpu_api import _planner
    g = load_golden('errgrad_B2_T16')
    planner = _planner(int(g['T']), 2)
    pl = planner.plan_layer
    cast = lambda k: torch.as_tensor(g[k]).to(dtype)
    with torch.no_grad():
        pl(cast('th'), cast('start'), cast('goal'), None, cast('sdf'), cast('qc'), cast('w'), cast('eps'))
    tol = 1e-9 if dtype == torch.float64 else 2e-5
    for k in ('ext', 'sg', 'gp', 'obs'):
        leaf = cast('th').requires_grad_(True)
        if k == 'ext':
            e = pl.error_ext_batch(leaf, cast('sdf'))
        else:
            e = dict(zip(('sg', 'gp', 'obs'), planner.unweighted_errors_batch(leaf, cast('sdf'))))[k]
        assert e.requires_grad and tuple(e.shape) == g['err_' + k].shape
        np.testing.assert_allclose(e.detach().double().numpy(), g['err_' + k], rtol=1e-11 if dtype == torch.float64 else 1e-6)
        (e * torch.as_tensor(g['c_' + k]).to(dtype)).sum().backward()
        _close(leaf.grad.double().numpy(), g['g_th_' + k], tol, k)
    leaf = cast('th').requires_grad_(True)
    assert not pl.error_batch(leaf, cast('sdf')).requires_grad            # computed under no_grad in the reference (:275)
    # the single-quantity accessors are differentiable too
    e = pl.gp_error(leaf).sum() + pl.start_goal_error(leaf).sum() + pl.obs_error(leaf, cast('sdf')).sum()
    e.backward()
    want = g['g_th_gp'] / g['c_gp'].reshape(-1, 1, 1) + g['g_th_sg'] / g['c_sg'].reshape(-1, 1, 1) + g['g_th_obs'] / g['c_obs'].reshape(-1, 1, 1)
    _close(leaf.grad.double().numpy(), want, tol * 10, 'sum of accessors')


@pytest.mark.gpu
def test_broadcast_weights_receive_summed_gradients():
    """Expanded / broadcast-shaped weights ((1,T-1,dof,dof), (1,T,1,1), (B,1,1,1)) are accepted by forward without
    copies (make_weights, stride 0); backward must hand back their own shape, summed over the broadcast dimensions."""
    from tests.test_gpu_api import _planner
    g = load_golden('grad_B2_T16')
    planner = _planner(int(g['T']), 2)
    B, T = 2, int(g['T'])
    th, start, goal, sdf = (t64(g[n]) for n in ('th', 'start', 'goal', 'sdf'))
    qc1 = t64(g['qc'])[:1].clone().requires_grad_(True)            # (1,T-1,2,2)
    w1 = t64(g['w'])[:1].clone().requires_grad_(True)              # (1,T,1,1)
    e1 = t64(g['eps'])[:, :1].clone().requires_grad_(True)         # (B,1,1,1)
    dth, _, err_ext = planner.plan_layer(th, start, goal, None, sdf, qc1, w1, e1)
    loss = (dth * t64(g['G'])).sum() + (err_ext * t64(g['g_err_ext'])).sum()
    loss.backward()
    qcf = qc1.detach().expand(B, -1, -1, -1).clone().requires_grad_(True)
    wf = w1.detach().expand(B, -1, -1, -1).clone().requires_grad_(True)
    ef = e1.detach().expand(-1, T, -1, -1).clone().requires_grad_(True)
    dth2, _, err_ext2 = planner.plan_layer(th, start, goal, None, sdf, qcf, wf, ef)
    ((dth2 * t64(g['G'])).sum() + (err_ext2 * t64(g['g_err_ext'])).sum()).backward()
    assert qc1.grad.shape == qc1.shape and w1.grad.shape == w1.shape and e1.grad.shape == e1.shape
    _close(qc1.grad.numpy(), qcf.grad.sum(0, keepdim=True).numpy(), 1e-12, 'qc')
    _close(w1.grad.numpy(), wf.grad.sum(0, keepdim=True).numpy(), 1e-12, 'w')
    _close(e1.grad.numpy(), ef.grad.sum(1, keepdim=True).numpy(), 1e-12, 'eps')
