"""Fused learned-covariance head on the GPU (DGPMP2_FLAG_HEAD): the kernels read the learned module's raw
output and form Qc^-1 / obscov_inv / eps themselves.  Checked (i) against the LIVE reference's
get_covariances -> PlanLayer.forward -> autograd chain (tests/golden/head_*.npz), tolerances as in
test_gpu_parity.TOL (north_star: 1e-4 rel), and (ii) BITWISE against this package's explicit-covariance path
fed with torch's own products in the same element type (the head must round exactly like torch.mul)."""
import numpy as np
import pytest
import torch

from tests.helpers import head_cases, load_golden, rel_err, t64

pytestmark = pytest.mark.gpu

TOL = {torch.float64: dict(dth=1e-9, err=1e-11), torch.float32: dict(dth=1e-5, err=1e-6)}


def _inputs(g, dtype):
    from tests.gpu_helpers import cparams, dev
    mode, learn_eps = str(g['mode']), bool(g['learn_eps'])
    T = int(g['T'])
    th, start, goal, sdf, out = (dev(g[k], dtype) for k in ('th', 'start', 'goal', 'sdf', 'out'))
    mk = lambda: cparams(T, x_lims=g['x_lims'], y_lims=g['y_lims'], q_full=(mode == 'q_full'))
    return mode, learn_eps, T, th, start, goal, sdf, out, mk


def _raw(out, T, mode, learn_eps):
    """the three zero-copy slices of ``out`` (what PlanLayer.split_head returns)"""
    from dgpmp2_b200 import _lib
    B = out.shape[0]
    n = _lib.head_block(mode, 2)
    G, S = T - 1, T
    flat = out[:, 0]
    q = flat[:, :G * n].reshape(B, G, n) if n else None
    o = flat[:, G * n:G * n + S]
    e = flat[:, G * n + S:G * n + 2 * S] if learn_eps else None
    return q, o, e


def _explicit(out, T, mode, learn_eps):
    """covariances formed by torch in out's dtype, exactly as the reference's get_covariances does"""
    q, o, e = _raw(out, T, mode, learn_eps)
    qc = None
    if q is not None:
        qc = q.unsqueeze(-1) * q.unsqueeze(-2)
        if mode == 'diag_identity':
            qc = qc * torch.eye(2, device=out.device, dtype=out.dtype)
    w = (o * o).reshape(out.shape[0], T, 1, 1)
    eps = (e * e).reshape(out.shape[0], T, 1, 1) if e is not None else None
    return qc, w, eps


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
@pytest.mark.parametrize('name', head_cases())
def test_head_step_vs_reference_golden(name, dtype):
    from dgpmp2_b200 import ops
    g = load_golden(name)
    mode, learn_eps, T, th, start, goal, sdf, out, mk = _inputs(g, dtype)
    q, o, e = _raw(out, T, mode, learn_eps)
    dth, err, err_ext, status = ops.gn_step(mk(), th, start, goal, sdf, qc_inv=q, w_obs=o, eps=e, head=mode)
    assert int(status.abs().max()) == 0
    tol = TOL[dtype]
    assert rel_err(dth.cpu(), g['dth']) < tol['dth']
    np.testing.assert_allclose(err.cpu().double().numpy(), g['err'].reshape(-1), rtol=tol['err'])
    np.testing.assert_allclose(err_ext.cpu().double().numpy(), g['err_ext'].reshape(-1), rtol=tol['err'])


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
@pytest.mark.parametrize('name', head_cases())
def test_head_is_bitwise_the_explicit_covariance_path(name, dtype):
    """step, errors, band and the persistent solve: raw output + FLAG_HEAD == torch products + no flag, bit for bit."""
    from dgpmp2_b200 import ops
    g = load_golden(name)
    mode, learn_eps, T, th, start, goal, sdf, out, mk = _inputs(g, dtype)
    q, o, e = _raw(out, T, mode, learn_eps)
    qc, w, eps = _explicit(out, T, mode, learn_eps)
    a = ops.gn_step(mk(), th, start, goal, sdf, qc_inv=q, w_obs=o, eps=e, head=mode)
    b = ops.gn_step(mk(), th, start, goal, sdf, qc_inv=qc, w_obs=w, eps=eps)
    for x, y in zip(a[:3], b[:3]):
        assert torch.equal(x, y)
    a = ops.errors(mk(), th, start, goal, sdf, qc_inv=q, w_obs=o, eps=e, head=mode)
    b = ops.errors(mk(), th, start, goal, sdf, qc_inv=qc, w_obs=w, eps=eps)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    a = ops.band(mk(), th, start, goal, sdf, qc_inv=q, w_obs=o, eps=e, head=mode)
    b = ops.band(mk(), th, start, goal, sdf, qc_inv=qc, w_obs=w, eps=eps)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    a = ops.gn_solve(mk(), th, start, goal, sdf, 3, 0.0, qc_inv=q, w_obs=o, eps=e, head=mode)
    b = ops.gn_solve(mk(), th, start, goal, sdf, 3, 0.0, qc_inv=qc, w_obs=w, eps=eps)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert torch.equal(a[2][:, :3], b[2][:, :3])


@pytest.mark.parametrize('name', head_cases())
def test_head_backward_kernel_is_bitwise_the_explicit_one(name):
    """The backward launch returns gradients w.r.t. the COVARIANCES in head mode too."""
    from dgpmp2_b200 import ops
    from tests.gpu_helpers import dev
    g = load_golden(name)
    dt = torch.float64
    mode, learn_eps, T, th, start, goal, sdf, out, mk = _inputs(g, dt)
    q, o, e = _raw(out, T, mode, learn_eps)
    qc, w, eps = _explicit(out, T, mode, learn_eps)
    dth = ops.gn_step(mk(), th, start, goal, sdf, qc_inv=qc, w_obs=w, eps=eps)[0]
    G, ge = dev(g['G'], dt), dev(g['g_err_ext'], dt).reshape(-1)
    need = dict(need_th=True, need_start=True, need_goal=True, need_qc=qc is not None, need_w=True,
                need_eps=eps is not None, need_sdf=True)
    a = ops.gn_step_backward(mk(), th, start, goal, sdf, dth, G, ge, qc_inv=q, w_obs=o, eps=e, head=mode, **need)
    b = ops.gn_step_backward(mk(), th, start, goal, sdf, dth, G, ge, qc_inv=qc, w_obs=w, eps=eps, **need)
    for x, y in zip(a[:6], b[:6]):
        assert (x is None and y is None) or torch.equal(x, y)
    # the SDF gradient is accumulated with atomics: equal up to summation order
    np.testing.assert_allclose(a[6].cpu().numpy(), b[6].cpu().numpy(), rtol=1e-12, atol=1e-12 * float(b[6].abs().max()))


@pytest.mark.parametrize('device', ['cpu', 'cuda'])
@pytest.mark.parametrize('name', head_cases())
def test_step_head_autograd_vs_reference_autograd(name, device):
    """d(loss)/d(out) through planner.step_head == the reference's autograd through get_covariances + its dense solve."""
    from tests.test_gpu_api import _planner
    g = load_golden(name)
    mode, learn_eps = str(g['mode']), bool(g['learn_eps'])
    B, T = g['th'].shape[0], int(g['T'])
    planner = _planner(T, B)
    mv = lambda k: t64(g[k]).to(device)
    out = mv('out').requires_grad_(True)
    dth, err, err_ext = planner.step_head(mv('th'), mv('start'), mv('goal'), None, mv('sdf'), out, mode, learn_eps)
    assert dth.device.type == device and dth.requires_grad and not err.requires_grad
    assert rel_err(dth.detach().cpu(), g['dth']) < 1e-9
    loss = (dth * mv('G')).sum() + (err_ext * mv('g_err_ext')).sum()
    loss.backward()
    scale = np.abs(g['g_out']).max()
    assert np.abs(out.grad.cpu().numpy() - g['g_out']).max() <= 1e-7 * scale
    # error_batch after the call evaluates with the head's covariances (the reference's statefulness)
    e2 = planner.error_batch(mv('th'), mv('sdf'))
    np.testing.assert_allclose(e2.cpu().numpy(), g['err'], rtol=1e-11)


def test_planner_step_and_forward_with_a_learn_module():
    """set_learn_module: step() returns the reference's 7-tuple (covariances for reporting) from one fused launch;
    forward() re-predicts at every iterate (diff_gpmp2_planner.py:128-147) and stays differentiable."""
    from tests.test_gpu_api import _planner
    g = load_golden('head_diag_identity_B3_T16')
    B, T = g['th'].shape[0], int(g['T'])
    planner = _planner(T, B, max_iters=3, tol_delta=0.0)
    out = t64(g['out']).cuda()
    calls = []

    class Module(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.scale = torch.nn.Parameter(torch.ones((), dtype=torch.float64, device='cuda'))

        def forward(self, th, im, sdf):
            calls.append(tuple(th.shape))
            return out * self.scale

    mod = Module()
    planner.set_learn_module(mod, 'diag_identity')
    mv = lambda k: t64(g[k]).cuda()
    dth, hidden, err, err_ext, qc, w, eps = planner.step(mv('th'), mv('start'), mv('goal'), None, mv('sdf'))
    assert hidden is None and rel_err(dth.detach().cpu(), g['dth']) < 1e-9
    np.testing.assert_array_equal(qc.cpu().numpy(), g['qc'])
    np.testing.assert_array_equal(w.cpu().numpy(), g['w'])
    np.testing.assert_allclose(eps.cpu().numpy(), g['eps'], rtol=1e-7)      # eps_traj is built in torch's default dtype (:47)
    res = planner.forward(mv('th'), mv('start'), mv('goal'), None, mv('sdf'))
    assert res[6] == [3] * B and len(calls) == 1 + 3
    res[0].sum().backward()
    assert mod.scale.grad is not None and torch.isfinite(mod.scale.grad) and float(mod.scale.grad.abs()) > 0


def test_head_flag_validation_in_the_c_abi():
    """A head flag without weights, or QC_VEC without HEAD, is an argument error (no launch)."""
    from dgpmp2_b200 import _lib, ops
    from tests.gpu_helpers import cparams, dev
    g = load_golden('head_diag_identity_B3_T16')
    dt = torch.float32
    th, start, goal, sdf = (dev(g[k], dt) for k in ('th', 'start', 'goal', 'sdf'))
    p = cparams(int(g['T']))
    p.flags |= _lib.FLAG_HEAD
    with pytest.raises(_lib.Dgpmp2Error):
        ops.gn_step(p, th, start, goal, sdf)
    p = cparams(int(g['T']))
    p.flags |= _lib.FLAG_HEAD_QC_VEC
    with pytest.raises(_lib.Dgpmp2Error):
        ops.gn_step(p, th, start, goal, sdf, w_obs=torch.ones(3, 16, 1, 1, device='cuda'))
