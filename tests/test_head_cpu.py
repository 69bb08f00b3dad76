"""Fused learned-covariance head, CPU side: the oracle's restatement of the reference's
``get_covariances`` and this package's host mirror are pinned against the LIVE reference
(tests/golden/head_*.npz, oracle/make_golden_head.py); the zero-copy slicing of the raw output
that the kernels read (``PlanLayer.split_head`` -> ``_lib.make_head_weights``) is checked without a GPU."""
import numpy as np
import pytest
import torch

from oracle import gn_oracle
from tests.helpers import head_cases, head_oracle_weights, load_golden, oracle_params, rel_err, t64


def test_head_fixtures_exist():
    assert len(head_cases()) == 4


@pytest.mark.parametrize('name', head_cases())
def test_oracle_covariances_and_step_match_reference(name):
    g = load_golden(name)
    p = oracle_params(g['T'], g['x_lims'], g['y_lims'])
    qc, w, eps, q_full = head_oracle_weights(g, p)
    np.testing.assert_array_equal(qc.numpy(), g['qc'])          # same products, same order: bit-exact
    np.testing.assert_array_equal(w.numpy(), g['w'])
    np.testing.assert_array_equal(eps.numpy(), g['eps'])
    dth, err, err_ext = gn_oracle.gn_step(t64(g['th']), t64(g['start']), t64(g['goal']), t64(g['sdf']), qc, w, eps, p, q_full)
    assert rel_err(dth, g['dth']) < 1e-9
    np.testing.assert_allclose(err.numpy(), g['err'], rtol=1e-12)
    np.testing.assert_allclose(err_ext.numpy(), g['err_ext'], rtol=1e-12)


def _planner(T):
    from diff_gpmp2.gpmp2.diff_gpmp2_planner import DiffGPMP2Planner
    from diff_gpmp2.robot_models import PointRobot2D
    from tests.test_gpu_api import _dicts
    gp, ob, pp, op, ev = _dicts(T, dtype=torch.float64)
    return DiffGPMP2Planner(gp, ob, pp, op, ev, PointRobot2D(torch.tensor(0.4)))


@pytest.mark.parametrize('name', head_cases())
def test_host_get_covariances_matches_reference(name):
    g = load_golden(name)
    mode, learn_eps = str(g['mode']), bool(g['learn_eps'])
    planner = _planner(int(g['T']))
    cov = planner.get_covariances(t64(g['out']), mode, learn_eps)
    cov = list(cov) if isinstance(cov, tuple) else [cov]
    if mode != 'fix_dynamics':
        np.testing.assert_array_equal(cov.pop(0).numpy(), g['qc'])
    np.testing.assert_array_equal(cov.pop(0).numpy(), g['w'])
    if learn_eps:
        np.testing.assert_array_equal(cov.pop(0).numpy(), g['eps'])
    assert not cov


@pytest.mark.parametrize('name', head_cases())
def test_split_head_views_feed_the_kernel_without_copies(name):
    from dgpmp2_b200 import _lib
    g = load_golden(name)
    mode, learn_eps = str(g['mode']), bool(g['learn_eps'])
    T = int(g['T'])
    out = torch.from_numpy(g['out'])
    B, K = out.shape[0], out.shape[2]
    planner = _planner(T)
    q, o, e = planner.plan_layer.split_head(out, mode, learn_eps)
    n = _lib.head_block(mode, 2)
    assert (q is None) == (n == 0) and (e is None) == (not learn_eps)
    cw, keep = _lib.make_head_weights(q, o, e, B, T, n)
    esz = out.element_size()
    if n:
        assert q.shape == (B, T - 1, n) and cw.qc_inv == out.data_ptr()
        assert (cw.qc_stride_b, cw.qc_stride_t) == (K, n)
    assert cw.w_obs == out.data_ptr() + (T - 1) * n * esz and (cw.w_stride_b, cw.w_stride_t) == (K, 1)
    if learn_eps:
        assert cw.eps == out.data_ptr() + ((T - 1) * n + T) * esz and (cw.eps_stride_b, cw.eps_stride_t) == (K, 1)
    else:
        assert not cw.eps
    # the squares / outer products of those views are the reference's covariances
    if n == 1:
        np.testing.assert_array_equal((q.double() ** 2).numpy()[..., 0], g['qc'][..., 0, 0])
    elif n:
        np.testing.assert_array_equal((q.double().unsqueeze(-1) * q.double().unsqueeze(-2)).numpy(), g['qc'])
    np.testing.assert_array_equal((o.double() ** 2).numpy(), g['w'][..., 0, 0])


def test_head_argument_errors():
    from dgpmp2_b200 import _lib
    from tests.gpu_helpers import cparams
    planner = _planner(8)
    with pytest.raises(ValueError):
        planner.plan_layer.split_head(torch.zeros(2, 1, 7 + 8 - 1), 'diag_identity')        # too short
    with pytest.raises(ValueError):
        planner.plan_layer.split_head(torch.zeros(2, 1, 7 + 8 + 9), 'diag_identity', learn_eps=True)   # ragged eps part
    with pytest.raises(NotImplementedError):
        _lib.head_block('diag', 2)                                                        # the reference raises too (:266)
    p = cparams(8)
    with pytest.raises(ValueError):
        _lib.set_head_flags(p, 'q_full')                                                  # params not built for q_full
    _lib.set_head_flags(p, 'qc_full')
    assert p.flags & _lib.FLAG_HEAD and p.flags & _lib.FLAG_HEAD_QC_VEC
    p = cparams(8, q_full=True)
    _lib.set_head_flags(p, 'q_full')
    assert p.flags == _lib.FLAG_Q_FULL | _lib.FLAG_HEAD
    with pytest.raises(ValueError):
        _lib.make_head_weights(torch.zeros(2, 7, 3), None, None, 2, 8, 2)


# ---------------------------------------------------------------------------------------------------------------
# Host logic of the head path (slicing, autograd chain rule through q q^T / o^2 / e^2, planner plumbing) with the
# CUDA entry points replaced by the oracle: what reaches ops.* and what comes back is exercised without a GPU.
# ---------------------------------------------------------------------------------------------------------------
def _fake_ops(monkeypatch, T, x_lims, y_lims):
    from dgpmp2_b200 import _lib, ops
    from dgpmp2_b200.gpmp2 import plan_layer as pl_mod

    def cov(p, qc_inv, w_obs, eps, head, B):
        # the constructor-time constants travel in the params struct (1/sigma^2, Qc^-1, eps), as for the kernels
        op = oracle_params(T, x_lims, y_lims, cost_sigma=p.w_obs_fix ** -0.5,
                           Q_c_inv=[[p.qc_inv_fix[0], p.qc_inv_fix[1]], [p.qc_inv_fix[2], p.qc_inv_fix[3]]], epsilon_dist=p.eps)
        if head is not None:
            n = _lib.head_block(head, 2)
            assert (qc_inv is None) == (n == 0)
            qc = None
            if n:
                assert qc_inv.shape == (B, T - 1, n)
                qc = qc_inv.unsqueeze(-1) * qc_inv.unsqueeze(-2)
                if head == 'diag_identity':
                    qc = qc * torch.eye(2, dtype=qc.dtype)
            w = (w_obs * w_obs).reshape(B, T, 1, 1)
            e = (eps * eps).reshape(B, T, 1, 1) if eps is not None else None
        else:
            qc, w, e = qc_inv, w_obs, eps
        if qc is None:
            qc = torch.tensor([p.qc_inv[i] for i in range(4)], dtype=torch.float64).reshape(2, 2).expand(B, T - 1, 2, 2)
        if w is None:
            w = torch.full((B, T, 1, 1), p.w_obs, dtype=torch.float64)
        if e is None:
            e = torch.full((B, T, 1, 1), p.eps, dtype=torch.float64)
        return op, qc, w, e

    def gn_step(p, th, start, goal, sdf, qc_inv=None, w_obs=None, eps=None, want_status=True, out=None, head=None):
        B = th.shape[0]
        op, qc, w, e = cov(p, qc_inv, w_obs, eps, head, B)
        dth, err, err_ext = gn_oracle.gn_step(th, start, goal, sdf, qc, w, e, op, bool(p.flags & _lib.FLAG_Q_FULL))
        return dth, err.reshape(B), err_ext.reshape(B), torch.zeros(B, dtype=torch.int32)

    def gn_step_backward(p, th, start, goal, sdf, dth, g_dth, g_err_ext=None, qc_inv=None, w_obs=None, eps=None,
                         need_th=True, need_start=False, need_goal=False, need_qc=False, need_w=False, need_eps=False,
                         need_sdf=False, head=None):
        B = th.shape[0]
        with torch.enable_grad():
            op, qc, w, e = cov(p, qc_inv, w_obs, eps, head, B)
            leaves = [t.detach().clone().requires_grad_(True) for t in (th, start, goal, qc, w, e, sdf)]
            d2, _, ee = gn_oracle.gn_step(leaves[0], leaves[1], leaves[2], leaves[6], leaves[3], leaves[4], leaves[5], op,
                                          bool(p.flags & _lib.FLAG_Q_FULL))
            loss = (d2 * g_dth).sum()
            if g_err_ext is not None:
                loss = loss + (ee.reshape(B) * g_err_ext).sum()
            gr = list(torch.autograd.grad(loss, leaves, allow_unused=True))
        gr[4], gr[5] = gr[4].reshape(B, T), gr[5].reshape(B, T)
        needs = [need_th, need_start, need_goal, need_qc, need_w, need_eps, need_sdf]
        return tuple(x if n else None for x, n in zip(gr, needs))

    def errors(p, th, start, goal, sdf, qc_inv=None, w_obs=None, eps=None, head=None):
        B = th.shape[0]
        op, qc, w, e = cov(p, qc_inv, w_obs, eps, head, B)
        _, err, err_ext = gn_oracle.gn_step(th, start, goal, sdf, qc, w, e, op, bool(p.flags & _lib.FLAG_Q_FULL))
        z = torch.zeros(B, dtype=torch.float64)
        return err.reshape(B), err_ext.reshape(B), z, z, z

    def gn_solve(p, th, start, goal, sdf, max_iters, tol_delta, qc_inv=None, w_obs=None, eps=None, head=None):
        B = th.shape[0]
        op, qc, w, e = cov(p, qc_inv, w_obs, eps, head, B)
        th_f, e0, ef, epi, eepi, iters = gn_oracle.gn_solve(th, start, goal, sdf, qc, w, e, op, max_iters, tol_delta)
        pad = lambda rows: torch.tensor([r + [float('nan')] * (max_iters - len(r)) for r in rows], dtype=torch.float64)
        return (th_f, torch.tensor(iters, dtype=torch.int32), pad(epi), pad(eepi), torch.tensor(ef, dtype=torch.float64),
                torch.tensor(ef, dtype=torch.float64), torch.zeros(B, dtype=torch.int32))

    monkeypatch.setattr(ops, 'gn_solve', gn_solve)
    monkeypatch.setattr(ops, 'gn_step', gn_step)
    monkeypatch.setattr(ops, 'gn_step_backward', gn_step_backward)
    monkeypatch.setattr(ops, 'errors', errors)
    monkeypatch.setattr(pl_mod, 'to_cuda', lambda t, dt=None: None if t is None else t.detach().to(dt))


@pytest.mark.parametrize('name', head_cases())
def test_step_head_host_logic_with_oracle_backed_ops(name, monkeypatch):
    g = load_golden(name)
    mode, learn_eps = str(g['mode']), bool(g['learn_eps'])
    T = int(g['T'])
    _fake_ops(monkeypatch, T, g['x_lims'], g['y_lims'])
    planner = _planner(T)
    out = t64(g['out']).requires_grad_(True)
    dth, err, err_ext = planner.step_head(t64(g['th']), t64(g['start']), t64(g['goal']), None, t64(g['sdf']), out, mode, learn_eps)
    assert rel_err(dth.detach(), g['dth']) < 1e-9 and dth.requires_grad and not err.requires_grad
    np.testing.assert_allclose(err.numpy(), g['err'], rtol=1e-12)
    loss = (dth * t64(g['G'])).sum() + (err_ext * t64(g['g_err_ext'])).sum()
    loss.backward()
    scale = np.abs(g['g_out']).max()
    assert np.abs(out.grad.numpy() - g['g_out']).max() <= 1e-8 * scale       # chain rule through the head's products
    e2 = planner.error_batch(t64(g['th']), t64(g['sdf']))
    np.testing.assert_allclose(e2.numpy(), g['err'], rtol=1e-12)


def test_planner_step_with_learn_module_host_logic(monkeypatch):
    g = load_golden('head_diag_identity_B3_T16')
    T, B = int(g['T']), g['th'].shape[0]
    _fake_ops(monkeypatch, T, g['x_lims'], g['y_lims'])
    planner = _planner(T)
    planner.optim_params.update(max_iters=2, tol_delta=0.0)
    scale = torch.ones((), dtype=torch.float64, requires_grad=True)
    planner.set_learn_module(lambda th, im, sdf: t64(g['out']) * scale, 'diag_identity')
    dth, hidden, err, err_ext, qc, w, eps = planner.step(t64(g['th']), t64(g['start']), t64(g['goal']), None, t64(g['sdf']))
    assert hidden is None and rel_err(dth.detach(), g['dth']) < 1e-9
    np.testing.assert_array_equal(qc.numpy(), g['qc'])
    np.testing.assert_array_equal(w.numpy(), g['w'])
    np.testing.assert_allclose(eps.numpy(), g['eps'], rtol=1e-7)      # eps_traj is built in torch's default dtype (:47)
    res = planner.forward(t64(g['th']), t64(g['start']), t64(g['goal']), None, t64(g['sdf']))
    assert res[6] == [2] * B
    res[0].sum().backward()
    assert scale.grad is not None and float(scale.grad.abs()) > 0


def test_header_constants_match_the_python_binding_and_head_flags_are_checked_without_a_gpu():
    """include/dgpmp2_b200.h is the contract: its #defines equal the constants the ctypes binding uses, and the
    C ABI rejects inconsistent head flags before anything is launched (runs on the CPU box)."""
    import ctypes
    import os
    import re
    from dgpmp2_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, 'include', 'dgpmp2_b200.h')).read()
    defs = {k: int(v.strip('()')) for k, v in re.findall(r'#define\s+(DGPMP2_[A-Z_0-9]+)\s+(\(?-?\d+\)?)', src)}
    assert defs['DGPMP2_FLAG_NONHOLONOMIC'] == _lib.FLAG_NONHOLONOMIC and defs['DGPMP2_FLAG_VEL_LIMITS'] == _lib.FLAG_VEL_LIMITS
    assert defs['DGPMP2_FLAG_Q_FULL'] == _lib.FLAG_Q_FULL
    assert defs['DGPMP2_FLAG_HEAD'] == _lib.FLAG_HEAD and defs['DGPMP2_FLAG_HEAD_QC_VEC'] == _lib.FLAG_HEAD_QC_VEC
    assert defs['DGPMP2_OK'] == _lib.OK and defs['DGPMP2_ERR_ARG'] == _lib.ERR_ARG
    assert defs['DGPMP2_ERR_UNSUPPORTED'] == _lib.ERR_UNSUPPORTED and defs['DGPMP2_ERR_CUDA'] == _lib.ERR_CUDA
    flags = [v for k, v in defs.items() if k.startswith('DGPMP2_FLAG_')]
    assert len(set(flags)) == len(flags) and all(f & (f - 1) == 0 for f in flags)      # distinct single bits

    lib = _lib.load()
    p = _lib.make_params(4, 16, 2, 16, 16, (-5, 5), (-5, 5), 10.0, 0.4, 0.01, 0.01, 0.1, torch.eye(2), 0.01, 0.4)
    null = ctypes.c_void_p(0)
    call = lambda w: lib.dgpmp2_gn_step_f32(ctypes.byref(p), null, null, null, null, w, null, null, null, null, null)
    w = _lib.CWeights()
    p.flags = _lib.FLAG_HEAD
    assert call(None) == _lib.ERR_ARG                       # a head without outputs
    p.flags = _lib.FLAG_HEAD_QC_VEC
    assert call(ctypes.byref(w)) == _lib.ERR_ARG            # QC_VEC qualifies HEAD
    p.flags = 32
    assert call(ctypes.byref(w)) == _lib.ERR_ARG            # unknown flag bit
    p.flags = _lib.FLAG_HEAD | _lib.FLAG_HEAD_QC_VEC
    p.B = 0
    assert call(ctypes.byref(w)) == _lib.OK                 # consistent flags, empty batch: accepted, nothing launched


def test_learning_example_runs_with_oracle_backed_ops(monkeypatch):
    """examples/learn_covariances_headless.py (expert labels by the persistent solve, K unrolled differentiable steps
    through the fused head, one flat gradient all-reduce per optimiser step): its host logic end to end, CPU tensors,
    the CUDA entry points replaced by the oracle."""
    import importlib.util
    import os
    from dgpmp2_b200 import _dev
    from dgpmp2_b200.gpmp2 import diff_gpmp2_planner as pm
    T = 8
    _fake_ops(monkeypatch, T, (-5.0, 5.0), (-5.0, 5.0))
    monkeypatch.setattr(pm, 'to_cuda', lambda t, dt=None: None if t is None else t.detach().to(dt))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'examples', 'learn_covariances_headless.py')
    spec = importlib.util.spec_from_file_location('learn_example', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    losses, head = mod.main(['--batch', '3', '--states', str(T), '--unroll', '2', '--iters', '4', '--device', 'cpu', '--f64',
                             '--ext-weight', '0'])       # (the external-loss term runs dgpmp2_errors_backward: -m gpu tests)
    assert len(losses) == 4 and all(l == l and l < float('inf') for l in losses)
    assert losses[-1] < losses[0]                              # the imitation loss goes down
    assert float(head.lin.bias.grad.abs().max()) > 0 and float(head.lin.weight.grad.abs().max()) > 0
    assert _dev is not None
