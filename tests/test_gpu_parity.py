"""-m gpu parity tests proper: the CUDA path (through the C ABI) against the golden fixtures of the
live reference and against the CPU oracle on seeded inputs.

Tolerances (north_star: dtheta within 1e-4 relative for fp32 I/O):
  float64 I/O kernels : dtheta rel <= 1e-9, errors rel <= 1e-11   (same arithmetic type as the reference)
  float32 I/O kernels : dtheta rel <= 1e-5 (only the final rounding to float32 differs; inputs are
                        identical float32 values), errors rel <= 1e-6
"""
import numpy as np
import pytest
import torch

from oracle import gn_oracle
from tests.helpers import golden_weights, load_golden, oracle_params, rel_err, step_cases, t64

pytestmark = pytest.mark.gpu

TOL = {torch.float64: dict(dth=1e-9, err=1e-11, band=1e-12), torch.float32: dict(dth=1e-5, err=1e-6, band=1e-12)}


def _run_case(g, dtype):
    from dgpmp2_b200 import ops
    from tests.gpu_helpers import cparams, dev
    p = oracle_params(g['T'], g['x_lims'], g['y_lims'])
    static = bool(g['static'])
    q_full = bool(g['q_full'])
    cp = cparams(g['T'], x_lims=g['x_lims'], y_lims=g['y_lims'], q_full=q_full)
    th, start, goal, sdf = (dev(g[k], dtype) for k in ('th', 'start', 'goal', 'sdf'))
    kw = {}
    if not static:
        kw = dict(qc_inv=dev(g['qc'], dtype), w_obs=dev(g['w'], dtype), eps=dev(g['eps'], dtype))
    return ops, cp, th, start, goal, sdf, kw


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
@pytest.mark.parametrize('name', step_cases())
def test_gn_step_vs_reference_golden(name, dtype):
    g = load_golden(name)
    ops, cp, th, start, goal, sdf, kw = _run_case(g, dtype)
    dth, err, err_ext, status = ops.gn_step(cp, th, start, goal, sdf, **kw)
    tol = TOL[dtype]
    assert int(status.abs().max()) == 0
    assert rel_err(dth.cpu(), g['dth']) < tol['dth']
    np.testing.assert_allclose(err.cpu().double().numpy().reshape(-1), g['err'].reshape(-1), rtol=tol['err'])
    np.testing.assert_allclose(err_ext.cpu().double().numpy().reshape(-1), g['err_ext'].reshape(-1), rtol=tol['err'])


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
@pytest.mark.parametrize('name', step_cases())
def test_band_factors_errors_vs_reference_golden(name, dtype):
    g = load_golden(name)
    ops, cp, th, start, goal, sdf, kw = _run_case(g, dtype)
    D, U, r = ops.band(cp, th, start, goal, sdf, **kw)
    scale = np.abs(g['band_D']).max()
    assert np.abs(D.cpu().numpy() - g['band_D']).max() < 1e-11 * scale
    assert np.abs(U.cpu().numpy() - g['band_U']).max() < 1e-11 * scale
    assert np.abs(r.cpu().numpy() - g['band_r']).max() < 1e-11 * max(1.0, np.abs(g['band_r']).max())
    gp, oc, oh, _, _ = ops.factors(cp, th, sdf, eps=kw.get('eps'))
    tol = 1e-13 if dtype == torch.float64 else 2e-6
    np.testing.assert_allclose(gp.cpu().double().numpy(), g['gp_err'][..., 0], rtol=0, atol=tol * 10)
    # identical active hinge set (branch-exact SDF lookup)
    np.testing.assert_array_equal(oc.cpu().numpy() > 0, g['obs_cost'][:, :, 0, 0] > 0)
    np.testing.assert_allclose(oc.cpu().double().numpy(), g['obs_cost'][:, :, 0, 0], rtol=0, atol=tol)
    np.testing.assert_allclose(oh.cpu().double().numpy(), g['obs_H'][:, :, 0, :], rtol=tol, atol=tol)
    e, ee, esg, egp, eobs = ops.errors(cp, th, start, goal, sdf, **kw)
    rt = TOL[dtype]['err']
    np.testing.assert_allclose(e.cpu().double().numpy(), g['err'].reshape(-1), rtol=rt)
    np.testing.assert_allclose(ee.cpu().double().numpy(), g['err_ext'].reshape(-1), rtol=rt)
    np.testing.assert_allclose(esg.cpu().double().numpy(), g['err_sg'].reshape(-1), rtol=rt)
    np.testing.assert_allclose(egp.cpu().double().numpy(), g['err_gp'].reshape(-1), rtol=max(rt, 1e-7), atol=1e-12)
    np.testing.assert_allclose(eobs.cpu().double().numpy(), g['err_obs'].reshape(-1), rtol=rt, atol=1e-30)


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
def test_bilinear_bit_exact_vs_reference(dtype):
    from dgpmp2_b200 import ops
    from tests.gpu_helpers import dev
    g = load_golden('bilinear_B3_N40')
    dist, J = ops.sdf_lookup(dev(g['sdf'], dtype), dev(g['pts'], dtype), float(g['res']), g['x_lims'][0], g['y_lims'][0])
    if dtype == torch.float64:
        np.testing.assert_array_equal(dist.cpu().numpy(), g['dist'])       # bit-exact incl. the dist == 0 artefact
        np.testing.assert_array_equal(J.cpu().numpy(), g['J'])
    else:
        np.testing.assert_array_equal(dist.cpu().numpy(), g['dist'].astype(np.float32))
        np.testing.assert_array_equal(J.cpu().numpy(), g['J'].astype(np.float32))


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
@pytest.mark.parametrize('B,T,H,W,seed', [(1, 2, 16, 16, 0), (5, 7, 20, 33, 1), (9, 33, 64, 64, 2), (16, 64, 128, 128, 3),
                                         (3, 100, 48, 48, 4), (2, 257, 64, 64, 5), (33, 16, 32, 32, 6)])
def test_gn_step_vs_oracle_seeded(B, T, H, W, seed, dtype):
    """Random (not straight-line) trajectories, random SDFs, random per-state weights, ragged sizes."""
    from dgpmp2_b200 import ops
    from tests.gpu_helpers import cparams, dev
    rng = np.random.default_rng(100 + seed)
    f32 = lambda a: torch.as_tensor(a).float().double()
    th = f32(rng.uniform(-5.5, 5.5, (B, T, 4)))
    start = f32(rng.uniform(-5, 5, (B, 1, 4)))
    goal = f32(rng.uniform(-5, 5, (B, 1, 4)))
    sdf = f32(rng.uniform(-1.0, 3.0, (B, 1, H, W)))
    q = rng.standard_normal((B, T - 1, 2, 1))
    qc = f32(q @ q.transpose(0, 1, 3, 2) + np.eye(2) * rng.uniform(0.3, 2.0, (B, T - 1, 1, 1)))
    w = f32(rng.uniform(10.0, 2e4, (B, T, 1, 1)))
    eps = f32(rng.uniform(0.0, 1.0, (B, T, 1, 1)))
    p = oracle_params(T)
    ref = gn_oracle.gn_step(th, start, goal, sdf, qc, w, eps, p)
    cp = cparams(T)
    dth, err, err_ext, status = ops.gn_step(cp, dev(th, dtype), dev(start, dtype), dev(goal, dtype), dev(sdf, dtype),
                                            qc_inv=dev(qc, dtype), w_obs=dev(w, dtype), eps=dev(eps, dtype))
    tol = TOL[dtype]
    assert int(status.abs().max()) == 0
    assert rel_err(dth.cpu(), ref[0]) < tol['dth']
    np.testing.assert_allclose(err.cpu().double().numpy(), ref[1].reshape(-1).numpy(), rtol=tol['err'])
    np.testing.assert_allclose(err_ext.cpu().double().numpy(), ref[2].reshape(-1).numpy(), rtol=tol['err'])


def test_forward_config1_vs_reference_golden():
    """The reference's own example flow (config 1): 100 GN iterations in one persistent launch (f64 I/O)."""
    from dgpmp2_b200 import ops
    from tests.gpu_helpers import cparams, dev
    g = load_golden('config1_step_T64')
    cp = cparams(64)
    dt = torch.float64
    out = ops.gn_solve(cp, dev(g['th'], dt), dev(g['start'], dt), dev(g['goal'], dt), dev(g['sdf'], dt), 100, 1e-4)
    th_final, iters, epi, eepi, ef, eef, status = out
    assert iters.cpu().tolist() == list(g['fwd_iters'])
    n = int(iters[0])
    np.testing.assert_allclose(epi[0, :n].cpu().numpy(), g['fwd_err_per_iter'], rtol=1e-7)
    np.testing.assert_allclose(eepi[0, :n].cpu().numpy(), g['fwd_err_ext_per_iter'], rtol=1e-7)
    np.testing.assert_allclose(ef.cpu().numpy(), g['fwd_err_final'], rtol=1e-7)
    assert rel_err(th_final.cpu(), g['fwd_th_final']) < 1e-8


def test_forward_batch_vs_reference_golden():
    from dgpmp2_b200 import ops
    from tests.gpu_helpers import cparams, dev
    g = load_golden('forward_B3_T32')
    cp = cparams(32)
    dt = torch.float64
    out = ops.gn_solve(cp, dev(g['th'], dt), dev(g['start'], dt), dev(g['goal'], dt), dev(g['sdf'], dt),
                       int(g['max_iters']), float(g['tol_delta']))
    th_final, iters, epi, eepi, ef, eef, status = out
    assert iters.cpu().tolist() == list(g['fwd_iters'])
    for b in range(3):
        n = int(iters[b])
        np.testing.assert_allclose(epi[b, :n].cpu().numpy(), g['fwd_err_per_iter'][b, :n], rtol=1e-7)
        assert torch.isnan(epi[b, n:]).all()
    np.testing.assert_allclose(ef.cpu().numpy(), g['fwd_err_final'], rtol=1e-7)
    assert rel_err(th_final.cpu(), g['fwd_th_final']) < 1e-8


@pytest.mark.parametrize('dof,T', [(2, 16), (3, 12)])
def test_gn_step_output_pointer_alignment_does_not_matter(dof, T):
    """The C ABI takes plain pointers: a dth buffer that is only element-aligned (here 4 bytes off a
    16-byte boundary) must give the bits of an aligned one (wide stores are an internal fast path)."""
    from dgpmp2_b200 import ops
    from dgpmp2_b200.datasets.synthetic import make_problems
    B, d = 5, 2 * dof
    pr = make_problems(B, T, dof=dof, unique_envs=2, seed=3)
    th, start, goal, sdf = (pr[k].cuda().float() for k in ('th_init', 'start', 'goal', 'sdf'))
    from tests.gpu_helpers import cparams
    from tests.helpers import XYH, YAML
    cp = cparams(T, dof=dof, base=XYH if dof == 3 else YAML)
    ref = ops.gn_step(cp, th, start, goal, sdf)
    buf = torch.zeros(B * T * d + 8, dtype=torch.float32, device='cuda')
    for off in (1, 2, 4):
        out = buf[off:off + B * T * d].view(B, T, d)
        got = ops.gn_step(cp, th, start, goal, sdf, out=out)
        assert got[0].data_ptr() == buf.data_ptr() + 4 * off
        assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])


@pytest.mark.parametrize('dtype', [torch.float64, torch.float32])
@pytest.mark.parametrize('name', ['custom_nonholonomic_B2_T12', 'custom_nonholonomic_B3_T96', 'custom_vel_limits_B3_T64'])
def test_custom_factor_configs_vs_executed_reference(name, dtype):
    """BASELINE configs 4 and 5 against fixtures produced by EXECUTING the live reference's masks, factor objects,
    construct_linear_system_batch, solve_linear_system_batch and error_batch (oracle/make_golden_r2.py; shape-adapter
    shims only).  cond(Lambda) is 1e5 .. 3e6 here (reg = 0 for config 4), hence 1e-8 rather than 1e-9 for float64."""
    from dgpmp2_b200 import ops
    from tests.gpu_helpers import cparams, dev
    from tests.helpers import custom_base, custom_flags
    g = load_golden(name)
    base, flags, dof = custom_base(g), custom_flags(g), int(g['dof'])
    cp = cparams(g['T'], x_lims=g['x_lims'], y_lims=g['y_lims'], base=base, dof=dof, **flags)
    th, start, goal, sdf = (dev(g[k], dtype) for k in ('th', 'start', 'goal', 'sdf'))
    dth, err, err_ext, status = ops.gn_step(cp, th, start, goal, sdf)
    assert int(status.abs().max()) == 0
    assert rel_err(dth.cpu(), g['dth']) < (1e-8 if dtype == torch.float64 else 1e-5)
    rt = TOL[dtype]['err']
    np.testing.assert_allclose(err.cpu().double().numpy(), g['err'].reshape(-1), rtol=rt)
    np.testing.assert_allclose(err_ext.cpu().double().numpy(), g['err_ext'].reshape(-1), rtol=rt)
    D, U, r = ops.band(cp, th, start, goal, sdf)
    scale = np.abs(g['band_D']).max()
    assert np.abs(D.cpu().numpy() - g['band_D']).max() < 1e-11 * scale
    assert np.abs(U.cpu().numpy() - g['band_U']).max() < 1e-11 * scale
    assert np.abs(r.cpu().numpy() - g['band_r']).max() < 1e-11 * max(1.0, np.abs(g['band_r']).max())
    _, _, _, ce, ch = ops.factors(cp, th, sdf, want_gp=False, want_obs=False, want_custom=True)
    tol = 1e-13 if dtype == torch.float64 else 2e-6
    np.testing.assert_allclose(ce.cpu().double().numpy().reshape(g['cust_err'].shape), g['cust_err'], rtol=0, atol=tol)
    np.testing.assert_allclose(ch.cpu().double().numpy().reshape(g['cust_H'].shape), g['cust_H'], rtol=tol, atol=tol)
    e2 = ops.errors(cp, th, start, goal, sdf)
    np.testing.assert_allclose(e2[0].cpu().double().numpy(), g['err'].reshape(-1), rtol=rt)
