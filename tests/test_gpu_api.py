"""-m gpu: the reference-facing Python API (diff_gpmp2.* import paths) against the live-reference goldens.
These read like the reference's example flows: build the planner from the same param dicts, call
step / forward / the factor classes."""
import numpy as np
import pytest
import torch

from tests.helpers import YAML, load_golden, rel_err

pytestmark = pytest.mark.gpu


def _dicts(T, x_lims=(-5.0, 5.0), y_lims=(-5.0, 5.0), dtype=torch.float64, **optim):
    gp_params = {'Q_c_inv': torch.tensor(YAML['Q_c_inv'], dtype=dtype), 'K_s': torch.tensor(YAML['K_s'], dtype=dtype),
                 'K_g': torch.tensor(YAML['K_g'], dtype=dtype), 'K_v': 0.01, 'v_x': 1.0, 'v_y': 1.0}
    obs_params = {'cost_sigma': torch.tensor(YAML['cost_sigma'], dtype=dtype), 'epsilon_dist': torch.tensor(YAML['epsilon_dist'], dtype=dtype)}
    planner_params = {'dof': 2, 'state_dim': 4, 'total_time_sec': 10, 'total_time_step': T - 1}
    optim_params = {'method': 'gauss_newton', 'reg': 0.1, 'plan_time': 'inf', 'max_iters': 100, 'tol_err': 1e-3, 'tol_delta': 1e-4}
    optim_params.update(optim)
    env_params = {'x_lims': list(x_lims), 'y_lims': list(y_lims)}
    return gp_params, obs_params, planner_params, optim_params, env_params


def _planner(T, B, **kw):
    from diff_gpmp2.gpmp2.diff_gpmp2_planner import DiffGPMP2Planner
    from diff_gpmp2.robot_models import PointRobot2D
    gp, ob, pp, op, ev = _dicts(T, **kw)
    robot = PointRobot2D(torch.tensor(YAML['sphere_radius'], dtype=kw.get('dtype', torch.float64)), B, T)
    return DiffGPMP2Planner(gp, ob, pp, op, ev, robot, batch_size=B)


@pytest.mark.parametrize('device', ['cpu', 'cuda'])
@pytest.mark.parametrize('name', ['step_static_B4_T64_k0', 'step_static_B4_T64_k5', 'config1_step_T64', 'step_static_B1_T101_k3'])
def test_planner_step_matches_reference(name, device):
    """DiffGPMP2Planner.step with fp64 tensors on either device (CPU tensors take the staged e2e path)."""
    g = load_golden(name)
    B, T = g['th'].shape[0], int(g['T'])
    planner = _planner(T, B)
    th, start, goal, sdf = (torch.from_numpy(g[k]).double().to(device) for k in ('th', 'start', 'goal', 'sdf'))
    im = torch.zeros_like(sdf)
    dth, hidden, err, err_ext, qc, w, eps = planner.step(th, start, goal, im, sdf)
    assert hidden is None and dth.device.type == device and dth.dtype == torch.float64
    assert dth.shape == (B, T, 4) and err.shape == (B, 1, 1) and err_ext.shape == (B, 1, 1)
    assert qc.shape == (B, T - 1, 2, 2) and w.shape == (B, T, 1, 1) and eps.shape == (B, T, 1, 1)
    assert rel_err(dth.cpu(), g['dth']) < 1e-9
    np.testing.assert_allclose(err.cpu().numpy(), g['err'], rtol=1e-11)
    np.testing.assert_allclose(err_ext.cpu().numpy(), g['err_ext'], rtol=1e-11)
    # error_batch / error_ext_batch / unweighted errors use the state installed by the step
    np.testing.assert_allclose(planner.error_batch(th, sdf).cpu().numpy(), g['err'], rtol=1e-11)
    np.testing.assert_allclose(planner.error_ext_batch(th, sdf).cpu().numpy(), g['err_ext'], rtol=1e-11)
    e_sg, e_gp, e_obs = planner.unweighted_errors_batch(th, sdf)
    np.testing.assert_allclose(e_sg.cpu().numpy(), g['err_sg'], rtol=1e-11)
    np.testing.assert_allclose(e_gp.cpu().numpy(), g['err_gp'], rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(e_obs.cpu().numpy(), g['err_obs'], rtol=1e-11, atol=1e-30)


def test_planner_step_float32_within_north_star_tolerance():
    g = load_golden('step_static_B4_T64_k5')
    planner = _planner(64, 4, dtype=torch.float32)
    th, start, goal, sdf = (torch.from_numpy(g[k]).float().cuda() for k in ('th', 'start', 'goal', 'sdf'))
    dth, _, err, err_ext, _, _, _ = planner.step(th, start, goal, torch.zeros_like(sdf), sdf)
    assert dth.dtype == torch.float32
    assert rel_err(dth.cpu(), g['dth']) < 1e-4          # north_star: 1e-4 relative for fp32 I/O
    np.testing.assert_allclose(err.cpu().double().numpy(), g['err'], rtol=1e-5)


def test_plan_layer_with_learned_weights_matches_reference():
    g = load_golden('step_learned_B3_T16')
    planner = _planner(16, 3)
    t = lambda k: torch.from_numpy(g[k]).double().cuda()
    dth, err, err_ext = planner.plan_layer(t('th'), t('start'), t('goal'), None, t('sdf'), t('qc'), t('w'), t('eps'))
    assert rel_err(dth.cpu(), g['dth']) < 1e-9
    np.testing.assert_allclose(err.cpu().numpy(), g['err'], rtol=1e-11)
    np.testing.assert_allclose(err_ext.cpu().numpy(), g['err_ext'], rtol=1e-11)
    D, U, r = planner.plan_layer.information_band(t('th'), t('sdf'))
    assert np.abs(D.cpu().numpy() - g['band_D']).max() < 1e-11 * np.abs(g['band_D']).max()


def test_plan_layer_q_full_mode_matches_reference():
    from diff_gpmp2.gpmp2 import PlanLayer
    from diff_gpmp2.robot_models import PointRobot2D
    g = load_golden('step_qfull_B2_T16')
    gp, ob, pp, op, ev = _dicts(16)
    layer = PlanLayer(gp, ob, pp, op, ev, PointRobot2D(torch.tensor(0.4, dtype=torch.float64), 2, 16),
                      learn_params={'dgpmp2': {'dynamics_mode': 'q_full'}}, batch_size=2)
    t = lambda k: torch.from_numpy(g[k]).double().cuda()
    dth, err, err_ext = layer(t('th'), t('start'), t('goal'), None, t('sdf'), t('qc'), t('w'), t('eps'))
    assert rel_err(dth.cpu(), g['dth']) < 1e-9
    np.testing.assert_allclose(err.cpu().numpy(), g['err'], rtol=1e-11)
    np.testing.assert_allclose(err_ext.cpu().numpy(), g['err_ext'], rtol=1e-11)


def test_planner_forward_config1_example_flow():
    """examples/diff_gpmp2_2d_example.py:40-67 on env/simple_2d/5.png (SDF from the fixture)."""
    from diff_gpmp2.utils.planner_utils import straight_line_traj
    g = load_golden('config1_step_T64')
    planner = _planner(64, 1)
    start = torch.from_numpy(g['start']).double()
    goal = torch.from_numpy(g['goal']).double()
    torch.set_default_dtype(torch.float64)
    try:
        th_init = straight_line_traj(start[0, :, :2], goal[0, :, :2], 10, 63, 2)
    finally:
        torch.set_default_dtype(torch.float32)
    np.testing.assert_allclose(th_init.numpy(), g['th'][0], atol=1e-6)
    sdf = torch.from_numpy(g['sdf']).double()
    out = planner.forward(torch.from_numpy(g['th']).double(), start, goal, torch.zeros_like(sdf), sdf)
    th_final, hidden, err_init, err_final, err_pi, err_ext_pi, k, timeb = out
    assert hidden is None and k == list(g['fwd_iters']) and len(timeb) == 1
    np.testing.assert_allclose(err_init, g['fwd_err_init'], rtol=1e-9)
    np.testing.assert_allclose(err_final, g['fwd_err_final'], rtol=1e-7)
    np.testing.assert_allclose(err_pi[0], g['fwd_err_per_iter'], rtol=1e-7)
    np.testing.assert_allclose(err_ext_pi[0], g['fwd_err_ext_per_iter'], rtol=1e-7)
    assert rel_err(th_final, g['fwd_th_final']) < 1e-8


def test_planner_forward_batch_early_convergence():
    g = load_golden('forward_B3_T32')
    planner = _planner(32, 3, max_iters=int(g['max_iters']), tol_delta=float(g['tol_delta']))
    t = lambda k: torch.from_numpy(g[k]).double().cuda()
    th_final, _, err_init, err_final, err_pi, _, k, _ = planner.forward(t('th'), t('start'), t('goal'), None, t('sdf'))
    assert k == list(g['fwd_iters']) and len(set(k)) > 1
    for b in range(3):
        np.testing.assert_allclose(err_pi[b], g['fwd_err_per_iter'][b, :k[b]], rtol=1e-7)
    np.testing.assert_allclose(err_final, g['fwd_err_final'], rtol=1e-7)
    assert rel_err(th_final.cpu(), g['fwd_th_final']) < 1e-8
    # a finite plan_time takes the batched step()-loop path: same iterates
    planner2 = _planner(32, 3, max_iters=int(g['max_iters']), tol_delta=float(g['tol_delta']), plan_time=1e9)
    out2 = planner2.forward(t('th'), t('start'), t('goal'), None, t('sdf'))
    assert out2[6] == k
    assert rel_err(out2[0].cpu(), g['fwd_th_final']) < 1e-8


def test_factor_classes_individually_callable():
    from diff_gpmp2.gpmp2.gp import GPFactor, PriorFactor
    from diff_gpmp2.gpmp2.obstacle import ObstacleFactor
    from diff_gpmp2.robot_models import PointRobot2D
    g = load_golden('step_static_B4_T64_k5')
    B, T = 4, 64
    th = torch.from_numpy(g['th']).double().cuda()
    sdf = torch.from_numpy(g['sdf']).double().cuda()
    gpf = GPFactor(2, 10.0 / 63, T - 1)
    e, H1, H2 = gpf.get_error(th)
    np.testing.assert_allclose(e.cpu().numpy(), g['gp_err'], atol=1e-13)
    assert H1.shape == (B, T - 1, 4, 4) and float(H1[0, 0, 0, 2]) == pytest.approx(10.0 / 63)
    assert torch.equal(H2[1, 3].cpu(), -torch.eye(4, dtype=torch.float64))
    gpf.set_Q_c_inv(torch.eye(2, dtype=torch.float64).expand(B, T - 1, 2, 2))
    dt = 10.0 / 63
    assert gpf.get_inv_cov_full()[0, 0, 0, 0] == pytest.approx(12.0 * dt ** -3.0)
    of = ObstacleFactor(4, T, torch.tensor(0.4, dtype=torch.float64), {'x_lims': [-5.0, 5.0], 'y_lims': [-5.0, 5.0]}, PointRobot2D(torch.tensor(0.4, dtype=torch.float64), B, T))
    c, H = of.get_error(th, sdf)
    assert c.shape == (B, T, 1, 1) and H.shape == (B, T, 1, 4)
    np.testing.assert_allclose(c.cpu().numpy(), g['obs_cost'], atol=1e-13)
    np.testing.assert_allclose(H.cpu().numpy(), g['obs_H'], atol=1e-12)
    # hinge cost on explicit sphere centres
    centres, Jfk = PointRobot2D(torch.tensor(0.4, dtype=torch.float64), B, T).get_sphere_centers_batch(th)
    c2, He = of.obs_cost.hinge_loss_signed_batch(centres, torch.tensor(0.4, dtype=torch.float64), torch.full((B, T, 1, 1), 0.4, dtype=torch.float64), sdf)
    np.testing.assert_allclose(c2.cpu().numpy(), g['obs_cost'], atol=1e-13)
    np.testing.assert_allclose(torch.einsum('bsij,bsjk->bsik', He, Jfk).cpu().numpy(), g['obs_H'], atol=1e-12)
    pf = PriorFactor(4, torch.tensor(0.01))
    pf.set_mean(torch.from_numpy(g['start']).double().cuda())
    ep, Hp = pf.get_error(th[:, 0:1])
    np.testing.assert_allclose(ep.cpu().numpy()[:, :, 0], (g['start'][:, 0].astype(np.float64) - g['th'][:, 0].astype(np.float64)))


def test_bilinear_interpolate_api():
    from diff_gpmp2.utils.sdf_utils import bilinear_interpolate
    g = load_golden('bilinear_B3_N40')
    d, J = bilinear_interpolate(torch.from_numpy(g['sdf']).double(), torch.from_numpy(g['pts']).double(), float(g['res']),
                                list(g['x_lims']), list(g['y_lims']))
    assert d.device.type == 'cpu' and d.shape == (3, 40, 1) and J.shape == (3, 40, 2)
    np.testing.assert_array_equal(d.numpy(), g['dist'])
    np.testing.assert_allclose(J.numpy(), g['J'], rtol=1e-15, atol=0)


def test_nonholonomic_factor_api():
    from diff_gpmp2.gpmp2.custom_factors import NonHolonomicFactor
    g = load_golden('nonholonomic_T12')
    f = NonHolonomicFactor(3, torch.tensor(0.01), 12)
    e, H = f.get_error_full(torch.from_numpy(g['traj']).double())
    np.testing.assert_allclose(e.numpy(), g['err'], atol=1e-15)
    np.testing.assert_allclose(H.numpy(), g['H'], atol=1e-15)
    np.testing.assert_allclose(f.get_inv_cov_full().numpy(), g['inv_cov'])


def test_not_positive_definite_raises_like_the_reference():
    planner = _planner(8, 2, reg=-1.0e9)          # makes Lambda indefinite: torch.cholesky raises in the reference
    th = torch.zeros(2, 8, 4, dtype=torch.float64, device='cuda')
    z = torch.zeros(2, 1, 4, dtype=torch.float64, device='cuda')
    sdf = torch.ones(2, 1, 16, 16, dtype=torch.float64, device='cuda')
    with pytest.raises(RuntimeError, match='positive-definite'):
        planner.step(th, z, z, None, sdf)
    planner.plan_layer.strict = False              # opt out of the host sync: status stays on the device
    planner.step(th, z, z, None, sdf)
    assert int(planner.plan_layer.last_status.min()) > 0


@pytest.mark.parametrize('H,W,pad', [(128, 128, 0), (200, 200, 1), (37, 61, 2), (16, 16, 0)])
def test_gpu_sdf_generation_bit_exact_vs_scipy_sdf_2d(H, W, pad):
    """Row (f3): sdf_2d on the GPU == the reference's scipy EDT path (exact EDT => bit-identical in fp64)."""
    from diff_gpmp2.utils.sdf_utils import sdf_2d, sdf_2d_gpu
    from dgpmp2_b200.datasets.synthetic import random_obstacle_map
    rng = np.random.default_rng(H + W)
    ims = [random_obstacle_map(rng, H, 'forest')[:H, :W] if H == W else (rng.random((H, W)) > 0.3).astype(np.float64) for _ in range(5)]
    ims.append(np.ones((H, W)))            # no obstacle at all: scipy's virtual background pixel at (-1,-1)
    ims.append(np.zeros((H, W)))           # fully occupied
    one = np.ones((H, W)); one[H // 2, W // 3] = 0.0
    ims.append(one)
    res = 10.0 / W
    got = sdf_2d_gpu(np.stack(ims), padlen=pad, res=res).cpu().numpy()
    for k, im in enumerate(ims):
        ref = sdf_2d(im, padlen=pad, res=res)
        assert got[k].shape == ref.shape
        np.testing.assert_array_equal(got[k], ref)
    got32 = sdf_2d_gpu(torch.tensor(np.stack(ims), dtype=torch.float32), padlen=pad, res=res).cpu().numpy()
    np.testing.assert_allclose(got32, got, rtol=1e-6, atol=1e-6)
    # uint8 occupancy images (a quarter of the bytes when the maps come from the host): same field in float32
    from dgpmp2_b200 import ops
    u8 = torch.tensor((np.stack(ims) > 0.75).astype(np.uint8) * 255).cuda()
    got8 = ops.sdf_from_occupancy(u8, padlen=pad, res=res, thresh=0.75 * 255).cpu().numpy()
    np.testing.assert_array_equal(got8, got32)


def test_headless_batch_example_runs_end_to_end():
    """The reference's batch-example flow (YAML -> PlanningDataset/DataLoader -> straight lines -> planner.forward)
    through the diff_gpmp2 import paths, on a dataset written in the reference's on-disk format."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'examples', 'diff_gpmp2_2d_batch_example_headless.py')
    spec = importlib.util.spec_from_file_location('headless_example', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        th_final, e0, e1, jb = mod.main(['--batch', '3', '--steps', '31'])
    finally:
        torch.set_default_dtype(torch.float32)
    assert th_final.shape == (3, 32, 4) and th_final.device.type == 'cpu' and len(jb) == 3
    assert all(b < a for a, b in zip(e0, e1))          # Gauss-Newton lowered the cost of every problem
