"""CPU-only tests of the host-side logic and of the C-ABI library's exported surface."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from tests.helpers import YAML

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, 'include', 'dgpmp2_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dgpmp2_[a-z0-9_]+)\s*\(', src)))


def test_library_loads_and_exports_every_declared_symbol():
    from dgpmp2_b200 import _lib
    lib = _lib.load()
    declared = _declared_functions()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(lib, name), 'missing export %s' % name
        assert name in _lib.PROTOTYPES, 'no ctypes prototype for %s' % name
    assert sorted(_lib.PROTOTYPES) == declared
    assert lib.dgpmp2_abi_version() == 1
    assert b'invalid' in lib.dgpmp2_status_string(-1)


def test_argument_checking_without_a_gpu():
    """Bad arguments are rejected before anything is launched (so this runs on the CPU box)."""
    from dgpmp2_b200 import _lib
    lib = _lib.load()
    p = _lib.make_params(4, 64, 2, 16, 16, (-5, 5), (-5, 5), 10.0, 0.4, 0.01, 0.01, 0.1, torch.eye(2), 0.01, 0.4)
    null = ctypes.c_void_p(0)
    assert lib.dgpmp2_gn_step_f32(ctypes.byref(p), null, null, null, null, None, null, null, null, null, null) == _lib.ERR_ARG
    p.dof = 5
    assert lib.dgpmp2_gn_step_f32(ctypes.byref(p), null, null, null, null, None, null, null, null, null, null) == _lib.ERR_UNSUPPORTED
    p.dof = 2
    p.flags = _lib.FLAG_NONHOLONOMIC          # needs dof == 3
    assert lib.dgpmp2_errors_f64(ctypes.byref(p), null, null, null, null, None, null, null, null, null, null, null) == _lib.ERR_ARG
    p.flags = _lib.FLAG_Q_FULL                # needs weights
    assert lib.dgpmp2_band_f64(ctypes.byref(p), null, null, null, null, None, null, null, null, null) == _lib.ERR_ARG
    p.flags = 0
    p.B = 0                                   # empty batch is a no-op, not an error
    assert lib.dgpmp2_gn_step_f64(ctypes.byref(p), null, null, null, null, None, null, null, null, null, null) == _lib.OK
    p.B, p.T = 4, 1
    assert lib.dgpmp2_gn_step_f64(ctypes.byref(p), null, null, null, null, None, null, null, null, null, null) == _lib.ERR_ARG
    with pytest.raises(_lib.Dgpmp2Error):
        _lib.check(_lib.ERR_UNSUPPORTED)


def test_launch_shape_and_limits():
    from dgpmp2_b200 import _lib, ops
    mk = lambda T, dof=2, B=1024: _lib.make_params(B, T, dof, 16, 16, (-5, 5), (-5, 5), 10.0, 0.4, 0.01, 0.01, 0.1, torch.eye(dof), 0.01, 0.4)
    f64 = torch.float64
    s = ops.launch_shape(mk(64), f64)
    # all-double kernel (float64 I/O): the ceil(1024 / 148) = 7 problems an SM has to process share ONE CTA
    # (packed BCR items), 147 CTAs on 148 SMs
    assert s['problems_per_cta'] == 7 and s['grid'] == 147 and s['threads'] == 448
    assert s['smem_bytes'] <= 232448
    assert ops.launch_shape(mk(64, B=8), f64)['problems_per_cta'] == 1   # small batches: one problem per CTA, 4 lanes per item
    assert ops.launch_shape(mk(64, B=8), f64)['threads'] == 128
    assert ops.launch_shape(mk(128), f64)['problems_per_cta'] == 3       # bounded by shared memory (7 would not fit)
    assert ops.launch_shape(mk(96, dof=3, B=512), f64)['problems_per_cta'] == 2
    assert ops.launch_shape(mk(500), f64)['smem_bytes'] <= 232448
    # float32 I/O uses the same kernel by default (the staged trajectory is half as large)
    assert ops.launch_shape(mk(64))['problems_per_cta'] == 7 and ops.launch_shape(mk(128))['problems_per_cta'] == 4
    # opt-in mixed-precision kernel (DGPMP2_PRECISION=32): one thread per state, ceil32(T) threads per problem, fp32
    # records of 208 (d = 4) / 456 (d = 6) bytes per state -> the SM's whole share fits one CTA for every BASELINE config
    import os
    saved = os.environ.get('DGPMP2_PRECISION')
    os.environ['DGPMP2_PRECISION'] = '32'
    try:
        s = ops.launch_shape(mk(64))
        assert s['problems_per_cta'] == 7 and s['grid'] == 147 and s['threads'] == 448 and s['smem_bytes'] <= 232448
        assert ops.launch_shape(mk(64, B=8))['threads'] == 64
        s = ops.launch_shape(mk(128))
        assert s['problems_per_cta'] == 7 and s['threads'] == 896 and s['grid'] == 147
        s = ops.launch_shape(mk(96, dof=3, B=512))
        assert s['problems_per_cta'] == 4 and s['threads'] == 384 and s['grid'] == 128
        assert ops.launch_shape(mk(33, B=100000))['problems_per_cta'] == 15   # one named barrier per problem
    finally:
        os.environ.pop('DGPMP2_PRECISION', None)
        if saved is not None:
            os.environ['DGPMP2_PRECISION'] = saved
    assert ops.launch_shape(mk(512))['smem_bytes'] <= 232448
    with pytest.raises(_lib.Dgpmp2Error):
        ops.launch_shape(mk(600))             # longer than the on-chip band: documented limit
    with pytest.raises(_lib.Dgpmp2Error):
        ops.launch_shape(mk(300, dof=3))


def test_params_follow_the_reference_formulas():
    from dgpmp2_b200 import _lib
    f64 = lambda v: torch.tensor(v, dtype=torch.float64)
    p = _lib.make_params(2, 101, 2, 202, 202, (-5.0, 5.0), (-5.0, 5.0), 10, f64(0.4), f64(0.01),
                         f64(0.01), 0.1, torch.eye(2), f64(0.01), f64(0.4))
    assert p.dt == 10 * 1.0 / 100 * 1.0
    assert p.res == 10.0 / 202
    assert p.ks_inv2 == 1.0 / 0.01 ** 2.0 and p.w_obs_fix == 1.0 / 0.01 ** 2.0
    assert _lib.num_factor_rows(p) == 4 * 102 + 101
    _lib.set_sdf_shape(p, 128, 64, 128 * 64)
    assert p.res == 10.0 / 64                  # cell size follows the SDF width only (obstacle_cost.py:34)
    q = _lib.make_params(2, 96, 3, 8, 8, (-5, 5), (-5, 5), 10, 0.4, 0.01, 0.01, 0.0, torch.eye(3), 0.01, 0.2,
                         non_holonomic=True, K_d=0.01)
    assert _lib.num_factor_rows(q) == 6 * 97 + 96 + 96 and q.kd_inv2 == 1.0 / 0.01 ** 2.0


def test_weights_struct_uses_strides_without_copying():
    from dgpmp2_b200 import _lib
    B, T = 5, 9
    qc = torch.eye(2).reshape(1, 1, 2, 2).expand(B, T - 1, 2, 2)
    w = torch.rand(B, T, 1, 1)
    eps = torch.rand(1, T, 1, 1).expand(B, T, 1, 1)
    cw, keep = _lib.make_weights(qc, w, eps, B, T, 2)
    assert (cw.qc_stride_b, cw.qc_stride_t) == (0, 0) and cw.qc_inv == qc.data_ptr()
    assert (cw.w_stride_b, cw.w_stride_t) == (T, 1)
    assert (cw.eps_stride_b, cw.eps_stride_t) == (0, 1)
    with pytest.raises(ValueError):
        _lib.make_weights(torch.zeros(B, T, 2, 2), None, None, B, T, 2)


def test_straight_line_and_convergence_helpers():
    from diff_gpmp2.utils.planner_utils import check_convergence, check_convergence_batch, straight_line_traj, straight_line_trajb
    s = torch.tensor([[[-4.0, -4.0, 0.0, 0.0]], [[1.0, 2.0, 0.0, 0.0]]])
    g = torch.tensor([[[4.0, 4.0, 0.0, 0.0]], [[3.0, -2.0, 0.0, 0.0]]])
    thb = straight_line_trajb(s, g, 10.0, 63, 2)
    assert thb.shape == (2, 64, 4)
    assert torch.allclose(thb[:, 0, :2], s[:, 0, :2]) and torch.allclose(thb[:, -1, :2], g[:, 0, :2])
    assert torch.allclose(thb[0, :, 2:], torch.tensor([0.8, 0.8]).expand(64, 2))
    th = straight_line_traj(s[0, :, :2], g[0, :, :2], 10.0, 63, 2)
    assert torch.allclose(th, thb[0])
    assert check_convergence(torch.zeros(3), 1, None, 1e-3, 1e-4, 100, verbose=False)
    assert not check_convergence(torch.ones(3), 1, None, 1e-3, 1e-4, 100, verbose=False)
    assert check_convergence(torch.ones(3), 100, None, 1e-3, 1e-4, 100, verbose=False)
    cv = check_convergence_batch(torch.ones(2, 4, 4), 3, torch.tensor([[[1e-6]], [[1.0]]]), 1e-3, 1e-4, 100)
    assert cv.reshape(-1).tolist() == [1, 0]        # only the error-delta test counts (reference quirk)


def test_get_covariances_modes():
    from diff_gpmp2.gpmp2.diff_gpmp2_planner import DiffGPMP2Planner
    from diff_gpmp2.robot_models import PointRobot2D
    from tests.test_gpu_api import _dicts
    gp, ob, pp, op, ev = _dicts(6, dtype=torch.float32)
    planner = DiffGPMP2Planner(gp, ob, pp, op, ev, PointRobot2D(torch.tensor(0.4)))
    G, S, B = 5, 6, 3
    out = torch.randn(B, 1, G * 2 + S + S)
    qc, w, eps = planner.get_covariances(out, 'qc_full', learn_eps=True)
    assert qc.shape == (B, G, 2, 2) and w.shape == (B, S, 1, 1) and eps.shape == (B, S, 1, 1)
    q = out[:, 0, :G * 2].reshape(B, G, 2, 1)
    assert torch.allclose(qc, q @ q.transpose(2, 3))
    qd, wd = planner.get_covariances(torch.randn(B, 1, G + S), 'diag_identity')
    assert torch.all(qd[..., 0, 1] == 0) and torch.all(qd[..., 0, 0] >= 0)
    qf, _ = planner.get_covariances(torch.randn(B, 1, G * 4 + S), 'q_full')
    assert qf.shape == (B, G, 4, 4)
    assert planner.get_covariances(torch.randn(B, 1, S), 'fix_dynamics').shape == (B, S, 1, 1)


def test_load_params_reads_the_reference_yaml_layout(tmp_path):
    from diff_gpmp2.utils.helpers import load_params
    (tmp_path / 'p.yaml').write_text(
        'gpmp2:\n  planner_params: {dof: 2, state_dim: 4, total_time_sec: 10, total_time_step: 63}\n'
        '  gp_params: {Q_c_inv: [[1.0, 0.0], [0.0, 1.0]], K_s: 0.01, K_g: 0.01}\n'
        '  obs_params: {cost_sigma: 0.01, epsilon_dist: 0.4}\n'
        '  optim_params: {method: gauss_newton, reg: 0.1, plan_time: inf, max_iters: 100, tol_err: 0.001, tol_delta: 0.0001}\n')
    (tmp_path / 'r.yaml').write_text('type: point_robot\ndof: 2\nsphere_radius: [0.4]\n')
    (tmp_path / 'e.yaml').write_text('dim: 2\nx_lims: [-5.0, 5.0]\ny_lims: [-5.0, 5.0]\n')
    env, pp, gp, ob, op, rob = load_params(str(tmp_path / 'p.yaml'), str(tmp_path / 'r.yaml'), str(tmp_path / 'e.yaml'), torch.device('cpu'))
    assert pp['total_time_step'] == 63 and isinstance(gp['K_s'], torch.Tensor) and float(rob['sphere_radius'][0]) == pytest.approx(0.4)
    assert env['x_lims'] == [-5.0, 5.0] and float(op['plan_time']) == float('inf')


def test_no_cpu_fallback_without_cuda():
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    from dgpmp2_b200 import _lib
    from diff_gpmp2.utils.sdf_utils import bilinear_interpolate
    with pytest.raises(_lib.Dgpmp2Error, match='no CPU fallback'):
        bilinear_interpolate(torch.zeros(1, 4, 4), torch.zeros(1, 2, 2), 1.0, [-1, 1], [-1, 1])


def test_synthetic_problem_generator_is_deterministic():
    from dgpmp2_b200.datasets.synthetic import make_problems
    a = make_problems(3, 16, im_size=32, seed=5)
    b = make_problems(3, 16, im_size=32, seed=5)
    for k in a:
        assert torch.equal(a[k], b[k])
    assert a['sdf'].shape == (3, 1, 32, 32) and a['th_init'].shape == (3, 16, 4)
    d = (a['goal'][:, 0, :2] - a['start'][:, 0, :2]).norm(dim=1)
    assert torch.all(d >= 0.6 * np.hypot(10, 10) - 1e-5)
    assert (a['sdf'] < 0).any() and (a['sdf'] > 0).any()


def test_bcr_numpy_model_matches_dense_solve():
    from tests.bcr_model import bcr_solve
    rng = np.random.default_rng(0)
    for T, d in [(2, 4), (3, 4), (5, 4), (8, 4), (9, 4), (64, 4), (96, 6), (101, 4), (128, 4)]:
        N = T * d
        A = np.zeros((N, N))
        U = rng.standard_normal((T - 1, d, d))
        for t in range(T - 1):
            A[t * d:(t + 1) * d, (t + 1) * d:(t + 2) * d] = U[t]
        A = A + A.T + np.eye(N) * 30.0
        D = np.stack([A[t * d:(t + 1) * d, t * d:(t + 1) * d] for t in range(T)])
        r = rng.standard_normal((T, d))
        ref = np.linalg.solve(A, r.reshape(-1)).reshape(T, d)
        for tail_max in (1, 2, 4, 8):      # 1 = classic root solve; 4 = the kernel's default
            x = bcr_solve(D, U, r, tail_max=tail_max)
            assert np.abs(x - ref).max() < 1e-12, (T, d, tail_max)


def test_slot_maps_closed_forms_are_inverse_bijections():
    """bcr_slot / bcr_state_of_slot (closed forms used on the device) against the level-table definition."""
    from tests.bcr_model import make_levels, slot, slot_closed_form, state_of_slot, state_of_slot_closed_form
    for T in list(range(2, 140)) + [255, 256, 257, 500, 518]:
        nlev, off = make_levels(T)
        seen = set()
        for t_ in range(T):
            m = slot_closed_form(T, t_)
            assert m == slot(off, T, t_) and 0 <= m < T
            assert state_of_slot_closed_form(T, m) == t_ == state_of_slot(off, nlev, T, m)
            seen.add(m)
        assert len(seen) == T


def test_launch_shape_query_balanced_cta_sizes():
    """dgpmp2_gn_step_launch_shape needs no GPU (148 SMs assumed without a device): one CTA per SM when the SM's share
    fits (B=1024, T=64: 7 problems x 147 CTAs); two CTA sizes when it does not (T=128: 4 problems of 55 KB per CTA ->
    148 x 4 + 144 x 3 = 292 CTAs instead of 256 x 4 in ragged waves); DGPMP2_BALANCED=2 gives the uniform grid."""
    import os
    from dgpmp2_b200 import _lib, ops
    from tests.helpers import YAML
    if torch.cuda.is_available():
        pytest.skip('the shapes below assume the 148-SM default used when no device is present')

    def shape(B, T):
        cp = _lib.make_params(B=B, T=T, dof=2, H=128, W=128, x_lims=(-5.0, 5.0), y_lims=(-5.0, 5.0), total_time_sec=10.0,
                              r_sphere=0.4, K_s=YAML['K_s'], K_g=YAML['K_g'], reg=YAML['reg'], Q_c_inv=YAML['Q_c_inv'],
                              cost_sigma=YAML['cost_sigma'], epsilon_dist=YAML['epsilon_dist'])
        return ops.launch_shape(cp)
    saved = os.environ.pop('DGPMP2_BALANCED', None)
    try:
        s = shape(1024, 64)
        assert (s['problems_per_cta'], s['grid'], s['threads']) == (7, 147, 448)
        s = shape(1024, 128)
        assert (s['problems_per_cta'], s['grid']) == (4, 292) and s['smem_bytes'] <= 232448
        assert 148 * 4 + (292 - 148) * 3 >= 1024 > 148 * 4 + (291 - 148) * 3
        os.environ['DGPMP2_BALANCED'] = '2'
        s = shape(1024, 128)
        assert (s['problems_per_cta'], s['grid']) == (4, 256)
    finally:
        os.environ.pop('DGPMP2_BALANCED', None)
        if saved is not None:
            os.environ['DGPMP2_BALANCED'] = saved
