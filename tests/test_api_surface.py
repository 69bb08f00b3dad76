"""Drop-in boundary: every class / function of the reference's public API on the hot path exists in
this package under the same import path with the same parameter names, in the same order, with defaults where the
reference has them.  The reference's side is a snapshot taken from the LIVE reference
(tests/golden/api_signatures.json, oracle/make_api_snapshot.py).  CPU only."""
import importlib
import inspect
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SNAP = json.load(open(os.path.join(HERE, 'golden', 'api_signatures.json')))


def _resolve(key):
    parts = key.split('.')
    for cut in range(len(parts), 0, -1):
        try:
            obj = importlib.import_module('.'.join(parts[:cut]))
        except ImportError:
            continue
        for name in parts[cut:]:
            obj = getattr(obj, name)
        return obj
    raise ImportError(key)


def _params(fn):
    sig = inspect.signature(fn)
    return [[n, p.default is not inspect.Parameter.empty] for n, p in sig.parameters.items()
            if p.kind in (p.POSITIONAL_OR_KEYWORD, p.KEYWORD_ONLY)]


@pytest.mark.parametrize('key', sorted(k for k, v in SNAP.items() if isinstance(v, list)))
def test_same_parameters_as_the_reference(key):
    ours = _params(_resolve(key))
    ref = SNAP[key]
    # same leading parameters (names and order); this package may append optional ones
    assert [n for n, _ in ours[:len(ref)]] == [n for n, _ in ref], (ours, ref)
    for (n, has_default), (_, ref_default) in zip(ours, ref):
        assert has_default or not ref_default, 'parameter %s lost its default' % n
    for n, has_default in ours[len(ref):]:
        assert has_default, 'extra parameter %s must be optional' % n


def test_the_module_is_this_package_not_the_reference():
    import diff_gpmp2
    import dgpmp2_b200
    assert os.path.dirname(os.path.abspath(diff_gpmp2.__file__)).startswith(os.path.dirname(HERE))
    from diff_gpmp2.gpmp2 import PlanLayer
    assert PlanLayer.__module__.startswith('dgpmp2_b200') or 'dgpmp2_b200' in inspect.getsourcefile(PlanLayer)
    assert dgpmp2_b200 is not None


def _learn_params(mode, learn_eps=False, model_type='feed_forward', **dg):
    """Shaped like the output of the reference's helpers.load_params_learn (utils/helpers.py:35-60) as
    learning/train_planner.py consumes it."""
    d = dict(dynamics_mode=mode, sdf_predict=True, learn_eps=learn_eps)
    d.update(dg)
    return {'model': {'type': model_type, 'dropout_prob': 0.0}, 'dgpmp2': d, 'data': {'im_size': 64, 'expert': 'gpmp2'},
            'optim': {'ext_obs_lambda': 1.0}, 'im_size': 64}


@pytest.mark.parametrize('mode,learn_eps,out_dim', [('fix_dynamics', False, 16), ('fix_dynamics', True, 32),
                                                   ('diag_identity', False, 31), ('qc_full', True, 62),
                                                   ('q_full', False, 76), ('diag', False, 46)])
def test_planner_constructor_accepts_learn_params(mode, learn_eps, out_dim):
    """Reference diff_gpmp2_planner.py:53-90 (used by learning/train_planner.py:691): the constructor derives
    dynamics_mode / learn_eps / out_dim / the constant trajectories from learn_params and exposes learn_module_conv /
    learn_module_fcn; here they are empty slots (the networks are outside this package) and step() raises until
    modules are assigned."""
    import torch
    from diff_gpmp2.gpmp2.diff_gpmp2_planner import DiffGPMP2Planner
    from diff_gpmp2.robot_models import PointRobot2D
    T = 16
    gp = {'Q_c_inv': torch.eye(2), 'K_s': torch.tensor(0.01), 'K_g': torch.tensor(0.01)}
    obs = {'cost_sigma': torch.tensor(0.01), 'epsilon_dist': torch.tensor(0.4)}
    pp = {'dof': 2, 'state_dim': 4, 'total_time_sec': 10.0, 'total_time_step': T - 1}
    op = {'method': 'gauss_newton', 'reg': 0.1, 'plan_time': 'inf', 'max_iters': 10, 'tol_err': 1e-3, 'tol_delta': 1e-4}
    env = {'x_lims': [-5.0, 5.0], 'y_lims': [-5.0, 5.0]}
    lp = _learn_params(mode, learn_eps, dtheta_predict=(mode == 'qc_full'))
    planner = DiffGPMP2Planner(gp, obs, pp, op, env, PointRobot2D(torch.tensor(0.4), 3, T), learn_params=lp, batch_size=3)
    assert planner.dynamics_mode == mode and planner.learn_eps == learn_eps and planner.sdf_predict is True
    assert lp['out_dim'] == out_dim and lp['state_dim'] == 4                       # updated in place, as the reference does
    assert lp['num_traj_states'] == (2 * T if mode == 'qc_full' else T)            # dtheta_predict doubles the input (:66)
    assert planner.res == 10.0 / 64 and planner.model_type == 'feed_forward' and planner.fixed_conv is False
    assert planner.learn_module_conv is None and planner.learn_module_fcn is None
    assert hasattr(planner, 'eps_traj') == (not learn_eps)
    assert hasattr(planner, 'qc_inv_traj') == (mode == 'fix_dynamics')
    assert planner.plan_layer.q_full == (mode == 'q_full')
    th = torch.zeros(3, T, 4)
    with pytest.raises(RuntimeError, match='no learned module is installed'):
        planner.step(th, th[:, :1], th[:, :1], torch.zeros(3, 1, 8, 8), torch.zeros(3, 1, 8, 8))
    # a reference checkpoint's keys line up once modules sit in the slots
    planner.learn_module_conv = torch.nn.Conv2d(2, 4, 3)
    planner.learn_module_fcn = torch.nn.Linear(8, out_dim)
    keys = set(planner.state_dict().keys())
    assert {'learn_module_conv.weight', 'learn_module_conv.bias', 'learn_module_fcn.weight', 'learn_module_fcn.bias'} <= keys
