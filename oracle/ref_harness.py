"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference (read-only at /root/reference)
with the three import-time shims it needs on a current Python stack.

Used by oracle/make_golden.py (to generate tests/golden/*.npz in the build container) and by
the optional live-reference tests.  The reference does not exist on the GPU box; nothing that
runs there imports this module.
"""
import os
import sys
import types


def reference_root():
    for cand in (os.environ.get('DGPMP2_REF'), '/root/reference'):
        if cand and os.path.isdir(os.path.join(cand, 'diff_gpmp2')):
            return cand
    return None


def import_reference():
    """Returns the reference's ``diff_gpmp2`` package or raises ImportError."""
    root = reference_root()
    if root is None:
        raise ImportError('reference tree not found')
    import torch
    import yaml
    # 1. matplotlib is absent (plan_layer.py:9, diff_gpmp2_planner.py:10, env_2d.py:13-14)
    if 'matplotlib' not in sys.modules:
        mpl = types.ModuleType('matplotlib')
        plt = types.ModuleType('matplotlib.pyplot')
        plt.style = types.SimpleNamespace(use=lambda *a, **k: None)
        plt.rcParams = {}
        mpl.pyplot = plt
        mpl.cm = types.ModuleType('matplotlib.cm')
        sys.modules['matplotlib'] = mpl
        sys.modules['matplotlib.pyplot'] = plt
        sys.modules['matplotlib.cm'] = mpl.cm
    # 2. uint8 masks are rejected by masked_select / masked_scatter_ in torch 2.x (plan_layer.py:392-451)
    torch.Tensor.byte = lambda self: self.bool()
    # 3. PyYAML 6 needs a Loader (helpers.py:11-15)
    if not getattr(yaml.load, '_dgpmp2_shim', False):
        _orig = yaml.load

        def _load(stream, Loader=yaml.SafeLoader):
            return _orig(stream, Loader=Loader)
        _load._dgpmp2_shim = True
        yaml.load = _load
    # our own repo also has a top-level ``diff_gpmp2`` alias package: make sure the reference wins here
    for name in [m for m in sys.modules if m == 'diff_gpmp2' or m.startswith('diff_gpmp2.')]:
        del sys.modules[name]
    sys.path.insert(0, root)
    try:
        import diff_gpmp2  # noqa: F401
        import diff_gpmp2.gpmp2.diff_gpmp2_planner  # noqa: F401
        import diff_gpmp2.robot_models  # noqa: F401
        ref = sys.modules['diff_gpmp2']
        assert os.path.abspath(ref.__file__).startswith(os.path.abspath(root)), ref.__file__
    finally:
        sys.path.remove(root)
    return ref


def make_reference_planner(B, T, params, x_lims=(-5.0, 5.0), y_lims=(-5.0, 5.0), learn_params=None):
    """Build the reference's PointRobot2D + DiffGPMP2Planner for batch size B (fp64 defaults)."""
    import torch
    import_reference()
    from diff_gpmp2.robot_models import PointRobot2D
    from diff_gpmp2.gpmp2.diff_gpmp2_planner import DiffGPMP2Planner
    torch.set_default_dtype(torch.float64)
    gp_params = {'Q_c_inv': torch.tensor(params['Q_c_inv']), 'K_s': torch.tensor(params['K_s']),
                 'K_g': torch.tensor(params['K_g'])}
    obs_params = {'cost_sigma': torch.tensor(params['cost_sigma']), 'epsilon_dist': torch.tensor(params['epsilon_dist'])}
    planner_params = {'dof': 2, 'state_dim': 4, 'total_time_sec': params['total_time_sec'],
                      'total_time_step': T - 1}
    optim_params = {'method': 'gauss_newton', 'reg': params['reg'], 'plan_time': 'inf',
                    'max_iters': params.get('max_iters', 100), 'tol_err': params.get('tol_err', 1e-3),
                    'tol_delta': params.get('tol_delta', 1e-4)}
    env_params = {'x_lims': list(x_lims), 'y_lims': list(y_lims)}
    robot = PointRobot2D(torch.tensor(params['sphere_radius']), B, T)
    planner = DiffGPMP2Planner(gp_params, obs_params, planner_params, optim_params, env_params, robot,
                               learn_params=learn_params, batch_size=B)
    return planner
