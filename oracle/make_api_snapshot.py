"""TEST INFRASTRUCTURE ONLY -- snapshot of the reference's public API on the hot path:
for every class / function the drop-in boundary names, the parameter names (and defaults' presence) of the LIVE
reference, written to tests/golden/api_signatures.json.  tests/test_api_surface.py compares this package's
``diff_gpmp2`` mirror against it.

    python -m oracle.make_api_snapshot       # from the repo root, in the build container
"""
import inspect
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

# (module, object, [methods]) -- SURVEY.md 8(b) "Python surface that must exist"
SURFACE = [
    ('diff_gpmp2.gpmp2.diff_gpmp2_planner', 'DiffGPMP2Planner',
     ['__init__', 'forward', 'step', 'error_batch', 'error_ext_batch', 'unweighted_errors_batch', 'get_covariances',
      'get_obs_covariance']),
    ('diff_gpmp2.gpmp2', 'PlanLayer', ['__init__', 'forward', 'error_batch', 'error_ext_batch', 'gp_error', 'obs_error',
                                       'start_goal_error']),
    ('diff_gpmp2.gpmp2.gp', 'GPFactor', ['__init__', 'get_error', 'set_Q_c_inv', 'set_inv_cov', 'calc_phi', 'calc_Q_inv_batch']),
    ('diff_gpmp2.gpmp2.gp', 'PriorFactor', ['__init__', 'get_error', 'set_mean', 'get_inv_cov_full']),
    ('diff_gpmp2.gpmp2.obstacle', 'ObstacleFactor', ['__init__', 'get_error', 'set_eps', 'set_inv_cov']),
    ('diff_gpmp2.gpmp2.obstacle', 'HingeLossObstacleCost', ['__init__', 'hinge_loss_signed_batch']),
    ('diff_gpmp2.gpmp2.custom_factors', 'NonHolonomicFactor', ['__init__', 'get_error_full', 'get_inv_cov_full']),
    ('diff_gpmp2.gpmp2.custom_factors', 'VelocityLimitFactor', ['__init__', 'get_error_full', 'get_inv_cov_full', 'set_v_traj']),
    ('diff_gpmp2.robot_models', 'PointRobot2D', ['__init__', 'get_sphere_centers_batch', 'forward_kinematics_batch', 'get_sphere_radii']),
    ('diff_gpmp2.robot_models', 'PointRobotXYH', ['__init__', 'get_sphere_radii']),
    ('diff_gpmp2.utils.helpers', 'load_params', None),
    ('diff_gpmp2.utils.sdf_utils', 'sdf_2d', None),
    ('diff_gpmp2.utils.sdf_utils', 'bilinear_interpolate', None),
    ('diff_gpmp2.utils.planner_utils', 'straight_line_traj', None),
    ('diff_gpmp2.utils.planner_utils', 'straight_line_trajb', None),
    ('diff_gpmp2.utils.planner_utils', 'check_convergence', None),
    ('diff_gpmp2.utils.planner_utils', 'check_convergence_batch', None),
    ('diff_gpmp2.utils.mat_utils', 'isotropic_matrix', None),
    ('diff_gpmp2.datasets', 'PlanningDataset', ['__init__', '__len__', '__getitem__']),
]


def describe(fn):
    sig = inspect.signature(fn)
    return [[n, p.default is not inspect.Parameter.empty] for n, p in sig.parameters.items()
            if p.kind in (p.POSITIONAL_OR_KEYWORD, p.KEYWORD_ONLY)]


def snapshot(importer):
    out = {}
    for mod, name, methods in SURFACE:
        try:
            m = importer(mod)
            obj = getattr(m, name)
        except Exception as ex:                      # e.g. datasets needs a package absent here
            out['%s.%s' % (mod, name)] = {'unavailable': type(ex).__name__}
            continue
        if methods is None:
            out['%s.%s' % (mod, name)] = describe(obj)
        else:
            for meth in methods:
                if hasattr(obj, meth):
                    out['%s.%s.%s' % (mod, name, meth)] = describe(getattr(obj, meth))
                else:
                    out['%s.%s.%s' % (mod, name, meth)] = {'unavailable': 'AttributeError'}
    return out


def main():
    import importlib
    from oracle import ref_harness
    ref_harness.import_reference()
    root = ref_harness.reference_root()

    def importer(mod):
        sys.path.insert(0, root)
        try:
            m = importlib.import_module(mod)
        finally:
            sys.path.remove(root)
        assert os.path.abspath(m.__file__).startswith(os.path.abspath(root)), m.__file__
        return m
    snap = snapshot(importer)
    path = os.path.join(ROOT, 'tests', 'golden', 'api_signatures.json')
    json.dump(snap, open(path, 'w'), indent=1, sort_keys=True)
    n_bad = sum(1 for v in snap.values() if isinstance(v, dict))
    print('wrote %s: %d entries, %d unavailable in the reference' % (path, len(snap), n_bad))
    for k, v in sorted(snap.items()):
        if isinstance(v, dict):
            print('  unavailable:', k, v)


if __name__ == '__main__':
    main()
