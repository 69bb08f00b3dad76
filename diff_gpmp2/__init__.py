"""Import-path compatibility layer: the reference's package name ``diff_gpmp2`` mapped onto the
B200-native implementation in ``dgpmp2_b200`` so that the reference's example scripts
(``from diff_gpmp2.gpmp2.diff_gpmp2_planner import DiffGPMP2Planner`` etc.) run unchanged.
Nothing is implemented here; every name resolves to a ``dgpmp2_b200`` module.
"""
import importlib
import sys

_ALIASES = [
    'env', 'env.env_2d',
    'robot_models', 'robot_models.robot_model', 'robot_models.point_robot_2d', 'robot_models.point_robot_xyh',
    'utils', 'utils.helpers', 'utils.sdf_utils', 'utils.planner_utils', 'utils.mat_utils',
    'gpmp2', 'gpmp2.plan_layer', 'gpmp2.diff_gpmp2_planner',
    'gpmp2.gp', 'gpmp2.gp.gp_factor', 'gpmp2.gp.prior_factor',
    'gpmp2.obstacle', 'gpmp2.obstacle.obstacle_factor', 'gpmp2.obstacle.obstacle_cost',
    'gpmp2.custom_factors', 'gpmp2.custom_factors.nonholonomic_factor', 'gpmp2.custom_factors.velocity_limit_factor',
    'datasets', 'datasets.planning_dataset', 'datasets.synthetic',
]

for _name in _ALIASES:
    _mod = importlib.import_module('dgpmp2_b200.' + _name)
    sys.modules[__name__ + '.' + _name] = _mod
    if '.' not in _name:
        setattr(sys.modules[__name__], _name, _mod)
