"""YAML parameter loading (API mirror of reference ``diff_gpmp2/utils/helpers.py:9-60``).

Returns the same 6-tuple (7 with learn params) of dicts with torch tensors where
the reference has them.  Uses ``yaml.safe_load`` (the reference's bare
``yaml.load(fp)`` is a TypeError on PyYAML >= 6).
"""
import numpy as np
import torch
import yaml


def rgb2gray(rgb):
    return np.dot(rgb[..., :3], [0.299, 0.587, 0.114])


def _read(path):
    with open(path, 'r') as fp:
        return yaml.safe_load(fp)


def _tensorise(planner_data, robot_data, device):
    g = planner_data['gpmp2']
    planner_params, gp_params = g['planner_params'], g['gp_params']
    obs_params, optim_params = g['obs_params'], g['optim_params']
    for key in ('Q_c_inv', 'K_s', 'K_g'):
        gp_params[key] = torch.tensor(gp_params[key], device=device)
    if planner_params.get('non_holonomic', False):
        gp_params['K_d'] = torch.tensor(gp_params['K_d'], device=device)
    if planner_params.get('use_vel_limits', False):
        gp_params['K_v'] = torch.tensor(gp_params['K_v'], device=device)
    for key in ('cost_sigma', 'epsilon_dist'):
        obs_params[key] = torch.tensor(obs_params[key], device=device)
    robot_data['sphere_radius'] = torch.tensor(robot_data['sphere_radius'], device=device)
    return planner_params, gp_params, obs_params, optim_params


def load_params(param_file, robot_file, env_file, device):
    planner_data, env_data, robot_data = _read(param_file), _read(env_file), _read(robot_file)
    planner_params, gp_params, obs_params, optim_params = _tensorise(planner_data, robot_data, device)
    return env_data, planner_params, gp_params, obs_params, optim_params, robot_data


def load_params_learn(param_file, robot_file, env_file, learn_params_file, device):
    out = load_params(param_file, robot_file, env_file, device)
    return out + (_read(learn_params_file),)
