"""Shared test helpers: golden-fixture loading and oracle parameter construction."""
import glob
import os

import numpy as np
import torch

from oracle import gn_oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

YAML = dict(Q_c_inv=[[1.0, 0.0], [0.0, 1.0]], K_s=0.01, K_g=0.01, cost_sigma=0.01, epsilon_dist=0.4,
            reg=0.1, total_time_sec=10.0, sphere_radius=0.4, max_iters=100, tol_delta=1e-4, tol_err=1e-3)
XYH = dict(Q_c_inv=[[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]], K_s=0.01, K_g=0.01, K_d=0.01,
           cost_sigma=0.01, epsilon_dist=0.2, reg=0.0, total_time_sec=10.0, sphere_radius=0.4)


def step_cases():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, '*step*.npz')))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    return {k: z[k] for k in z.files}


def oracle_params(T, x_lims=(-5.0, 5.0), y_lims=(-5.0, 5.0), base=YAML, dof=2, **over):
    kw = dict(dof=dof, T=int(T), total_time_sec=base['total_time_sec'], x_lims=[float(v) for v in x_lims],
              y_lims=[float(v) for v in y_lims], r_sphere=base['sphere_radius'], K_s=base['K_s'], K_g=base['K_g'],
              reg=base['reg'], Q_c_inv=base['Q_c_inv'], cost_sigma=base['cost_sigma'],
              epsilon_dist=base['epsilon_dist'])
    if 'K_d' in base:
        kw['K_d'] = base['K_d']
    kw.update(over)
    return gn_oracle.GNParams(**kw)


def golden_weights(g, p):
    """(qc, w, eps, q_full) as float64 tensors for a golden step case."""
    B = g['th'].shape[0]
    T = int(g['T'])
    if bool(g['static']):
        qc = torch.tensor(p.Q_c_inv, dtype=torch.float64).expand(B, T - 1, p.dof, p.dof)
        w = torch.full((B, T, 1, 1), 1.0 / p.cost_sigma ** 2, dtype=torch.float64)
        eps = torch.full((B, T, 1, 1), p.epsilon_dist, dtype=torch.float64)
        return qc, w, eps, False
    return (torch.from_numpy(g['qc']).double(), torch.from_numpy(g['w']).double(),
            torch.from_numpy(g['eps']).double(), bool(g['q_full']))


def t64(a):
    return torch.from_numpy(np.asarray(a)).double()


def rel_err(a, b):
    a = torch.as_tensor(a).double().reshape(a.shape[0], -1)
    b = torch.as_tensor(b).double().reshape(b.shape[0], -1)
    return (torch.linalg.norm(a - b, dim=1) / torch.linalg.norm(b, dim=1).clamp_min(1e-300)).max().item()


def head_cases():
    return sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLDEN, 'head_*.npz')))


def head_oracle_weights(g, p):
    """Covariances of a head_* golden case via the oracle's restatement of get_covariances, with the
    static values where the mode / learn_eps leaves them constant -> (qc, w, eps, q_full)."""
    mode, learn_eps = str(g['mode']), bool(g['learn_eps'])
    B, T = g['th'].shape[0], int(g['T'])
    qc, w, eps = gn_oracle.covariances_from_head(t64(g['out']), p, mode, learn_eps)
    if qc is None:
        qc = torch.tensor(p.Q_c_inv, dtype=torch.float64).expand(B, T - 1, p.dof, p.dof)
    if eps is None:
        eps = torch.full((B, T, 1, 1), p.epsilon_dist, dtype=torch.float64)
    return qc, w, eps, mode == 'q_full'


def custom_base(g):
    """Planner constants a custom_* golden case was generated with (oracle/make_golden_r2.py)."""
    kv = dict(zip([str(k) for k in g['params_keys']], [str(v) for v in g['params_vals']]))
    import ast
    return {k: ast.literal_eval(v) for k, v in kv.items()}


def custom_flags(g):
    f = {}
    if bool(g['non_holonomic']):
        f['non_holonomic'] = True
    if bool(g['use_vel_limits']):
        f['use_vel_limits'] = True
    return f


def custom_oracle_params(g):
    base = custom_base(g)
    extra = {k: base[k] for k in ('K_v', 'v_x', 'v_y') if k in base and bool(g['use_vel_limits'])}
    return oracle_params(g['T'], g['x_lims'], g['y_lims'], base=base, dof=int(g['dof']), **custom_flags(g), **extra)
