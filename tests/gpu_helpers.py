"""Helpers for the -m gpu parity tests: build C-ABI params from the same constants as the oracle."""
import torch

from dgpmp2_b200 import _lib
from tests.helpers import YAML


def cparams(T, B=1, H=1, W=1, x_lims=(-5.0, 5.0), y_lims=(-5.0, 5.0), base=YAML, dof=2, **over):
    kw = dict(B=B, T=int(T), dof=dof, H=H, W=W, x_lims=[float(v) for v in x_lims], y_lims=[float(v) for v in y_lims],
              total_time_sec=base['total_time_sec'], r_sphere=base['sphere_radius'], K_s=base['K_s'], K_g=base['K_g'],
              reg=base['reg'], Q_c_inv=base['Q_c_inv'], cost_sigma=base['cost_sigma'],
              epsilon_dist=base['epsilon_dist'], K_d=base.get('K_d'), K_v=base.get('K_v'),
              v_x=base.get('v_x'), v_y=base.get('v_y'))
    kw.update(over)
    return _lib.make_params(**kw)


def dev(a, dtype):
    return torch.as_tensor(a).to(device='cuda', dtype=dtype)
