"""Obstacle factor (API mirror of reference ``gpmp2/obstacle/obstacle_factor.py:9-60``).

``get_error(trajb, sdfb)`` = sphere centres -> bilinear SDF lookup -> hinge -> Jacobian chained
with the (constant) forward-kinematics Jacobian, all in the CUDA library (dgpmp2_factors_*).
"""
import torch

from ... import _lib, ops
from ..._dev import as_float, back, to_cuda, work_dtype
from .obstacle_cost import HingeLossObstacleCost


class ObstacleFactor(object):
    def __init__(self, state_dim, num_obs_factors, eps, env_params, robot_model, batch_size=1, use_cuda=False):
        self.use_cuda = torch.cuda.is_available() if use_cuda else False
        self.device = torch.device('cuda') if self.use_cuda else torch.device('cpu')
        self.robot_model = robot_model
        self.num_obs_factors = num_obs_factors
        self.state_dim = state_dim
        self.env_params = env_params
        self.eps = eps
        self.inv_cov = None
        self.obs_cost = HingeLossObstacleCost(env_params, use_cuda=self.use_cuda)

    def get_error(self, trajb, sdfb):
        """trajb (B,T,d), sdfb (B,1,H,W) -> error (B,T,1,1), H (B,T,1,d)."""
        B, T, d = trajb.shape
        dt = work_dtype(trajb, sdfb)
        x_lims, y_lims = self.env_params['x_lims'], self.env_params['y_lims']
        eps = self.eps
        eps_t, eps_c = None, 0.0
        if isinstance(eps, torch.Tensor) and eps.numel() > 1:
            eps_t = to_cuda(eps, dt)
            if eps_t.dim() == 3:            # (T, nlinks, 1) planner-style constant trajectory
                eps_t = eps_t.unsqueeze(0)
        else:
            eps_c = as_float(eps)
        p = _lib.make_params(B, T, d // 2, 1, 1, x_lims, y_lims, 1.0 * (T - 1), self.robot_model.get_sphere_radii(),
                             1.0, 1.0, 0.0, torch.eye(d // 2), 1.0, eps_c)
        _, oc, oh, _, _ = ops.factors(p, to_cuda(trajb, dt), to_cuda(sdfb, dt), eps=eps_t, want_gp=False)
        return (back(oc, trajb).to(trajb.dtype).reshape(B, T, 1, 1), back(oh, trajb).to(trajb.dtype).reshape(B, T, 1, d))

    def get_inv_cov(self, idx):
        return self.inv_cov[idx]

    def get_cov(self, idx):
        return torch.pinverse(self.inv_cov[idx])

    def set_inv_cov(self, inv_cov):
        self.inv_cov = inv_cov

    def get_inv_cov_full(self):
        return self.inv_cov

    def set_eps(self, eps):
        self.eps = eps
