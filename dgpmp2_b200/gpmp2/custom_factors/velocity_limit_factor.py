"""Velocity-limit hinge factor (API mirror of reference
``gpmp2/custom_factors/velocity_limit_factor.py:7-53``): cost = max(|v| - v_lim, 0) per axis,
active when |v| >= v_lim, H = -sign(v) on the velocity entry.  The reference's constructor does
not run on Python 3 (``ndims/2`` is a float); this mirror uses ``ndims // 2`` and evaluates in the
CUDA library (dgpmp2_factors_*).
"""
import torch

from ... import _lib, ops
from ..._dev import as_float, back, to_cuda, work_dtype
from ...utils import mat_utils


class VelocityLimitFactor(object):
    def __init__(self, ndims, num_vel_factors, sig, batch_size=1, use_cuda=False):
        self.use_cuda = torch.cuda.is_available() if use_cuda else False
        self.device = torch.device('cuda') if self.use_cuda else torch.device('cpu')
        self.num_vel_factors = num_vel_factors
        self.ndims = int(ndims)
        self.batch_size = batch_size
        sig = torch.as_tensor(sig)
        self.cov = mat_utils.isotropic_matrix(torch.pow(sig, 2.0), self.ndims, self.device)
        self.inv_cov = mat_utils.isotropic_matrix(1.0 / torch.pow(sig, 2.0), self.ndims // 2, self.device).unsqueeze(0).repeat(num_vel_factors, 1, 1)
        self.vx_traj = None
        self.vy_traj = None

    def get_error_full(self, traj):
        """traj (T,4) -> cost (T,2), H (T,2,4); a batch (B,T,4) gives (B,T,2), (B,T,2,4)."""
        single = traj.dim() == 2
        tb = traj.unsqueeze(0) if single else traj
        B, T, d = tb.shape
        dt = work_dtype(tb)
        p = _lib.make_params(B, T, 2, 1, 1, (0.0, 1.0), (0.0, 1.0), 1.0 * (T - 1), 0.0, 1.0, 1.0, 0.0, torch.eye(2), 1.0, 0.0,
                             use_vel_limits=True, K_v=1.0, v_x=as_float(self.vx_traj), v_y=as_float(self.vy_traj))
        _, _, _, ce, ch = ops.factors(p, to_cuda(tb, dt), want_gp=False, want_obs=False, want_custom=True)
        cost = back(ce, traj).to(traj.dtype)
        H = back(ch, traj).to(traj.dtype)
        return (cost[0], H[0]) if single else (cost, H)

    def get_cov(self):
        return self.cov

    def get_inv_cov_full(self):
        return self.inv_cov

    def set_v_traj(self, vx_traj, vy_traj):
        self.vx_traj = vx_traj
        self.vy_traj = vy_traj
