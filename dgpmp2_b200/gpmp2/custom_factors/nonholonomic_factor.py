"""Nonholonomic (Dubins-like) factor for the (x, y, h, vx, vy, w) robot (API mirror of reference
``gpmp2/custom_factors/nonholonomic_factor.py:7-39``): e = vy cos h - vx sin h and the Jacobian
row the reference writes (:22-29, reproduced literally).  Evaluated in the CUDA library
(dgpmp2_factors_*) per problem -- the reference's own method only works on one (T,6) trajectory.
"""
import torch

from ... import _lib, ops
from ..._dev import back, to_cuda, work_dtype
from ...utils import mat_utils


class NonHolonomicFactor(object):
    def __init__(self, dof, sig, num_dyn_factors, batch_size=1, use_cuda=False):
        self.use_cuda = torch.cuda.is_available() if use_cuda else False
        self.device = torch.device('cuda') if self.use_cuda else torch.device('cpu')
        self.dof = dof
        self.num_dyn_factors = num_dyn_factors
        sig = torch.as_tensor(sig)
        self.cov = mat_utils.isotropic_matrix(torch.pow(sig, 2.0), 1, self.device).unsqueeze(0).repeat(num_dyn_factors, 1, 1)
        self.inv_cov = mat_utils.isotropic_matrix(1.0 / torch.pow(sig, 2.0), 1, self.device).unsqueeze(0).repeat(num_dyn_factors, 1, 1)

    def get_error_full(self, traj):
        """traj (T,6) -> err (T,1), H (T,6); a batch (B,T,6) gives (B,T,1), (B,T,6)."""
        single = traj.dim() == 2
        tb = traj.unsqueeze(0) if single else traj
        B, T, d = tb.shape
        dt = work_dtype(tb)
        p = _lib.make_params(B, T, 3, 1, 1, (0.0, 1.0), (0.0, 1.0), 1.0 * (T - 1), 0.0, 1.0, 1.0, 0.0, torch.eye(3), 1.0, 0.0,
                             non_holonomic=True, K_d=1.0)
        _, _, _, ce, ch = ops.factors(p, to_cuda(tb, dt), want_gp=False, want_obs=False, want_custom=True)
        err = back(ce, traj).to(traj.dtype).reshape(B, T, 1)
        H = back(ch, traj).to(traj.dtype)
        return (err[0], H[0]) if single else (err, H)

    def get_cov(self):
        return self.cov

    def get_inv_cov(self):
        return self.inv_cov

    def get_inv_cov_full(self):
        return self.inv_cov
