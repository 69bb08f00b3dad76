"""Constant-velocity GP prior factor (API mirror of reference ``gpmp2/gp/gp_factor.py:8-135``).

``get_error`` runs in the CUDA library (dgpmp2_factors_*); the Jacobians H1 = Phi and H2 = -I are
constants returned as expanded views.  ``calc_Q_inv_batch`` is the closed-form Kronecker product
of the reference (:65-73), kept for callers that want the matrix; the GN kernels rebuild it from
Qc^-1 on the fly and never read this tensor.
"""
import torch

from ... import _lib, ops
from ..._dev import back, to_cuda, work_dtype


class GPFactor(object):
    def __init__(self, dof, delta_t, num_gp_factors, batch_size=1, use_cuda=False):
        self.use_cuda = torch.cuda.is_available() if use_cuda else False
        self.device = torch.device('cuda') if self.use_cuda else torch.device('cpu')
        self.dof = int(dof)
        self.delta_t = float(delta_t)
        self.state_dim = 2 * self.dof
        self.num_gp_factors = int(num_gp_factors)
        self.idx1 = torch.arange(0, self.num_gp_factors, device=self.device)
        self.idx2 = torch.arange(1, self.num_gp_factors + 1, device=self.device)
        self.Q_c_inv = None
        self.Q_inv = None

    def calc_phi(self, device=None, dtype=None):
        I = torch.eye(self.dof, device=device or self.device, dtype=dtype or torch.get_default_dtype())
        Z = torch.zeros_like(I)
        return torch.cat((torch.cat((I, self.delta_t * I), dim=1), torch.cat((Z, I), dim=1)), dim=0)

    def calc_Q_inv_batch(self):
        q = self.Q_c_inv
        m1 = 12.0 * (self.delta_t ** -3.0) * q
        m2 = -6.0 * (self.delta_t ** -2.0) * q
        m3 = 4.0 * (self.delta_t ** -1.0) * q
        return torch.cat((torch.cat((m1, m2), dim=-1), torch.cat((m2, m3), dim=-1)), dim=-2)

    calc_Q_inv_full = calc_Q_inv_batch

    def get_error(self, trajb):
        """trajb (B,T,d) -> error (B,T-1,d,1), H1 (B,T-1,d,d), H2 (B,T-1,d,d)."""
        B, T, d = trajb.shape
        dt = work_dtype(trajb)
        p = _lib.make_params(B, T, self.dof, 1, 1, (0.0, 1.0), (0.0, 1.0), self.delta_t * (T - 1), 0.0, 1.0, 1.0, 0.0,
                             torch.eye(self.dof), 1.0, 0.0)
        gp, _, _, _, _ = ops.factors(p, to_cuda(trajb, dt), want_obs=False)
        err = back(gp, trajb).to(trajb.dtype).unsqueeze(-1)
        phi = self.calc_phi(trajb.device, trajb.dtype)
        H1 = phi.reshape(1, 1, d, d).expand(B, T - 1, d, d)
        H2 = (-1.0 * torch.eye(d, device=trajb.device, dtype=trajb.dtype)).reshape(1, 1, d, d).expand(B, T - 1, d, d)
        return err, H1, H2

    def get_error_full(self, traj):
        err, H1, H2 = self.get_error(traj.unsqueeze(0))
        return err[0, :, :, 0], H1[0], H2[0]

    def get_inv_cov_full(self):
        return self.Q_inv

    def set_Q_c_inv(self, Q_c_inv):
        self.Q_c_inv = Q_c_inv
        self.Q_inv = self.calc_Q_inv_batch()

    def set_inv_cov(self, Q_inv):
        self.Q_inv = Q_inv
