"""Hinge-loss obstacle cost (API mirror of reference ``gpmp2/obstacle/obstacle_cost.py:7-38``).

Lookup + hinge + gradient are ONE pass of the CUDA library (dgpmp2_hinge_batch_*: 8 bytes of position in, four
SDF taps, 12 bytes out per point); inside the GN step the same arithmetic is fused into the assembly.
"""
import torch

from ... import ops
from ..._dev import back, to_cuda, work_dtype


class HingeLossObstacleCost(object):
    def __init__(self, env_params, batch_size=1, use_cuda=False):
        self.use_cuda = torch.cuda.is_available() if use_cuda else False
        self.device = torch.device('cuda') if self.use_cuda else torch.device('cpu')
        self.env_params = env_params

    def hinge_loss_signed_batch(self, sphere_centersb, r_vec, epsb, sdfb):
        """sphere_centersb (B,T,nlinks,2), epsb (B,T,nlinks,1)-like, sdfb (B,1,H,W)
        -> cost (B,T,nlinks,1), H (B,T,nlinks,2)."""
        B, T, nl, _ = sphere_centersb.shape
        dt = work_dtype(sphere_centersb, sdfb)
        x_lims, y_lims = self.env_params['x_lims'], self.env_params['y_lims']
        res = (x_lims[1] - x_lims[0]) / (sdfb.shape[-1])
        pts = to_cuda(sphere_centersb, dt).reshape(B, T * nl, 2)
        eps_t = torch.as_tensor(epsb)
        r = float(torch.as_tensor(r_vec).reshape(-1)[0])
        if eps_t.numel() == 1:
            cost, H = ops.hinge_batch(to_cuda(sdfb, dt), pts, res, x_lims[0], y_lims[0], r, eps_const=float(eps_t))
        else:
            e = to_cuda(eps_t, dt)
            if e.numel() == B * T * nl:
                e = e.reshape(B, T * nl)
            elif e.numel() == T * nl:
                e = e.reshape(1, T * nl)
            else:
                e = torch.broadcast_to(e, (B, T, nl, 1)).reshape(B, T * nl)
            cost, H = ops.hinge_batch(to_cuda(sdfb, dt), pts, res, x_lims[0], y_lims[0], r, eps=e)
        out_dt = sphere_centersb.dtype
        return (back(cost, sphere_centersb).to(out_dt).reshape(B, T, nl, 1),
                back(H, sphere_centersb).to(out_dt).reshape(B, T, nl, 2))
