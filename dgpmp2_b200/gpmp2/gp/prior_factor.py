"""Start / goal prior factor (API mirror of reference ``gpmp2/gp/prior_factor.py:6-35``).

e = mean - state, H = I, K = I / sigma^2.  Inside the GN step the prior is evaluated by the fused
CUDA kernel (csrc/factors.cuh, assemble_node); this class keeps the reference's stateful
interface (``set_mean`` / ``set_inv_cov``) and is individually callable for evaluation code.
"""
import torch

from ...utils import mat_utils


class PriorFactor(object):
    def __init__(self, ndims, sig, batch_size=1, use_cuda=False):
        self.use_cuda = torch.cuda.is_available() if use_cuda else False
        self.device = torch.device('cuda') if self.use_cuda else torch.device('cpu')
        self.ndims = int(ndims)
        self.sig = sig
        self.cov = mat_utils.isotropic_matrix(torch.pow(torch.as_tensor(sig), 2.0), self.ndims, self.device)
        self.meanb = None
        self.inv_cov = None

    def get_error(self, stateb):
        B = stateb.shape[0]
        err = (self.meanb.to(stateb.device) - stateb).reshape(B, self.ndims, 1)
        H = torch.eye(self.ndims, device=stateb.device, dtype=stateb.dtype).unsqueeze(0).expand(B, -1, -1)
        return err, H

    def get_cov(self):
        return self.cov

    def get_inv_cov(self):
        return self.inv_cov

    def set_mean(self, meanb):
        self.meanb = meanb

    def set_inv_cov(self, inv_covb):
        self.inv_cov = inv_covb
