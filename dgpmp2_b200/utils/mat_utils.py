"""Tiny matrix helper kept for API parity (reference ``diff_gpmp2/utils/mat_utils.py:4-6``)."""
import torch


def isotropic_matrix(sig, dim, device=torch.device('cpu')):
    """``sig * I_dim`` on ``device``; ``sig`` may be a python float or a 0-d tensor."""
    eye = torch.eye(int(dim), device=device)
    return eye * sig
