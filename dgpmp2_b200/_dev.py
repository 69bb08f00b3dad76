"""Device plumbing shared by the host-side API mirror: the compute path is CUDA only, so tensors
handed in on the CPU are staged to the current CUDA device and results are returned on the
caller's device (this is the end-to-end path bench.py times)."""
import torch

from . import _lib


def cuda_device():
    _lib.require_cuda()
    return torch.device('cuda', torch.cuda.current_device())


def to_cuda(t, dtype=None):
    if t is None:
        return None
    dev = cuda_device()
    if t.is_cuda and (dtype is None or t.dtype == dtype):
        return t
    return t.detach().to(device=dev, dtype=dtype if dtype is not None else t.dtype, non_blocking=True)


def back(t, like):
    """Return ``t`` on the device of ``like``."""
    if t is None or like is None or t.device == like.device:
        return t
    return t.to(like.device)


def as_float(x):
    if isinstance(x, torch.Tensor):
        return float(x.detach().double().reshape(-1)[0].item())
    return float(x)


def work_dtype(*tensors):
    """float64 if any floating input is float64 else float32 (the dtype the kernels run their I/O in)."""
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.dtype == torch.float64:
            return torch.float64
    return torch.float32
