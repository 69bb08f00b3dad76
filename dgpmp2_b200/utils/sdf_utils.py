"""SDF utilities (API mirror of reference ``diff_gpmp2/utils/sdf_utils.py``).

``sdf_2d`` (:6-21) is host-side input preparation (scipy exact EDT, as the reference).
``bilinear_interpolate`` (:38-107) runs in the CUDA library (dgpmp2_sdf_lookup_*).
"""
import numpy as np
import torch


def sdf_2d(image, padlen=1, res=1.0):
    """Signed Euclidean distance transform of an occupancy image (free > 0.75), metres when
    ``res`` is the cell size; positive in free space. ``padlen`` pads with free cells."""
    from scipy import ndimage
    free = np.array(np.asarray(image) > 0.75, dtype=np.float64)
    if padlen > 0:
        free = np.pad(free, (padlen, padlen), 'constant', constant_values=(1.0, 1.0))
    occ = np.array(1.0 - free, dtype=np.float64)
    edt = ndimage.distance_transform_edt
    return (edt(free) - edt(occ)) * res


def sdf_2d_gpu(images, padlen=1, res=1.0):
    """GPU version of ``sdf_2d`` for a batch of occupancy images (B,H,W) (tensor or array): exact
    Euclidean distance transform in the CUDA library (dgpmp2_sdf_from_occupancy_*), bit-identical to
    the scipy path in float64.  Returns a CUDA tensor (B, H+2*padlen, W+2*padlen)."""
    from .. import ops
    from .._dev import cuda_device
    t = torch.as_tensor(np.asarray(images) if not isinstance(images, torch.Tensor) else images)
    if not t.is_floating_point():
        t = t.double()
    return ops.sdf_from_occupancy(t.to(cuda_device()), padlen=padlen, res=res)


def rgb2gray(rgb):
    return np.dot(rgb[..., :3], [0.299, 0.587, 0.114])


def costmap_2d(sdf, eps):
    return (sdf <= eps).double() * (-1.0 * sdf + eps)


def safe_sdf(sdf, eps):
    return -1.0 * sdf + eps


def bilinear_interpolate(imb, stateb, res, x_lims, y_lims, use_cuda=False):
    """``imb`` (B,H,W) or (B,1,H,W) SDF, ``stateb`` (B,N,2) -> ``d_obs`` (B,N,1), ``J`` (B,N,2).

    Same contract as the reference, including its behaviour outside the image (the clamped taps
    give dist = 0, J = 0).  Computed by the CUDA library; CPU inputs are staged to the GPU and the
    results returned on the inputs' device."""
    from .. import ops
    dev = stateb.device
    cuda = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else None
    if cuda is None:
        ops._lib.require_cuda()
    if imb.dim() == 4:
        imb = imb.squeeze(1)
    dist, J = ops.sdf_lookup(imb.to(cuda, stateb.dtype), stateb.detach().to(cuda), float(res), float(x_lims[0]), float(y_lims[0]))
    return dist.to(dev), J.to(dev)
