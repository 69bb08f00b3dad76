"""Tensor-level entry points: torch CUDA tensors in, torch CUDA tensors out, all compute in
libdgpmp2_b200.so through the C ABI.  These are what the planner / factor classes call.
"""
import ctypes
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import CParams, check, load, make_weights, ptr, stream_ptr, suffix


def _on_tensor_device(fn):
    """Run the op with the CUDA device of its first CUDA tensor argument current: the C ABI launches on the current
    device (stream_ptr() hands it that device's current stream, c_abi.cu sizes its launches from cudaGetDevice), so
    tensors living on cuda:1 while cuda:0 is current would otherwise be launched on the wrong device."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                with torch.cuda.device(a.device):
                    return fn(*args, **kwargs)
        return fn(*args, **kwargs)
    return wrapper


def _prep(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.Dgpmp2Error('%s must be a CUDA tensor (no CPU path)' % name)
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _sdf3(sdf: torch.Tensor, B: int) -> Tuple[torch.Tensor, int]:
    """(B,1,H,W) / (B,H,W) / (1,H,W) / (H,W) -> contiguous (n,H,W) and problem stride in elements."""
    if sdf.dim() == 4:
        sdf = sdf[:, 0]
    if sdf.dim() == 2:
        sdf = sdf.unsqueeze(0)
    if sdf.dim() != 3 or sdf.shape[0] not in (1, B):
        raise ValueError('sdf must be (B,1,H,W), got %s' % (tuple(sdf.shape),))
    sdf = sdf.contiguous()
    stride = 0 if (sdf.shape[0] == 1 and B > 1) else sdf.shape[1] * sdf.shape[2]
    return sdf, stride


def _common(p: CParams, th, start, goal, sdf):
    _lib.require_cuda()
    B, T, d = th.shape
    if d != 2 * p.dof or T != p.T:
        raise ValueError('trajectory shape %s does not match params (T=%d, d=%d)' % (tuple(th.shape), p.T, 2 * p.dof))
    dtype = th.dtype
    th = _prep(th, dtype, 'th')
    start = _prep(start, dtype, 'start').reshape(B, d) if start is not None else None
    goal = _prep(goal, dtype, 'goal').reshape(B, d) if goal is not None else None
    sdf, sdf_sb = _sdf3(_prep(sdf, dtype, 'sdf'), B)
    for name, t in (('start', start), ('goal', goal), ('sdf', sdf)):
        if t is not None and t.device != th.device:
            raise ValueError('%s is on %s but th is on %s: all tensors of a call must share one device' % (name, t.device, th.device))
    p.B = B
    _lib.set_sdf_shape(p, sdf.shape[1], sdf.shape[2], sdf_sb)
    return th, start, goal, sdf


def _weights(p: CParams, dtype, qc_inv, w_obs, eps, B, T, head=None):
    """head: None -> qc_inv / w_obs / eps are covariances; a dynamics_mode string -> they are the RAW outputs
    of the learned module (q_raw (B,T-1,n), o_raw (B,T), e_raw (B,T)) and the kernels form the covariances
    themselves (fused get_covariances, DGPMP2_FLAG_HEAD)."""
    if head is not None:
        _lib.set_head_flags(p, head)
        n = _lib.head_block(head, p.dof)
        if (qc_inv is not None) != (n > 0):
            raise ValueError("dynamics_mode %r %s a raw Qc output" % (head, 'needs' if n else 'does not take'))
        if qc_inv is None and w_obs is None and eps is None:
            raise ValueError('a covariance head needs at least one raw output')
        w, keep = _lib.make_head_weights(_prep_w(qc_inv, dtype), _prep_w(w_obs, dtype), _prep_w(eps, dtype), B, T, n)
        return ctypes.byref(w), keep + [w]
    if qc_inv is None and w_obs is None and eps is None:
        return None, []
    blk = 2 * p.dof if (p.flags & _lib.FLAG_Q_FULL) else p.dof
    qc = _prep_w(qc_inv, dtype)
    wo = _prep_w(w_obs, dtype)
    ep = _prep_w(eps, dtype)
    w, keep = make_weights(qc, wo, ep, B, T, blk)
    return ctypes.byref(w), keep + [w]


def _prep_w(t, dtype):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.Dgpmp2Error('weights must be CUDA tensors')
    return t if t.dtype == dtype else t.to(dtype)


@_on_tensor_device
def gn_step(p: CParams, th, start, goal, sdf, qc_inv=None, w_obs=None, eps=None, want_status=True, out=None, head=None):
    """One batched GN iteration. Returns dth (B,T,d), err (B,), err_ext (B,), status (B,) int32 or None.
    `out`: optional preallocated contiguous CUDA tensor (B,T,d) of th's dtype that receives dth.
    `head`: dynamics_mode string -> qc_inv / w_obs / eps are raw head outputs (see _weights)."""
    th, start, goal, sdf = _common(p, th, start, goal, sdf)
    B, T, d = th.shape
    if out is not None:
        if out.shape != th.shape or out.dtype != th.dtype or not out.is_cuda or not out.is_contiguous():
            raise ValueError('out must be a contiguous CUDA tensor with the shape and dtype of th')
        dth = out
    else:
        dth = torch.empty_like(th)
    err = torch.empty(B, dtype=th.dtype, device=th.device)
    err_ext = torch.empty_like(err)
    status = torch.empty(B, dtype=torch.int32, device=th.device) if want_status else None
    wref, keep = _weights(p, th.dtype, qc_inv, w_obs, eps, B, T, head)
    fn = getattr(load(), 'dgpmp2_gn_step_' + suffix(th.dtype))
    check(fn(ctypes.byref(p), ptr(th), ptr(start), ptr(goal), ptr(sdf), wref, ptr(dth), ptr(err), ptr(err_ext),
             ptr(status), stream_ptr()))
    return dth, err, err_ext, status


@_on_tensor_device
def gn_step_diag(p: CParams, th, start, goal, sdf, qc_inv=None, w_obs=None, eps=None, head=None):
    """gn_step (float32 I/O) that also reports how the mixed-precision kernel solved each problem:
    returns dth, err, err_ext, status, refine (B,) int32 -- k > 0: accepted after k refinement iterations,
    k < 0: handed to the all-double path, 0: the launch used the all-double kernel."""
    th, start, goal, sdf = _common(p, th, start, goal, sdf)
    if th.dtype != torch.float32:
        raise TypeError('gn_step_diag is a float32 entry point')
    B, T, d = th.shape
    dth = torch.empty_like(th)
    err = torch.empty(B, dtype=th.dtype, device=th.device)
    err_ext = torch.empty_like(err)
    status = torch.empty(B, dtype=torch.int32, device=th.device)
    refine = torch.empty(B, dtype=torch.int32, device=th.device)
    wref, keep = _weights(p, th.dtype, qc_inv, w_obs, eps, B, T, head)
    check(load().dgpmp2_gn_step_diag_f32(ctypes.byref(p), ptr(th), ptr(start), ptr(goal), ptr(sdf), wref, ptr(dth),
                                          ptr(err), ptr(err_ext), ptr(status), ptr(refine), stream_ptr()))
    return dth, err, err_ext, status, refine


@_on_tensor_device
def gn_step_backward(p: CParams, th, start, goal, sdf, dth, g_dth, g_err_ext=None, qc_inv=None, w_obs=None, eps=None,
                     need_th=True, need_start=False, need_goal=False, need_qc=False, need_w=False, need_eps=False,
                     need_sdf=False, head=None):
    """Backward of gn_step: returns (g_th, g_start, g_goal, g_qc, g_w, g_eps, g_sdf); entries not asked for are None.
    g_qc is dense (B,T-1,blk,blk), g_w / g_eps are (B,T), g_sdf has the layout of the (contiguous) sdf."""
    th, start, goal, sdf = _common(p, th, start, goal, sdf)
    B, T, d = th.shape
    dt, dev = th.dtype, th.device
    dth = _prep(dth, dt, 'dth')
    g_dth = _prep(g_dth, dt, 'g_dth')
    g_ee = _prep(g_err_ext.reshape(B), dt, 'g_err_ext') if g_err_ext is not None else None
    blk = 2 * p.dof if (p.flags & _lib.FLAG_Q_FULL) else p.dof
    g_th = torch.empty_like(th) if need_th else None
    g_start = torch.empty(B, d, dtype=dt, device=dev) if need_start else None
    g_goal = torch.empty(B, d, dtype=dt, device=dev) if need_goal else None
    g_qc = torch.empty(B, T - 1, blk, blk, dtype=dt, device=dev) if need_qc else None
    g_w = torch.empty(B, T, dtype=dt, device=dev) if need_w else None
    g_eps = torch.empty(B, T, dtype=dt, device=dev) if need_eps else None
    g_sdf = torch.zeros_like(sdf) if need_sdf else None          # accumulated with atomics
    wref, keep = _weights(p, dt, qc_inv, w_obs, eps, B, T, head)
    fn = getattr(load(), 'dgpmp2_gn_step_backward_' + suffix(dt))
    check(fn(ctypes.byref(p), ptr(th), ptr(start), ptr(goal), ptr(sdf), wref, ptr(dth), ptr(g_dth), ptr(g_ee),
             ptr(g_th), ptr(g_start), ptr(g_goal), ptr(g_qc), ptr(g_w), ptr(g_eps), ptr(g_sdf), stream_ptr()))
    return g_th, g_start, g_goal, g_qc, g_w, g_eps, g_sdf


@_on_tensor_device
def gn_solve(p: CParams, th, start, goal, sdf, max_iters: int, tol_delta: float, qc_inv=None, w_obs=None, eps=None,
             head=None):
    """Persistent solve to convergence. Returns th_final, iters, err_per_iter (B,max_iters; NaN beyond iters),
    err_ext_per_iter, err_final, err_ext_final, status."""
    th, start, goal, sdf = _common(p, th, start, goal, sdf)
    B, T, d = th.shape
    dev, dt = th.device, th.dtype
    th_final = torch.empty_like(th)
    iters = torch.empty(B, dtype=torch.int32, device=dev)
    epi = torch.full((B, int(max_iters)), float('nan'), dtype=dt, device=dev)
    eepi = torch.full((B, int(max_iters)), float('nan'), dtype=dt, device=dev)
    ef = torch.empty(B, dtype=dt, device=dev)
    eef = torch.empty(B, dtype=dt, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    wref, keep = _weights(p, dt, qc_inv, w_obs, eps, B, T, head)
    fn = getattr(load(), 'dgpmp2_gn_solve_' + suffix(dt))
    check(fn(ctypes.byref(p), ptr(th), ptr(start), ptr(goal), ptr(sdf), wref, int(max_iters), float(tol_delta),
             ptr(th_final), ptr(iters), ptr(epi), ptr(eepi), ptr(ef), ptr(eef), ptr(status), stream_ptr()))
    return th_final, iters, epi, eepi, ef, eef, status


@_on_tensor_device
def errors(p: CParams, th, start, goal, sdf, qc_inv=None, w_obs=None, eps=None, head=None):
    """Factor sweep. Returns err, err_ext, err_sg, err_gp, err_obs, each (B,)."""
    th, start, goal, sdf = _common(p, th, start, goal, sdf)
    B, T, d = th.shape
    outs = [torch.empty(B, dtype=th.dtype, device=th.device) for _ in range(5)]
    wref, keep = _weights(p, th.dtype, qc_inv, w_obs, eps, B, T, head)
    fn = getattr(load(), 'dgpmp2_errors_' + suffix(th.dtype))
    check(fn(ctypes.byref(p), ptr(th), ptr(start), ptr(goal), ptr(sdf), wref, *[ptr(o) for o in outs], stream_ptr()))
    return tuple(outs)


@_on_tensor_device
def errors_backward(p: CParams, th, start, goal, sdf, g_ext=None, g_sg=None, g_gp=None, g_obs=None, eps=None, head=None):
    """Backward of ``errors`` w.r.t. th: g_* are (B,) upstream gradients of err_ext / err_sg / err_gp / err_obs
    (None = 0; err has no gradient, as in the reference).  Returns g_th (B,T,d)."""
    th, start, goal, sdf = _common(p, th, start, goal, sdf)
    B, T, d = th.shape
    gs = [None if g is None else _prep(g.reshape(B), th.dtype, 'upstream gradient') for g in (g_ext, g_sg, g_gp, g_obs)]
    g_th = torch.empty_like(th)
    wref, keep = None, []
    if eps is not None and head is not None:        # raw head output: the kernel squares it (DGPMP2_FLAG_HEAD)
        p.flags |= _lib.FLAG_HEAD
        w, keep = _lib.make_head_weights(None, None, _prep_w(eps, th.dtype), B, T, 0)
        wref, keep = ctypes.byref(w), keep + [w]
    elif eps is not None:
        wref, keep = _weights(p, th.dtype, None, None, eps, B, T)
    fn = getattr(load(), 'dgpmp2_errors_backward_' + suffix(th.dtype))
    check(fn(ctypes.byref(p), ptr(th), ptr(start), ptr(goal), ptr(sdf), wref, *[ptr(g) for g in gs], ptr(g_th), stream_ptr()))
    return g_th


@_on_tensor_device
def factors(p: CParams, th, sdf=None, eps=None, want_gp=True, want_obs=True, want_custom=False):
    """Stand-alone factor outputs: gp_err (B,T-1,d), obs_cost (B,T), obs_H (B,T,d), cust_err, cust_H (or None)."""
    _lib.require_cuda()
    B, T, d = th.shape
    dt, dev = th.dtype, th.device
    th = _prep(th, dt, 'th')
    p.B = B
    sdf3 = None
    if want_obs:
        sdf3, sb = _sdf3(_prep(sdf, dt, 'sdf'), B)
        _lib.set_sdf_shape(p, sdf3.shape[1], sdf3.shape[2], sb)
    gp = torch.empty(B, T - 1, d, dtype=dt, device=dev) if want_gp else None
    oc = torch.empty(B, T, dtype=dt, device=dev) if want_obs else None
    oh = torch.empty(B, T, d, dtype=dt, device=dev) if want_obs else None
    ce = ch = None
    if want_custom and (p.flags & _lib.FLAG_NONHOLONOMIC):
        ce = torch.empty(B, T, dtype=dt, device=dev)
        ch = torch.empty(B, T, d, dtype=dt, device=dev)
    elif want_custom and (p.flags & _lib.FLAG_VEL_LIMITS):
        ce = torch.empty(B, T, 2, dtype=dt, device=dev)
        ch = torch.empty(B, T, 2, d, dtype=dt, device=dev)
    wref, keep = _weights(p, dt, None, None, eps, B, T)
    fn = getattr(load(), 'dgpmp2_factors_' + suffix(dt))
    check(fn(ctypes.byref(p), ptr(th), ptr(sdf3), wref, ptr(gp), ptr(oc), ptr(oh), ptr(ce), ptr(ch), stream_ptr()))
    return gp, oc, oh, ce, ch


@_on_tensor_device
def sdf_lookup(sdf, pts, res: float, x_lo: float, y_lo: float):
    """bilinear_interpolate: sdf (B,H,W), pts (B,N,2) -> dist (B,N,1), J (B,N,2)."""
    _lib.require_cuda()
    dt = pts.dtype
    B, N, _ = pts.shape
    pts = _prep(pts, dt, 'pts')
    sdf3, sb = _sdf3(_prep(sdf, dt, 'sdf'), B)
    dist = torch.empty(B, N, 1, dtype=dt, device=pts.device)
    J = torch.empty(B, N, 2, dtype=dt, device=pts.device)
    fn = getattr(load(), 'dgpmp2_sdf_lookup_' + suffix(dt))
    check(fn(ptr(sdf3), B, int(sdf3.shape[1]), int(sdf3.shape[2]), sb, ptr(pts), N, float(res), float(x_lo),
             float(y_lo), ptr(dist), ptr(J), stream_ptr()))
    return dist, J


@_on_tensor_device
def hinge_batch(sdf, pts, res: float, x_lo: float, y_lo: float, r_sphere: float, eps=None, eps_const: float = 0.0):
    """hinge_loss_signed_batch: sdf (B,H,W) / (B,1,H,W) / (1,H,W), pts (B,N,2), eps None (-> eps_const) or (B|1,N|1)
    -> cost (B,N), He (B,N,2).  One fused pass: 36 algorithmic bytes per point (fp32)."""
    _lib.require_cuda()
    dt = pts.dtype
    B, N, _ = pts.shape
    pts = _prep(pts, dt, 'pts')
    sdf3, sb = _sdf3(_prep(sdf, dt, 'sdf'), B)
    esb = esn = 0
    if eps is not None:
        eps = _prep_w(eps, dt)
        if eps.dim() != 2 or eps.shape[0] not in (1, B) or eps.shape[1] not in (1, N):
            raise ValueError('eps must be (B|1, N|1), got %s' % (tuple(eps.shape),))
        esb = 0 if eps.shape[0] == 1 else eps.stride(0)
        esn = 0 if eps.shape[1] == 1 else eps.stride(1)
    cost = torch.empty(B, N, dtype=dt, device=pts.device)
    He = torch.empty(B, N, 2, dtype=dt, device=pts.device)
    fn = getattr(load(), 'dgpmp2_hinge_batch_' + suffix(dt))
    check(fn(ptr(sdf3), B, int(sdf3.shape[1]), int(sdf3.shape[2]), sb, ptr(pts), N, float(res), float(x_lo), float(y_lo),
             ptr(eps), esb, esn, float(eps_const), float(r_sphere), ptr(cost), ptr(He), stream_ptr()))
    return cost, He


@_on_tensor_device
def sdf_from_occupancy(im, padlen: int = 1, res: float = 1.0, thresh: float = 0.75):
    """Batched signed distance field on the GPU: im (B,H,W) or (H,W) CUDA tensor -> (B,H+2p,W+2p), exact EDT.
    float32 / float64 images give an SDF of the same dtype; a uint8 image (free = im > thresh) gives float32."""
    _lib.require_cuda()
    if im.dim() == 2:
        im = im.unsqueeze(0)
    if im.dtype == torch.uint8:
        if not im.is_cuda:
            raise _lib.Dgpmp2Error('im must be a CUDA tensor (no CPU path)')
        im = im.contiguous()
        B, H, W = im.shape
        out = torch.empty(B, H + 2 * padlen, W + 2 * padlen, dtype=torch.float32, device=im.device)
        check(load().dgpmp2_sdf_from_occupancy_u8_f32(ptr(im), B, H, W, int(padlen), float(thresh), float(res), ptr(out),
                                                      stream_ptr()))
        return out
    dt = im.dtype if im.dtype in (torch.float32, torch.float64) else torch.float32
    im = _prep(im, dt, 'im')
    B, H, W = im.shape
    out = torch.empty(B, H + 2 * padlen, W + 2 * padlen, dtype=dt, device=im.device)
    fn = getattr(load(), 'dgpmp2_sdf_from_occupancy_' + suffix(dt))
    check(fn(ptr(im), B, H, W, int(padlen), float(thresh), float(res), ptr(out), stream_ptr()))
    return out


@_on_tensor_device
def band(p: CParams, th, start, goal, sdf, qc_inv=None, w_obs=None, eps=None, head=None):
    """Information band in float64: D (B,T,d,d), U (B,T-1,d,d), r (B,T,d)."""
    th, start, goal, sdf = _common(p, th, start, goal, sdf)
    B, T, d = th.shape
    D = torch.empty(B, T, d, d, dtype=torch.float64, device=th.device)
    U = torch.empty(B, T - 1, d, d, dtype=torch.float64, device=th.device)
    r = torch.empty(B, T, d, dtype=torch.float64, device=th.device)
    wref, keep = _weights(p, th.dtype, qc_inv, w_obs, eps, B, T, head)
    fn = getattr(load(), 'dgpmp2_band_' + suffix(th.dtype))
    check(fn(ctypes.byref(p), ptr(th), ptr(start), ptr(goal), ptr(sdf), wref, ptr(D), ptr(U), ptr(r), stream_ptr()))
    return D, U, r


class HostStepper:
    """End-to-end GN step on HOST tensors through dgpmp2_gn_step_host_* (the e2e path of bench.py
    and of the planner when it is handed CPU tensors).  Owns the device workspace and pinned
    output buffers for one problem shape."""

    def __init__(self, p: CParams, dtype=torch.float32, device=None):
        _lib.require_cuda()
        self.p = p
        self.dtype = dtype
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else device
        nbytes = ctypes.c_size_t(0)
        es = 4 if dtype == torch.float32 else 8
        check(load().dgpmp2_host_step_workspace_bytes(ctypes.byref(p), es, ctypes.byref(nbytes)))
        self.ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        B, T, d = p.B, p.T, 2 * p.dof
        self.dth = torch.empty(B, T, d, dtype=dtype).pin_memory()
        self.err = torch.empty(B, dtype=dtype).pin_memory()
        self.err_ext = torch.empty(B, dtype=dtype).pin_memory()
        self.status = torch.empty(B, dtype=torch.int32).pin_memory()
        self.h2d_bytes = (B * T * d + 2 * B * d) * es
        self.sdf_bytes = (p.H * p.W if p.sdf_stride_b == 0 else p.sdf_stride_b * B) * es
        self.d2h_bytes = (B * T * d + 2 * B) * es + 4 * B
        self._sdf_staged = False

    def _check_host(self, t, name, numel):
        if not isinstance(t, torch.Tensor) or t.is_cuda:
            raise _lib.Dgpmp2Error('HostStepper.step: %s must be a HOST tensor' % name)
        if t.dtype != self.dtype:
            raise TypeError('HostStepper.step: %s is %s, the stepper was built for %s' % (name, t.dtype, self.dtype))
        if not t.is_contiguous():
            raise ValueError('HostStepper.step: %s must be contiguous' % name)
        if t.numel() != numel:
            raise ValueError('HostStepper.step: %s has %d elements, expected %d' % (name, t.numel(), numel))

    def step(self, th, start, goal, sdf, sdf_resident=False, in_place=False):
        """Host tensors (contiguous, ideally pinned) -> (dth, err, err_ext, status) pinned host tensors.
        Synchronous: returns when the results are in host memory.  ``sdf_resident=True`` reuses the SDF a previous
        (copying) call of this stepper put in the device workspace (``sdf`` may then be None).  ``in_place=True``
        (DGPMP2_SDF_IN_PLACE): a pinned ``sdf`` is not copied at all -- the kernel reads the taps it needs from the
        host buffer over PCIe (the fast way for an SDF that is used once); a pageable ``sdf`` is copied as usual."""
        p = self.p
        B, T, d = p.B, p.T, 2 * p.dof
        self._check_host(th, 'th', B * T * d)
        self._check_host(start, 'start', B * d)
        self._check_host(goal, 'goal', B * d)
        if sdf_resident and in_place:
            raise ValueError('HostStepper.step: sdf_resident and in_place exclude each other')
        if sdf_resident:
            if not self._sdf_staged:
                raise _lib.Dgpmp2Error('HostStepper.step: sdf_resident=True before any call has staged an SDF')
        else:
            self._check_host(sdf, 'sdf', p.H * p.W if p.sdf_stride_b == 0 else p.sdf_stride_b * B)
        fn = getattr(load(), 'dgpmp2_gn_step_host_' + suffix(self.dtype))
        with torch.cuda.device(self.device):
            read_in_place = bool(in_place and load().dgpmp2_host_pointer_is_mapped(ptr(sdf)))
            check(fn(ctypes.byref(p), ptr(th), ptr(start), ptr(goal), ptr(None if sdf_resident else sdf), ptr(self.dth),
                     ptr(self.err), ptr(self.err_ext), ptr(self.status), ctypes.c_void_p(self.ws.data_ptr()), self.ws.numel(),
                     1 if sdf_resident else (2 if in_place else 0), stream_ptr()))
        if not read_in_place:
            self._sdf_staged = True
        self.last_sdf_read_in_place = read_in_place
        return self.dth, self.err, self.err_ext, self.status


def pack_occupancy_bits(im, thresh: float = 0.75):
    """(B,H,W) occupancy images (free = im > thresh; any dtype, host or device) -> (B,H,ceil(W/32)) int32 words, bit
    (x & 31) of word x >> 5 set = FREE: the input layout of dgpmp2_sdf_from_occupancy_bits_f32 / HostOccStepper."""
    free = (im > thresh)
    B, H, W = free.shape
    nw = (W + 31) // 32
    if W != nw * 32:
        free = torch.nn.functional.pad(free, (0, nw * 32 - W))
    sh = (torch.ones(32, dtype=torch.int64, device=free.device) << torch.arange(32, dtype=torch.int64, device=free.device))
    words = (free.reshape(B, H, nw, 32).to(torch.int64) * sh).sum(-1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words)
    return words.to(torch.int32).contiguous()


@_on_tensor_device
def sdf_from_occupancy_bits(bits, W: int, res: float = 1.0):
    """Exact signed distance fields from bit-packed maps (pack_occupancy_bits): (B,H,ceil(W/32)) int32 -> (B,H,W) float32."""
    _lib.require_cuda()
    if not bits.is_cuda or bits.dtype != torch.int32:
        raise _lib.Dgpmp2Error('bits must be a CUDA int32 tensor')
    bits = bits.contiguous()
    B, H, nw = bits.shape
    if nw != (W + 31) // 32:
        raise ValueError('bits has %d words per row, W = %d needs %d' % (nw, W, (W + 31) // 32))
    out = torch.empty(B, H, W, dtype=torch.float32, device=bits.device)
    check(load().dgpmp2_sdf_from_occupancy_bits_f32(ptr(bits), B, H, int(W), float(res), ptr(out), stream_ptr()))
    return out


class HostOccStepper(HostStepper):
    """End-to-end GN step from HOST trajectories and HOST bit-packed occupancy maps (dgpmp2_gn_step_host_occ_f32): the maps
    cross the bus as 1 bit per pixel and the signed distance fields are built on the device."""

    def __init__(self, p: CParams, device=None):
        super().__init__(p, torch.float32, device)
        nbytes = ctypes.c_size_t(0)
        check(load().dgpmp2_host_step_occ_workspace_bytes(ctypes.byref(p), ctypes.byref(nbytes)))
        self.ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
        self.occ_words = p.B * p.H * ((p.W + 31) // 32)
        self.occ_bytes = 4 * self.occ_words

    def step(self, th, start, goal, occ_bits):
        p = self.p
        B, T, d = p.B, p.T, 2 * p.dof
        self._check_host(th, 'th', B * T * d)
        self._check_host(start, 'start', B * d)
        self._check_host(goal, 'goal', B * d)
        if not isinstance(occ_bits, torch.Tensor) or occ_bits.is_cuda or occ_bits.dtype != torch.int32 or \
                not occ_bits.is_contiguous() or occ_bits.numel() != self.occ_words:
            raise ValueError('occ_bits must be a contiguous host int32 tensor of %d words (pack_occupancy_bits)' % self.occ_words)
        with torch.cuda.device(self.device):
            check(load().dgpmp2_gn_step_host_occ_f32(ctypes.byref(p), ptr(th), ptr(start), ptr(goal), ptr(occ_bits), ptr(self.dth),
                                                     ptr(self.err), ptr(self.err_ext), ptr(self.status),
                                                     ctypes.c_void_p(self.ws.data_ptr()), self.ws.numel(), stream_ptr()))
        return self.dth, self.err, self.err_ext, self.status


def launch_shape(p: CParams, dtype=torch.float32):
    np_, thr, smem, grid = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    check(load().dgpmp2_gn_step_launch_shape(ctypes.byref(p), 4 if dtype == torch.float32 else 8, ctypes.byref(np_),
                                             ctypes.byref(thr), ctypes.byref(smem), ctypes.byref(grid)))
    return {'problems_per_cta': np_.value, 'threads': thr.value, 'smem_bytes': smem.value, 'grid': grid.value}
