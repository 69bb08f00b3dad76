"""ctypes binding of libdgpmp2_b200.so (the C ABI in include/dgpmp2_b200.h).

The library is the only compute path of this package: if it is missing, or no
CUDA device is usable, every operation raises -- there is no CPU / PyTorch
fallback.  torch is used for device memory and streams only.
"""
import ctypes
import math
import os
from typing import Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DGPMP2_LIB: load another build of the same ABI instead (A/B timing of experimental builds, scratch/*.sh)
LIB_PATH = os.environ.get('DGPMP2_LIB') or os.path.join(_HERE, 'lib', 'libdgpmp2_b200.so')

FLAG_NONHOLONOMIC = 1
FLAG_VEL_LIMITS = 2
FLAG_Q_FULL = 4
FLAG_HEAD = 8            # weights hold the raw outputs of the learned module (fused get_covariances)
FLAG_HEAD_QC_VEC = 16    # ... and Qc^-1 = v v^T from dof values ('qc_full'); alone: q^2 I ('diag_identity')
HEAD_MODES = ('fix_dynamics', 'diag_identity', 'qc_full', 'q_full')

OK, ERR_ARG, ERR_UNSUPPORTED, ERR_CUDA = 0, -1, -2, -3


class CParams(ctypes.Structure):
    """struct dgpmp2_params (include/dgpmp2_b200.h)."""
    _fields_ = [
        ('B', ctypes.c_int32), ('T', ctypes.c_int32), ('dof', ctypes.c_int32),
        ('H', ctypes.c_int32), ('W', ctypes.c_int32), ('flags', ctypes.c_int32),
        ('sdf_stride_b', ctypes.c_int64),
        ('x_lo', ctypes.c_double), ('y_lo', ctypes.c_double), ('res', ctypes.c_double),
        ('dt', ctypes.c_double), ('r_sphere', ctypes.c_double),
        ('ks_inv2', ctypes.c_double), ('kg_inv2', ctypes.c_double), ('reg', ctypes.c_double),
        ('kd_inv2', ctypes.c_double), ('kv_inv2', ctypes.c_double),
        ('vx_lim', ctypes.c_double), ('vy_lim', ctypes.c_double),
        ('qc_inv', ctypes.c_double * 9), ('w_obs', ctypes.c_double), ('eps', ctypes.c_double),
        ('qc_inv_fix', ctypes.c_double * 9), ('w_obs_fix', ctypes.c_double),
    ]


class CWeights(ctypes.Structure):
    """struct dgpmp2_weights."""
    _fields_ = [
        ('qc_inv', ctypes.c_void_p), ('qc_stride_b', ctypes.c_int64), ('qc_stride_t', ctypes.c_int64),
        ('w_obs', ctypes.c_void_p), ('w_stride_b', ctypes.c_int64), ('w_stride_t', ctypes.c_int64),
        ('eps', ctypes.c_void_p), ('eps_stride_b', ctypes.c_int64), ('eps_stride_t', ctypes.c_int64),
    ]


_P = ctypes.POINTER
_vp, _i32, _i64, _f64, _sz = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_size_t

# name -> argtypes (restype is int unless listed in _RESTYPES); every symbol declared in the header is here
PROTOTYPES = {
    'dgpmp2_abi_version': [],
    'dgpmp2_status_string': [ctypes.c_int],
    'dgpmp2_last_cuda_error': [],
    'dgpmp2_gn_step_f32': [_P(CParams), _vp, _vp, _vp, _vp, _P(CWeights), _vp, _vp, _vp, _vp, _vp],
    'dgpmp2_gn_step_f64': [_P(CParams), _vp, _vp, _vp, _vp, _P(CWeights), _vp, _vp, _vp, _vp, _vp],
    'dgpmp2_gn_step_diag_f32': [_P(CParams), _vp, _vp, _vp, _vp, _P(CWeights), _vp, _vp, _vp, _vp, _vp, _vp],
    'dgpmp2_gn_step_backward_f32': [_P(CParams), _vp, _vp, _vp, _vp, _P(CWeights), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    'dgpmp2_gn_step_backward_f64': [_P(CParams), _vp, _vp, _vp, _vp, _P(CWeights), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    'dgpmp2_gn_solve_f32': [_P(CParams), _vp, _vp, _vp, _vp, _P(CWeights), _i32, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    'dgpmp2_gn_solve_f64': [_P(CParams), _vp, _vp, _vp, _vp, _P(CWeights), _i32, _f64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    'dgpmp2_errors_f32': [_P(CParams), _vp, _vp, _vp, _vp, _P(CWeights), _vp, _vp, _vp, _vp, _vp, _vp],
    'dgpmp2_errors_f64': [_P(CParams), _vp, _vp, _vp, _vp, _P(CWeights), _vp, _vp, _vp, _vp, _vp, _vp],
    'dgpmp2_errors_backward_f32': [_P(CParams), _vp, _vp, _vp, _vp, _P(CWeights), _vp, _vp, _vp, _vp, _vp, _vp],
    'dgpmp2_errors_backward_f64': [_P(CParams), _vp, _vp, _vp, _vp, _P(CWeights), _vp, _vp, _vp, _vp, _vp, _vp],
    'dgpmp2_factors_f32': [_P(CParams), _vp, _vp, _P(CWeights), _vp, _vp, _vp, _vp, _vp, _vp],
    'dgpmp2_factors_f64': [_P(CParams), _vp, _vp, _P(CWeights), _vp, _vp, _vp, _vp, _vp, _vp],
    'dgpmp2_sdf_lookup_f32': [_vp, _i32, _i32, _i32, _i64, _vp, _i32, _f64, _f64, _f64, _vp, _vp, _vp],
    'dgpmp2_sdf_lookup_f64': [_vp, _i32, _i32, _i32, _i64, _vp, _i32, _f64, _f64, _f64, _vp, _vp, _vp],
    'dgpmp2_hinge_batch_f32': [_vp, _i32, _i32, _i32, _i64, _vp, _i32, _f64, _f64, _f64, _vp, _i64, _i64, _f64, _f64, _vp, _vp, _vp],
    'dgpmp2_hinge_batch_f64': [_vp, _i32, _i32, _i32, _i64, _vp, _i32, _f64, _f64, _f64, _vp, _i64, _i64, _f64, _f64, _vp, _vp, _vp],
    'dgpmp2_sdf_from_occupancy_f32': [_vp, _i32, _i32, _i32, _i32, _f64, _f64, _vp, _vp],
    'dgpmp2_sdf_from_occupancy_f64': [_vp, _i32, _i32, _i32, _i32, _f64, _f64, _vp, _vp],
    'dgpmp2_sdf_from_occupancy_u8_f32': [_vp, _i32, _i32, _i32, _i32, _f64, _f64, _vp, _vp],
    'dgpmp2_sdf_from_occupancy_bits_f32': [_vp, _i32, _i32, _i32, _f64, _vp, _vp],
    'dgpmp2_host_step_occ_workspace_bytes': [_P(CParams), _P(_sz)],
    'dgpmp2_gn_step_host_occ_f32': [_P(CParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp],
    'dgpmp2_band_f32': [_P(CParams), _vp, _vp, _vp, _vp, _P(CWeights), _vp, _vp, _vp, _vp],
    'dgpmp2_band_f64': [_P(CParams), _vp, _vp, _vp, _vp, _P(CWeights), _vp, _vp, _vp, _vp],
    'dgpmp2_host_step_workspace_bytes': [_P(CParams), _i32, _P(_sz)],
    'dgpmp2_host_pointer_is_mapped': [_vp],
    'dgpmp2_gn_step_host_f32': [_P(CParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i32, _vp],
    'dgpmp2_gn_step_host_f64': [_P(CParams), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _i32, _vp],
    'dgpmp2_gn_step_launch_shape': [_P(CParams), _i32, _P(_i32), _P(_i32), _P(_i32), _P(_i32)],
}
_RESTYPES = {'dgpmp2_status_string': ctypes.c_char_p, 'dgpmp2_last_cuda_error': ctypes.c_char_p}

_lib = None


class Dgpmp2Error(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Dgpmp2Error(
            'libdgpmp2_b200.so is not built (%s). Build it with `python -m dgpmp2_b200.build` '
            '(nvcc, sm_100a). There is no CPU fallback.' % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, ctypes.c_int)
    if lib.dgpmp2_abi_version() != 1:
        raise Dgpmp2Error('libdgpmp2_b200.so ABI version mismatch')
    _lib = lib
    return lib


def check(rc):
    if rc == OK:
        return
    lib = load()
    msg = lib.dgpmp2_status_string(rc).decode()
    if rc == ERR_CUDA:
        msg += ': ' + lib.dgpmp2_last_cuda_error().decode()
    raise Dgpmp2Error('dgpmp2_b200: %s (code %d)' % (msg, rc))


def require_cuda():
    if not torch.cuda.is_available():
        raise Dgpmp2Error('dgpmp2_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')


def _as_float(x):
    if isinstance(x, torch.Tensor):
        return float(x.detach().double().reshape(-1)[0].item()) if x.numel() == 1 else x
    return float(x)


def make_params(B, T, dof, H, W, x_lims, y_lims, total_time_sec, r_sphere, K_s, K_g, reg,
                Q_c_inv, cost_sigma, epsilon_dist, non_holonomic=False, K_d=None,
                use_vel_limits=False, K_v=None, v_x=None, v_y=None, q_full=False,
                sdf_stride_b=None, Q_c_inv_static=None, w_obs_static=None, eps_static=None) -> CParams:
    """Fill struct dgpmp2_params from planner constants, computing the derived scalars in
    double precision the way the reference computes them (file:line in the header)."""
    p = CParams()
    p.x_hi, p.y_hi = float(x_lims[1]), float(y_lims[1])     # python-side only (res depends on the SDF width)
    p.B, p.T, p.dof, p.H, p.W = int(B), int(T), int(dof), int(H), int(W)
    flags = 0
    if non_holonomic:
        flags |= FLAG_NONHOLONOMIC
    if use_vel_limits:
        flags |= FLAG_VEL_LIMITS
    if q_full:
        flags |= FLAG_Q_FULL
    p.flags = flags
    p.sdf_stride_b = int(H * W if sdf_stride_b is None else sdf_stride_b)
    p.x_lo, p.y_lo = float(x_lims[0]), float(y_lims[0])
    p.res = (float(x_lims[1]) - float(x_lims[0])) / (W)                # obstacle_cost.py:34
    total_time_step = T - 1
    p.dt = _as_float(total_time_sec) * 1.0 / total_time_step * 1.0      # plan_layer.py:31
    p.r_sphere = _as_float(r_sphere)
    p.ks_inv2 = 1.0 / math.pow(_as_float(K_s), 2.0)                     # plan_layer.py:64
    p.kg_inv2 = 1.0 / math.pow(_as_float(K_g), 2.0)                     # plan_layer.py:65
    p.reg = _as_float(reg)
    p.kd_inv2 = 1.0 / math.pow(_as_float(K_d), 2.0) if (non_holonomic and K_d is not None) else 0.0
    p.kv_inv2 = 1.0 / math.pow(_as_float(K_v), 2.0) if (use_vel_limits and K_v is not None) else 0.0
    p.vx_lim = _as_float(v_x) if v_x is not None else 0.0
    p.vy_lim = _as_float(v_y) if v_y is not None else 0.0
    qfix = torch.as_tensor(Q_c_inv).detach().double().cpu().reshape(-1)
    if qfix.numel() != dof * dof:
        raise ValueError('Q_c_inv must be dof x dof')
    qst = qfix if Q_c_inv_static is None else torch.as_tensor(Q_c_inv_static).detach().double().cpu().reshape(-1)
    for i in range(dof * dof):
        p.qc_inv_fix[i] = float(qfix[i])
        p.qc_inv[i] = float(qst[i])
    p.w_obs_fix = 1.0 / math.pow(_as_float(cost_sigma), 2.0)            # plan_layer.py:71-76
    p.w_obs = p.w_obs_fix if w_obs_static is None else _as_float(w_obs_static)
    p.eps = _as_float(epsilon_dist) if eps_static is None else _as_float(eps_static)
    return p


def set_sdf_shape(p: CParams, H: int, W: int, stride_b: int):
    """Record the SDF geometry of this call; the cell size follows the SDF WIDTH (obstacle_cost.py:34)."""
    p.H, p.W, p.sdf_stride_b = int(H), int(W), int(stride_b)
    p.res = (p.x_hi - p.x_lo) / (int(W))


def num_factor_rows(p: CParams) -> int:
    """M of plan_layer.py:39-45."""
    d = 2 * p.dof
    m = d * ((p.T - 1) + 2) + p.T
    if p.flags & FLAG_NONHOLONOMIC:
        m += p.T
    if p.flags & FLAG_VEL_LIMITS:
        m += p.dof * p.T
    return m


def suffix(dtype) -> str:
    if dtype == torch.float32:
        return 'f32'
    if dtype == torch.float64:
        return 'f64'
    raise TypeError('dgpmp2_b200 supports float32 and float64 tensors, got %s' % (dtype,))


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def make_weights(qc_inv: Optional[torch.Tensor], w_obs: Optional[torch.Tensor], eps: Optional[torch.Tensor],
                 B: int, T: int, blk: int):
    """struct dgpmp2_weights from (possibly expanded / strided) torch CUDA tensors.

    qc_inv: (B|1, T-1|1, blk, blk) ; w_obs, eps: anything reshapeable to (B|1, T|1) leading dims
    (e.g. the reference's (B,T,1,1)).  Broadcast dims get stride 0; no copies for expanded views.
    Returns (CWeights, keepalive list).
    """
    w = CWeights()
    keep = []

    def lead_strides(t, inner):
        # t has shape (b, n, *inner); the inner block must be contiguous
        tail = 1
        for k in range(t.dim() - 1, 1, -1):
            if t.shape[k] != 1 and t.stride(k) != tail:
                return None
            tail *= t.shape[k]
        sb = 0 if t.shape[0] == 1 else t.stride(0)
        st = 0 if t.shape[1] == 1 else t.stride(1)
        return sb, st

    if qc_inv is not None:
        q = qc_inv
        if q.dim() == 3:
            q = q.unsqueeze(0)
        if q.dim() != 4 or q.shape[-1] != blk or q.shape[-2] != blk or q.shape[0] not in (1, B) or q.shape[1] not in (1, T - 1):
            raise ValueError('qc_inv must be (B,T-1,%d,%d), got %s' % (blk, blk, tuple(qc_inv.shape)))
        s = lead_strides(q, blk * blk)
        if s is None:
            q = q.contiguous()
            s = lead_strides(q, blk * blk)
        keep.append(q)
        w.qc_inv, w.qc_stride_b, w.qc_stride_t = q.data_ptr(), s[0], s[1]

    def scalar_field(t, name):
        if t.dim() < 2:
            raise ValueError('%s must have leading dims (B,T)' % name)
        if t.shape[0] not in (1, B) or t.shape[1] not in (1, T) or t.numel() != t.shape[0] * t.shape[1]:
            raise ValueError('%s must be (B,T,1,1)-like, got %s' % (name, tuple(t.shape)))
        keep.append(t)
        sb = 0 if t.shape[0] == 1 else t.stride(0)
        st = 0 if t.shape[1] == 1 else t.stride(1)
        return t.data_ptr(), sb, st

    if w_obs is not None:
        w.w_obs, w.w_stride_b, w.w_stride_t = scalar_field(w_obs, 'obscov_inv')
    if eps is not None:
        w.eps, w.eps_stride_b, w.eps_stride_t = scalar_field(eps, 'eps')
    return w, keep


def head_block(mode: str, dof: int) -> int:
    """Raw values per GP factor in the learned module's output (diff_gpmp2_planner.py:250-278)."""
    if mode not in HEAD_MODES:
        raise NotImplementedError('dynamics_mode %r' % (mode,))
    return {'fix_dynamics': 0, 'diag_identity': 1, 'qc_full': dof, 'q_full': 2 * dof}[mode]


def set_head_flags(p: CParams, mode: str):
    """Select the fused covariance head of the kernels for this call (DGPMP2_FLAG_HEAD*)."""
    head_block(mode, p.dof)
    if (mode == 'q_full') != bool(p.flags & FLAG_Q_FULL):
        raise ValueError("dynamics_mode 'q_full' needs params built with q_full=True (and only that mode does)")
    p.flags |= FLAG_HEAD
    if mode == 'qc_full':
        p.flags |= FLAG_HEAD_QC_VEC


def make_head_weights(q_raw: Optional[torch.Tensor], o_raw: Optional[torch.Tensor], e_raw: Optional[torch.Tensor],
                      B: int, T: int, n: int):
    """struct dgpmp2_weights over RAW head outputs (views of the learned module's ``out``; no copies):
    q_raw (B|1, T-1|1, n) with n values per GP factor, o_raw / e_raw (B|1, T|1)."""
    w = CWeights()
    keep = []
    if q_raw is not None:
        q = q_raw
        if q.dim() != 3 or q.shape[2] != n or q.shape[0] not in (1, B) or q.shape[1] not in (1, T - 1):
            raise ValueError('raw Qc head output must be (B,T-1,%d), got %s' % (n, tuple(q_raw.shape)))
        if n > 1 and q.stride(2) != 1:
            q = q.contiguous()
        keep.append(q)
        w.qc_inv = q.data_ptr()
        w.qc_stride_b = 0 if q.shape[0] == 1 else q.stride(0)
        w.qc_stride_t = 0 if q.shape[1] == 1 else q.stride(1)
    for name, t in (('w_obs', o_raw), ('eps', e_raw)):
        if t is None:
            continue
        if t.dim() != 2 or t.shape[0] not in (1, B) or t.shape[1] not in (1, T):
            raise ValueError('raw %s head output must be (B,T), got %s' % (name, tuple(t.shape)))
        keep.append(t)
        sb = 0 if t.shape[0] == 1 else t.stride(0)
        st = 0 if t.shape[1] == 1 else t.stride(1)
        if name == 'w_obs':
            w.w_obs, w.w_stride_b, w.w_stride_t = t.data_ptr(), sb, st
        else:
            w.eps, w.eps_stride_b, w.eps_stride_t = t.data_ptr(), sb, st
    return w, keep
