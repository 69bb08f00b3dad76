"""TEST INFRASTRUCTURE ONLY -- CPU fp64 restatement of dGPMP2's inner Gauss-Newton step.

This module restates, in plain torch-CPU float64 tensor ops, the algorithm the
reference runs for ``PlanLayer.forward`` (reference
``diff_gpmp2/gpmp2/plan_layer.py:87-99``): evaluate every factor, place the
Jacobians / errors / inverse covariances in DENSE ``A (B,M,N)``, ``b (B,M,1)``,
``K (B,M,M)``, form ``A^T K A + reg*I`` and ``A^T K b`` with dense batched
matmuls and obtain ``dtheta`` with a dense Cholesky followed by two explicit
triangular inverses -- exactly the reference's (deliberately naive) algorithm,
so that timing it is a fair statement of the reference CPU path ("port").

Parity pinning: the reference ships no golden vectors, so
this oracle is pinned against OUTPUTS OF THE LIVE REFERENCE generated in the
build container by ``oracle/make_golden.py`` (committed under
``tests/golden/``; checked by ``tests/test_oracle_golden.py``).  The custom
factors (nonholonomic, velocity limit) cannot be executed batched by the
reference; for those the per-trajectory factor functions
are pinned against the live reference's per-trajectory output and the batched
system is "restated-oracle parity" only.

Nothing in the product package imports this file.

Every function cites the reference lines it follows.  All inputs are promoted
to float64 (the reference's examples run with DoubleTensor defaults).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import torch

F64 = torch.float64


@dataclass
class GNParams:
    """Constructor-time constants of the planner (reference ``plan_layer.py:14-85``)."""
    dof: int
    T: int                       # num_traj_states = total_time_step + 1   (plan_layer.py:29-30)
    total_time_sec: float
    x_lims: Sequence[float]
    y_lims: Sequence[float]
    r_sphere: float              # robot_model.get_sphere_radii()           (obstacle_factor.py:37)
    K_s: float
    K_g: float
    reg: float                   # optim_params['reg']                      (plan_layer.py:96)
    Q_c_inv: Sequence[Sequence[float]]   # fixed Qc^-1 for err_ext          (plan_layer.py:70-73)
    cost_sigma: float            # fixed obstacle sigma for err_ext         (plan_layer.py:71-76)
    epsilon_dist: float
    non_holonomic: bool = False
    K_d: float = 0.01
    use_vel_limits: bool = False
    K_v: float = 0.01
    v_x: float = 1.0
    v_y: float = 1.0
    nlinks: int = 1

    @property
    def d(self) -> int:
        return 2 * self.dof

    @property
    def dt(self) -> float:
        # plan_layer.py:31  dt = total_time_sec / total_time_step
        return self.total_time_sec * 1.0 / (self.T - 1) * 1.0

    @property
    def M(self) -> int:
        # plan_layer.py:39-45
        m = self.d * ((self.T - 1) + 2) + self.T * self.nlinks
        if self.non_holonomic:
            m += self.T
        if self.use_vel_limits:
            m += self.dof * self.T
        return m

    @property
    def N(self) -> int:
        return self.d * self.T   # plan_layer.py:46


def _f64(x) -> torch.Tensor:
    return torch.as_tensor(x).to(F64)


# --------------------------------------------------------------------------
# SDF lookup + hinge (reference utils/sdf_utils.py:38-107, obstacle_cost.py:29-38)
# --------------------------------------------------------------------------
def bilinear_sdf(sdf: torch.Tensor, pts: torch.Tensor, res: float,
                 x_lims: Sequence[float], y_lims: Sequence[float]
                 ) -> Tuple[torch.Tensor, torch.Tensor]:
    """``sdf (B,H,W)``, ``pts (B,N,2)`` -> ``dist (B,N,1)``, ``J (B,N,2)``.

    Follows sdf_utils.py:57-94.  The in-limits mask of :96-106 is a no-op under
    bool tensor semantics (``a + b`` on bools is OR, ``== 1`` keeps it), so the
    out-of-image behaviour is the clamp artefact: both clamped indices coincide,
    the weights cancel and ``dist = 0``, ``J = 0``.  (Pinned by golden case
    "oob".)
    """
    sdf = _f64(sdf)
    pts = _f64(pts)
    B, H, W = sdf.shape
    N = pts.shape[1]
    orig_x = 0.0 - x_lims[0] / res                       # :57
    orig_y = 0.0 - y_lims[0] / res                       # :58
    px = (orig_x + pts[:, :, 0] / res).reshape(-1)       # :61
    py = (orig_y - pts[:, :, 1] / res).reshape(-1)       # :62  (y axis flipped)
    x1 = torch.floor(px).long()                          # :64-67
    x2 = x1 + 1
    y1 = torch.floor(py).long()
    y2 = y1 + 1
    x1 = x1.clamp(0, W - 1)                              # :69-72
    x2 = x2.clamp(0, W - 1)
    y1 = y1.clamp(0, H - 1)
    y2 = y2.clamp(0, H - 1)
    bz = torch.arange(B).repeat_interleave(N)            # :73-74
    v11 = sdf[bz, y1, x1]                                # :76-79
    v21 = sdf[bz, y1, x2]
    v12 = sdf[bz, y2, x1]
    v22 = sdf[bz, y2, x2]
    ax = x2.to(F64) - px                                 # weights use the CLAMPED indices (:81-89)
    bx = px - x1.to(F64)
    ay = y2.to(F64) - py
    by = py - y1.to(F64)
    dist = ax * ay * v11 + bx * ay * v21 + ax * by * v12 + bx * by * v22   # :90
    Jx = -1.0 * (ay * (v21 - v11) + by * (v22 - v12)) / res               # :93
    Jy = (ax * (v12 - v11) + bx * (v22 - v21)) / res                      # :94
    J = torch.stack((Jx, Jy), dim=-1).reshape(B, N, 2)
    return dist.reshape(B, N, 1), J


def sdf_resolution(p: GNParams, sdf_width: int) -> float:
    """obstacle_cost.py:34 -- derived from the SDF *width* only, used for both axes."""
    return (p.x_lims[1] - p.x_lims[0]) / sdf_width


def obstacle_factor(th: torch.Tensor, sdf: torch.Tensor, eps: torch.Tensor, p: GNParams
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """``ObstacleFactor.get_error`` (obstacle_factor.py:35-40) for a one-sphere robot.

    th (B,T,d), sdf (B,H,W) or (B,1,H,W), eps broadcastable to (B,T,1,1)
    -> cost (B,T,1,1), H (B,T,1,d).  Sphere centre = first two state entries and
    J_fk = [I2 0] for both PointRobot2D (point_robot_2d.py:58-63) and
    PointRobotXYH (point_robot_xyh.py:20-38).
    """
    th = _f64(th)
    sdf = _f64(sdf)
    if sdf.dim() == 4:
        sdf = sdf[:, 0]
    B, T, d = th.shape
    eps_tot = (_f64(eps) + p.r_sphere).expand(B, T, 1, 1).reshape(B, T, 1)   # obstacle_cost.py:30,33
    res = sdf_resolution(p, sdf.shape[-1])
    dist, J = bilinear_sdf(sdf, th[:, :, 0:2], res, p.x_lims, p.y_lims)
    active = dist <= eps_tot                                                  # :36 (<=)
    cost = torch.where(active, eps_tot - dist, torch.zeros_like(dist))
    He = torch.where(active, -1.0 * J, torch.zeros_like(J))                  # :37
    H = torch.zeros(B, T, 1, d, dtype=F64)
    H[:, :, 0, 0:2] = He                                                      # H_e . J_fk  (obstacle_factor.py:39)
    return cost.reshape(B, T, 1, 1), H


# --------------------------------------------------------------------------
# GP prior + start/goal priors (gp/gp_factor.py, gp/prior_factor.py)
# --------------------------------------------------------------------------
def gp_phi(dof: int, dt: float) -> torch.Tensor:
    """gp_factor.py:31-37."""
    I = torch.eye(dof, dtype=F64)
    Z = torch.zeros(dof, dof, dtype=F64)
    return torch.cat((torch.cat((I, dt * I), dim=1), torch.cat((Z, I), dim=1)), dim=0)


def gp_inv_cov(qc_inv: torch.Tensor, dt: float) -> torch.Tensor:
    """gp_factor.py:65-73: Q^-1 = [[12 dt^-3, -6 dt^-2],[-6 dt^-2, 4 dt^-1]] (x) Qc^-1."""
    qc_inv = _f64(qc_inv)
    m1 = 12.0 * (dt ** -3.0) * qc_inv
    m2 = -6.0 * (dt ** -2.0) * qc_inv
    m3 = 4.0 * (dt ** -1.0) * qc_inv
    up = torch.cat((m1, m2), dim=-1)
    lo = torch.cat((m2, m3), dim=-1)
    return torch.cat((up, lo), dim=-2)


def gp_factor(th: torch.Tensor, p: GNParams):
    """``GPFactor.get_error`` (gp_factor.py:100-110): e_i = th_{i+1} - Phi th_i, H1 = Phi, H2 = -I."""
    th = _f64(th)
    B, T, d = th.shape
    phi = gp_phi(p.dof, p.dt)
    s1 = th[:, :-1]
    s2 = th[:, 1:]
    e = s2 - torch.einsum('ij,btj->bti', phi, s1)
    H1 = phi.expand(B, T - 1, d, d)
    H2 = (-1.0 * torch.eye(d, dtype=F64)).expand(B, T - 1, d, d)
    return e.unsqueeze(-1), H1, H2


def prior_factor(state: torch.Tensor, mean: torch.Tensor):
    """``PriorFactor.get_error`` (prior_factor.py:15-18): e = mean - state, H = I."""
    state = _f64(state)
    mean = _f64(mean)
    B = state.shape[0]
    d = state.shape[-1]
    e = (mean - state).reshape(B, d, 1)
    H = torch.eye(d, dtype=F64).expand(B, d, d)
    return e, H


# --------------------------------------------------------------------------
# custom factors, literal per-trajectory restatements
# --------------------------------------------------------------------------
def nonholonomic_factor_traj(traj: torch.Tensor):
    """``NonHolonomicFactor.get_error_full`` on ONE (T,6) trajectory
    (nonholonomic_factor.py:16-30), state = (x, y, h, vx, vy, w).
    The Jacobian row is reproduced literally."""
    traj = _f64(traj)
    vx = traj[:, 3:4]
    vy = traj[:, 4:5]
    h = traj[:, 2:3]
    err = vy * torch.cos(h) - vx * torch.sin(h)                          # :20
    h1 = torch.zeros(traj.shape[0], 2, dtype=F64)                        # :22
    h2 = -vy * torch.sin(h) + vx * torch.cos(h)                          # :23
    h3 = torch.cat((-torch.sin(h), torch.cos(h)), -1)                    # :24
    h4 = torch.zeros(traj.shape[0], 1, dtype=F64)                        # :25
    H = torch.cat((torch.cat((h1, h2), -1), torch.cat((h3, h4), -1)), -1)  # :26-29
    return err, H


def velocity_limit_factor_traj(traj: torch.Tensor, vx_lim: float, vy_lim: float):
    """``VelocityLimitFactor.get_error_full`` on ONE (T,4) trajectory
    (velocity_limit_factor.py:17-29) with integer ``ndims//2``.
    cost (T,2), H (T,2,4); active when |v| >= limit (note: >=)."""
    traj = _f64(traj)
    T = traj.shape[0]
    vx = traj[:, 2:3]
    vy = traj[:, 3:4]
    ax = torch.abs(vx) >= vx_lim
    ay = torch.abs(vy) >= vy_lim
    cost_x = torch.where(ax, torch.abs(vx) - vx_lim, torch.zeros_like(vx))
    cost_y = torch.where(ay, torch.abs(vy) - vy_lim, torch.zeros_like(vy))
    H = torch.zeros(T, 2, 4, dtype=F64)
    H[:, 0, 2] = torch.where(ax, -torch.sign(vx), torch.zeros_like(vx))[:, 0]
    H[:, 1, 3] = torch.where(ay, -torch.sign(vy), torch.zeros_like(vy))[:, 0]
    return torch.cat((cost_x, cost_y), dim=1), H


def nonholonomic_factor(th: torch.Tensor):
    """Per-batch-element application of the per-trajectory factor."""
    outs = [nonholonomic_factor_traj(t) for t in _f64(th)]
    return torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs])


def velocity_limit_factor(th: torch.Tensor, vx_lim: float, vy_lim: float):
    outs = [velocity_limit_factor_traj(t, vx_lim, vy_lim) for t in _f64(th)]
    return torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs])


# --------------------------------------------------------------------------
# dense linear system (plan_layer.py:152-200 with the row layout of :391-451)
# --------------------------------------------------------------------------
def _expand_w(w, B, T) -> torch.Tensor:
    return _f64(w).expand(B, T, 1, 1).reshape(B, T)


def build_dense_system(th, start, goal, sdf, q_inv, w_obs, eps, p: GNParams):
    """Dense A (B,M,N), b (B,M,1), K (B,M,M).

    Row layout (plan_layer.py:417-451): start prior rows [0,d); GP factor i rows
    [(i+1)d,(i+2)d) with H1 in column block i and H2 in column block i+1; goal
    prior rows [T d,(T+1) d) in the last column block; obstacle factor t at row
    (T+1) d + t in column block t; then nonholonomic rows (one per state) or
    velocity-limit rows (dof per state) at offset (T+1) d + T (the reference
    does not advance the offset between the two, :442, so they are exclusive).
    q_inv is the full (B,T-1,d,d) GP inverse covariance.
    """
    th = _f64(th)
    B, T, d = th.shape
    assert T == p.T and d == p.d
    M, N = p.M, p.N
    A = torch.zeros(B, M, N, dtype=F64)
    b = torch.zeros(B, M, 1, dtype=F64)
    K = torch.zeros(B, M, M, dtype=F64)
    Id = torch.eye(d, dtype=F64)

    # start prior (:157-158,:169-171); K = I / K_s^2 (:64,:67)
    e_s, H_s = prior_factor(th[:, 0], _f64(start).reshape(B, d))
    A[:, 0:d, 0:d] = H_s
    b[:, 0:d] = e_s
    K[:, 0:d, 0:d] = Id * (1.0 / math.pow(p.K_s, 2.0))

    # GP factors (:160-161,:173-176)
    e_gp, H1, H2 = gp_factor(th, p)
    q_inv = _f64(q_inv).expand(B, T - 1, d, d)
    for i in range(T - 1):
        r0 = (i + 1) * d
        A[:, r0:r0 + d, i * d:(i + 1) * d] = H1[:, i]
        A[:, r0:r0 + d, (i + 1) * d:(i + 2) * d] = H2[:, i]
        b[:, r0:r0 + d] = e_gp[:, i]
        K[:, r0:r0 + d, r0:r0 + d] = q_inv[:, i]

    # goal prior (:163-164,:178-180)
    off = d * T
    e_g, H_g = prior_factor(th[:, T - 1], _f64(goal).reshape(B, d))
    A[:, off:off + d, N - d:N] = H_g
    b[:, off:off + d] = e_g
    K[:, off:off + d, off:off + d] = Id * (1.0 / math.pow(p.K_g, 2.0))

    # obstacle factors (:166-167,:182-184), nlinks == 1
    off = off + d
    c_obs, H_obs = obstacle_factor(th, sdf, eps, p)
    w = _expand_w(w_obs, B, T)
    for t in range(T):
        A[:, off + t, t * d:(t + 1) * d] = H_obs[:, t, 0]
        b[:, off + t, 0] = c_obs[:, t, 0, 0]
        K[:, off + t, off + t] = w[:, t]
    off = off + T

    if p.non_holonomic:                                   # :186-191, rows :433-441
        e_nh, H_nh = nonholonomic_factor(th)
        kd = 1.0 / math.pow(p.K_d, 2.0)                   # nonholonomic_factor.py:14
        for t in range(T):
            A[:, off + t, t * d:(t + 1) * d] = H_nh[:, t]
            b[:, off + t, 0] = e_nh[:, t, 0]
            K[:, off + t, off + t] = kd
    if p.use_vel_limits:                                  # :193-198, rows :443-451
        c_v, H_v = velocity_limit_factor(th, p.v_x, p.v_y)
        kv = 1.0 / math.pow(p.K_v, 2.0)                   # velocity_limit_factor.py:15
        for t in range(T):
            r0 = off + t * p.dof
            A[:, r0:r0 + p.dof, t * d:(t + 1) * d] = H_v[:, t]
            b[:, r0:r0 + p.dof, 0] = c_v[:, t]
            for k in range(p.dof):
                K[:, r0 + k, r0 + k] = kv
    return A, b, K


def normal_equations(A, b, K, reg: float):
    """plan_layer.py:215-220."""
    AtK = torch.bmm(A.transpose(1, 2), K)
    LAM = torch.bmm(AtK, A) + reg * torch.eye(A.shape[-1], dtype=F64)
    R = torch.bmm(AtK, b)
    return LAM, R


def solve_dense(A, b, K, reg: float) -> torch.Tensor:
    """plan_layer.py:214-234: upper Cholesky, then the two explicit N x N inverses of the triangular
    factor applied with bmm.  The reference forms the inverses with ``torch.inverse`` (LU); here the
    same inverses are formed with a triangular solve against the identity, because MKL's batched
    getrf/getri path hangs on the GPU box's host CPU ("oneMKL ERROR: Parameter 6 ... DLASWP").
    Same arithmetic structure (explicit inverse + dense bmm), same result to rounding."""
    LAM, R = normal_equations(A, b, K, reg)
    u = torch.linalg.cholesky(LAM).transpose(1, 2).contiguous()   # torch.cholesky(LAM, upper=True)
    eye = torch.eye(LAM.shape[-1], dtype=F64).expand_as(LAM)
    ut_inv = torch.linalg.solve_triangular(u.transpose(1, 2), eye, upper=False)
    u_inv = torch.linalg.solve_triangular(u, eye, upper=True)
    z = torch.bmm(ut_inv, R)
    dth = torch.bmm(u_inv, z)
    return dth


def weighted_error(th, start, goal, sdf, q_inv, w_obs, eps, p: GNParams) -> torch.Tensor:
    """plan_layer.py:273-308 (and :310-345 when called with the fixed covariances):
    err = 0.5 * sum_f e_f^T K_f e_f / M  -> (B,1,1)."""
    th = _f64(th)
    B, T, d = th.shape
    e_s, _ = prior_factor(th[:, 0], _f64(start).reshape(B, d))
    e_g, _ = prior_factor(th[:, T - 1], _f64(goal).reshape(B, d))
    err = 0.5 * (e_s.squeeze(-1) ** 2).sum(-1) / math.pow(p.K_s, 2.0)
    err = err + 0.5 * (e_g.squeeze(-1) ** 2).sum(-1) / math.pow(p.K_g, 2.0)
    e_gp, _, _ = gp_factor(th, p)
    q_inv = _f64(q_inv).expand(B, T - 1, d, d)
    err = err + 0.5 * torch.einsum('bti,btij,btj->b', e_gp.squeeze(-1), q_inv, e_gp.squeeze(-1))
    c_obs, _ = obstacle_factor(th, sdf, eps, p)
    w = _expand_w(w_obs, B, T)
    err = err + 0.5 * (w * c_obs.reshape(B, T) ** 2).sum(-1)
    if p.non_holonomic:
        e_nh, _ = nonholonomic_factor(th)
        err = err + 0.5 * (e_nh.reshape(B, T) ** 2).sum(-1) / math.pow(p.K_d, 2.0)
    if p.use_vel_limits:
        c_v, _ = velocity_limit_factor(th, p.v_x, p.v_y)
        err = err + 0.5 * (c_v.reshape(B, -1) ** 2).sum(-1) / math.pow(p.K_v, 2.0)
    return (err / p.M).reshape(B, 1, 1)


def fixed_covariances(p: GNParams, B: int):
    """Constructor-time covariances used by ``error_ext_batch`` (plan_layer.py:70-81)."""
    qc = _f64(p.Q_c_inv).expand(B, p.T - 1, p.dof, p.dof)
    w = torch.full((B, p.T, 1, 1), 1.0 / math.pow(p.cost_sigma, 2.0), dtype=F64)
    return gp_inv_cov(qc, p.dt), w


def covariances_from_head(out, p: GNParams, mode: str = 'diag_identity', learn_eps: bool = False):
    """``DiffGPMP2Planner.get_covariances`` (diff_gpmp2_planner.py:247-283): the learned module's output
    ``out`` (B,1,out_dim) = [q | o | e] -> (qc_inv or None, obscov_inv (B,T,1,1), eps (B,T,1,1) or None).

    'fix_dynamics': no q part (:250-253); 'diag_identity': one value per GP factor, Qc^-1 = q*q * I_dof
    (:254-260); 'qc_full': dof values, Qc^-1 = v v^T (:267-271); 'q_full': state_dim values, Q^-1 = v v^T
    (:272-276); obscov_inv = o*o (:278); eps = e*e (:280-282).  nlinks = 1.  Pinned against the live
    reference by tests/golden/head_*.npz (oracle/make_golden_head.py).
    """
    out = _f64(out)
    B = out.shape[0]
    G, S, d = p.T - 1, p.T, 2 * p.dof
    n = {'fix_dynamics': 0, 'diag_identity': 1, 'qc_full': p.dof, 'q_full': d}[mode]
    row = out[:, 0, :]
    qc = None
    if n:
        v = row[:, :G * n].reshape(B, G, n, 1)
        qc = v * v.transpose(2, 3)
        if mode == 'diag_identity':
            qc = qc * torch.eye(p.dof, dtype=F64)
    o = row[:, G * n:G * n + S].reshape(B, S, 1, 1)
    w = o * o
    eps = None
    if learn_eps:
        e = row[:, G * n + S:].reshape(B, S, 1, 1)
        eps = e * e
    return qc, w, eps


def gn_step(th, start, goal, sdf, qc_inv, w_obs, eps, p: GNParams, q_full: bool = False):
    """``PlanLayer.forward`` (plan_layer.py:87-99) -> (dtheta (B,T,d), err (B,1,1), err_ext (B,1,1)).

    ``qc_inv`` is (B,T-1,dof,dof) (or broadcastable) unless ``q_full`` in which
    case it is the full (B,T-1,d,d) GP inverse covariance (:90).
    """
    th = _f64(th)
    B, T, d = th.shape
    q_inv = _f64(qc_inv) if q_full else gp_inv_cov(_f64(qc_inv), p.dt)
    A, b, K = build_dense_system(th, start, goal, sdf, q_inv, w_obs, eps, p)
    dth = solve_dense(A, b, K, p.reg).reshape(B, T, d)
    err = weighted_error(th, start, goal, sdf, q_inv, w_obs, eps, p)
    q_fix, w_fix = fixed_covariances(p, B)
    err_ext = weighted_error(th, start, goal, sdf, q_fix, w_fix, eps, p)
    return dth, err, err_ext


def unweighted_errors(th, start, goal, sdf, eps, p: GNParams):
    """plan_layer.py:374-388 -> err_sg (B,1), err_gp (B,1,1), err_obs (B,1,1)."""
    th = _f64(th)
    B, T, d = th.shape
    e_s, _ = prior_factor(th[:, 0], _f64(start).reshape(B, d))
    e_g, _ = prior_factor(th[:, T - 1], _f64(goal).reshape(B, d))
    err_sg = (0.5 * (e_s ** 2).sum(1) + 0.5 * (e_g ** 2).sum(1)).reshape(B, 1)
    e_gp, _, _ = gp_factor(th, p)
    err_gp = (0.5 * (e_gp.squeeze(-1) ** 2).sum(-1)).mean(dim=1).reshape(B, 1, 1)
    c_obs, _ = obstacle_factor(th, sdf, eps, p)
    err_obs = (0.5 * c_obs.reshape(B, T) ** 2).mean(dim=1).reshape(B, 1, 1)
    return err_sg, err_gp, err_obs


def band_from_dense(LAM: torch.Tensor, R: torch.Tensor, T: int, d: int):
    """Extract the block-tridiagonal band (D (B,T,d,d), U (B,T-1,d,d), r (B,T,d)) of a
    dense information matrix and return the largest |entry| outside the band."""
    B = LAM.shape[0]
    D = torch.stack([LAM[:, t * d:(t + 1) * d, t * d:(t + 1) * d] for t in range(T)], dim=1)
    U = torch.stack([LAM[:, t * d:(t + 1) * d, (t + 1) * d:(t + 2) * d] for t in range(T - 1)], dim=1)
    mask = torch.ones_like(LAM, dtype=torch.bool)
    for t in range(T):
        lo = max(0, (t - 1) * d)
        hi = min(T * d, (t + 2) * d)
        mask[:, t * d:(t + 1) * d, lo:hi] = False
    off = LAM[mask].abs().max().item() if mask.any() else 0.0
    return D, U, R.reshape(B, T, d), off


def gn_solve(th0, start, goal, sdf, qc_inv, w_obs, eps, p: GNParams,
             max_iters: int, tol_delta: float, q_full: bool = False):
    """``DiffGPMP2Planner.forward`` (diff_gpmp2_planner.py:92-174) with static covariances:
    per problem, iterate th <- th + dtheta until ||dtheta||_2 < tol_delta or j >= max_iters
    (planner_utils.py:3-16).  Returns th_final (B,T,d), err_init[B], err_final[B],
    err_per_iter[B][j], err_ext_per_iter[B][j], iters[B]."""
    th0 = _f64(th0)
    B = th0.shape[0]
    out_th = torch.zeros_like(th0)
    err_init, err_final, err_pi, err_ext_pi, iters = [], [], [], [], []
    for i in range(B):
        sl = slice(i, i + 1)

        def pick(w):
            w = _f64(w)
            return w[sl] if (w.dim() >= 1 and w.shape[0] == B and w.dim() == 4) else w
        th = th0[sl].clone()
        epi, eepi = [], []
        j = 0
        while True:
            dth, e_old, ee_old = gn_step(th, _f64(start)[sl], _f64(goal)[sl], _f64(sdf)[sl],
                                         pick(qc_inv), pick(w_obs), pick(eps), p, q_full)
            epi.append(e_old.item())
            eepi.append(ee_old.item())
            th = th + dth
            j += 1
            if torch.norm(dth) < tol_delta or j >= max_iters:
                break
        q_inv = pick(qc_inv) if q_full else gp_inv_cov(pick(qc_inv), p.dt)
        e_new = weighted_error(th, _f64(start)[sl], _f64(goal)[sl], _f64(sdf)[sl], q_inv, pick(w_obs), pick(eps), p)
        out_th[i] = th[0]
        err_init.append(epi[0])
        err_final.append(e_new.item())
        err_pi.append(epi)
        err_ext_pi.append(eepi)
        iters.append(j)
    return out_th, err_init, err_final, err_pi, err_ext_pi, iters
