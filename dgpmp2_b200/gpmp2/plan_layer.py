"""One batched Gauss-Newton step as a torch module (API mirror of reference
``diff_gpmp2/gpmp2/plan_layer.py:13-99``).

Same constructor, same ``forward(thb, startb, goalb, imb, sdfb, qc_inv_trajb, obscov_inv_trajb,
eps_trajb) -> (dthetab, err, err_ext)``, same factor attributes (``start_prior``, ``goal_prior``,
``gp_prior``, ``obs_factor``, ``gp_prior_fix``, ``obs_factor_fix``, ``dyn_factor``, ``vel_factor``)
and the same statefulness: ``forward`` installs the per-call means / covariances / eps on the
factor objects and ``error_batch`` / ``error_ext_batch`` evaluate against what was last installed.

What differs is everything underneath: the reference builds dense A (B,M,N), b, K with
masked_scatter_ and solves dense normal equations (:152-234); here one fused CUDA kernel
(dgpmp2_gn_step_*) evaluates the factors, keeps the block-tridiagonal band in shared memory,
solves it by block cyclic reduction and returns dtheta and both errors.  ``batch_size`` is
accepted for compatibility but any batch size works.  There is no CPU path: CPU tensors are
staged to the GPU and the results handed back on the caller's device.
"""
import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from .. import _lib, ops
from .._dev import as_float, back, to_cuda, work_dtype
from .custom_factors import NonHolonomicFactor, VelocityLimitFactor
from .gp import GPFactor, PriorFactor
from .obstacle import ObstacleFactor


class _GNStep(torch.autograd.Function):
    """dtheta, err, err_ext = GN step, differentiable w.r.t. th, start, goal, sdf, qc_inv, obscov_inv, eps.
    Forward: dgpmp2_gn_step_*; backward: dgpmp2_gn_step_backward_* (adjoint solve with the same block
    cyclic reduction + factor VJPs).  err is not differentiable (the reference computes it under no_grad).
    With ``head`` (a dynamics_mode string) qc / w / eps are the RAW outputs of the learned module
    (q (B,T-1,n), o (B,T), e (B,T); any may be None) and the kernels form the covariances themselves;
    the backward kernel returns the gradients w.r.t. the covariances and the chain rule through the
    products (q q^T, o^2, e^2) is applied here."""

    @staticmethod
    def forward(ctx, layer, static, head, th, start, goal, sdf, qc, w, eps):
        dt = work_dtype(th, sdf)
        p = layer.cparams(q_full=(head == 'q_full') if head is not None else None)
        c = lambda t: to_cuda(t, dt) if t is not None else None
        thc, stc, goc, sdfc = c(th), c(start), c(goal), c(sdf)
        kw = {} if static else dict(qc_inv=c(qc), w_obs=c(w), eps=c(eps), head=head)
        dth, err, err_ext, status = ops.gn_step(p, thc, stc, goc, sdfc, want_status=True, **kw)
        layer._check(status)
        ctx.layer, ctx.static, ctx.head, ctx.dt = layer, static, head, dt
        ctx.meta = [(t.device, t.dtype, tuple(t.shape)) if isinstance(t, torch.Tensor) else None for t in (th, start, goal, sdf, qc, w, eps)]
        ctx.save_for_backward(thc, stc, goc, sdfc, dth, *([] if static else [kw['qc_inv'], kw['w_obs'], kw['eps']]))
        B = th.shape[0]
        out = (back(dth, th).to(th.dtype), back(err, th).to(th.dtype).reshape(B, 1, 1), back(err_ext, th).to(th.dtype).reshape(B, 1, 1))
        ctx.mark_non_differentiable(out[1])
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g_dth, g_err, g_err_ext):
        layer, static, head, dt = ctx.layer, ctx.static, ctx.head, ctx.dt
        saved = ctx.saved_tensors
        thc, stc, goc, sdfc, dth = saved[:5]
        kw = {} if static else dict(qc_inv=saved[5], w_obs=saved[6], eps=saved[7], head=head)
        need = ctx.needs_input_grad          # (layer, static, head, th, start, goal, sdf, qc, w, eps)
        B = thc.shape[0]
        gd = to_cuda(g_dth, dt) if g_dth is not None else torch.zeros_like(dth)
        ge = to_cuda(g_err_ext, dt).reshape(B) if g_err_ext is not None else None
        p = layer.cparams(q_full=(head == 'q_full') if head is not None else None)
        outs = ops.gn_step_backward(p, thc, stc, goc, sdfc, dth, gd, ge,
                                    need_th=need[3], need_start=need[4], need_goal=need[5], need_sdf=need[6],
                                    need_qc=need[7] and not static, need_w=need[8] and not static,
                                    need_eps=need[9] and not static, **kw)
        g_th, g_start, g_goal, g_qc, g_w, g_eps, g_sdf = outs
        if head is not None:
            # chain rule through the head's products: Qc^-1 = q^2 I or v v^T, w = o^2, eps = e^2
            q, o, e = saved[5], saved[6], saved[7]
            if g_qc is not None:
                if head == 'diag_identity':
                    g_qc = 2.0 * q * torch.diagonal(g_qc, dim1=-2, dim2=-1).sum(-1, keepdim=True)
                else:
                    g_qc = torch.matmul(g_qc + g_qc.transpose(-1, -2), q.unsqueeze(-1)).squeeze(-1)
            if g_w is not None:
                g_w = 2.0 * o * g_w
            if g_eps is not None:
                g_eps = 2.0 * e * g_eps

        def fit(g, k):
            """kernel gradient (dense per (b, t)) -> the input's own shape: broadcast inputs ((1,T-1,dof,dof), (B,1,..),
            (1,T,1,1): make_weights passes them with stride 0) receive the SUM over the dimensions they were
            broadcast along, as autograd would give for an expanded tensor."""
            if g is None or ctx.meta[k] is None:
                return None
            dev, dtype, shape = ctx.meta[k]
            if k >= 4 and g.numel() != int(torch.Size(shape).numel()) and len(shape) >= 2:
                full = list(shape)
                full[0], full[1] = g.shape[0], g.shape[1]      # the kernel's gradient is dense over (problem, factor / state)
                g = g.reshape(full).sum_to_size(shape)
            return g.reshape(shape).to(device=dev, dtype=dtype)
        return (None, None, None, fit(g_th, 0), fit(g_start, 1), fit(g_goal, 2), fit(g_sdf, 3), fit(g_qc, 4), fit(g_w, 5), fit(g_eps, 6))


class _Errors(torch.autograd.Function):
    """(err, err_ext, err_sg, err_gp, err_obs) of one factor sweep (dgpmp2_errors_*), differentiable w.r.t. the
    trajectory like the reference's error_ext_batch / gp_error / obs_error / start_goal_error (plan_layer.py:310-388;
    the training loss of learning/train_planner.py:327-346 is built from them).  err is not differentiable (the
    reference computes it under no_grad, :275).  Backward: dgpmp2_errors_backward_*."""

    @staticmethod
    def forward(ctx, layer, th, sdf):
        s = layer._state
        dt = work_dtype(th, sdf)
        p, kw = layer._installed(dt)
        thc, stc, goc, sdfc = to_cuda(th, dt), to_cuda(s['start'], dt), to_cuda(s['goal'], dt), to_cuda(sdf, dt)
        outs = ops.errors(p, thc, stc, goc, sdfc, **kw)
        ctx.layer, ctx.dt, ctx.head = layer, dt, kw.get('head')
        ctx.meta = (th.device, th.dtype, tuple(th.shape))
        eps = kw.get('eps')
        ctx.has_eps = eps is not None
        ctx.save_for_backward(thc, stc, goc, sdfc, *([eps] if eps is not None else []))
        res = tuple(back(o, th).to(th.dtype) for o in outs)
        ctx.mark_non_differentiable(res[0])
        return res

    @staticmethod
    @once_differentiable
    def backward(ctx, g_err, g_ext, g_sg, g_gp, g_obs):
        saved = ctx.saved_tensors
        thc, stc, goc, sdfc = saved[:4]
        eps = saved[4] if ctx.has_eps else None
        if not ctx.needs_input_grad[1]:
            return None, None, None
        p = ctx.layer.cparams(q_full=False)
        c = lambda g: None if g is None else to_cuda(g, ctx.dt).reshape(-1)
        g_th = ops.errors_backward(p, thc, stc, goc, sdfc, c(g_ext), c(g_sg), c(g_gp), c(g_obs), eps=eps, head=ctx.head)
        dev, dtype, shape = ctx.meta
        return None, g_th.reshape(shape).to(device=dev, dtype=dtype), None


class PlanLayer(nn.Module):
    def __init__(self, gp_params, obs_params, planner_params, optim_params, env_params, robot_model,
                 learn_params=None, batch_size=1, use_cuda=False):
        super(PlanLayer, self).__init__()
        self.use_cuda = torch.cuda.is_available() if use_cuda else False
        self.device = torch.device('cuda') if self.use_cuda else torch.device('cpu')
        self.gp_params, self.obs_params, self.planner_params = gp_params, obs_params, planner_params
        self.optim_params, self.env_params, self.robot_model = optim_params, env_params, robot_model
        self.learn_params = learn_params
        self.batch_size = batch_size
        self.dof = int(planner_params['dof'])
        self.state_dim = int(planner_params['state_dim'])
        self.total_time_sec = planner_params['total_time_sec']
        self.total_time_step = int(planner_params['total_time_step'])
        self.num_traj_states = self.total_time_step + 1
        self.dt = self.total_time_sec * 1.0 / self.total_time_step * 1.0
        self.non_holonomic = bool(planner_params.get('non_holonomic', False))
        self.use_vel_limits = bool(planner_params.get('use_vel_limits', False))
        self.num_gp_factors = self.num_traj_states - 1
        self.num_prior_factors = 2
        self.num_obs_factors = self.num_traj_states
        self.nlinks = self.robot_model.nlinks
        if self.nlinks != 1:
            raise NotImplementedError('the GN kernels implement one collision sphere per state (nlinks == 1)')
        if self.state_dim != 2 * self.dof or self.dof not in (2, 3):
            raise NotImplementedError('state_dim must be 2*dof with dof in {2, 3}')
        self.M = self.state_dim * (self.num_gp_factors + self.num_prior_factors) + self.num_obs_factors * self.nlinks
        if self.non_holonomic:
            self.num_dynamics_factors = self.num_traj_states
            self.M += self.num_dynamics_factors
        if self.use_vel_limits:
            self.num_vel_factors = self.num_traj_states
            self.M += self.dof * self.num_vel_factors
        self.N = self.state_dim * self.num_traj_states
        self.dynamics_mode = learn_params['dgpmp2']['dynamics_mode'] if learn_params is not None else None
        self.q_full = self.dynamics_mode == 'q_full'

        # factor objects with the reference's names (they carry the per-call state)
        T = self.num_traj_states
        self.gp_prior = GPFactor(self.dof, self.dt, self.num_gp_factors, batch_size, self.use_cuda)
        self.start_prior = PriorFactor(self.state_dim, gp_params['K_s'], batch_size, self.use_cuda)
        self.goal_prior = PriorFactor(self.state_dim, gp_params['K_g'], batch_size, self.use_cuda)
        self.obs_factor = ObstacleFactor(self.state_dim, T, obs_params['epsilon_dist'], env_params, robot_model, batch_size, self.use_cuda)
        self.gp_prior_fix = GPFactor(self.dof, self.dt, self.num_gp_factors, batch_size, self.use_cuda)
        self.obs_factor_fix = ObstacleFactor(self.state_dim, T, obs_params['epsilon_dist'], env_params, robot_model, batch_size, self.use_cuda)
        if self.non_holonomic:
            self.dyn_factor = NonHolonomicFactor(self.dof, gp_params['K_d'], T, batch_size, self.use_cuda)
        if self.use_vel_limits:
            self.vel_factor = VelocityLimitFactor(self.state_dim, T, gp_params['K_v'], batch_size, self.use_cuda)
            self.vel_factor.set_v_traj(torch.as_tensor(gp_params['v_x']), torch.as_tensor(gp_params['v_y']))
        self.qc_inv_fix = torch.as_tensor(gp_params['Q_c_inv']).detach().double().cpu()
        self.gp_prior_fix.Q_c_inv = self.qc_inv_fix
        self.strict = True          # raise (like torch.cholesky) when a problem's system is not positive definite
        self.last_status = None
        self._state = None          # weights installed by the last forward()
        self._static_views = {}

    # ------------------------------------------------------------------ parameters
    def cparams(self, q_full=None, eps_static=None, w_static=None, qc_static=None):
        gp, ob = self.gp_params, self.obs_params
        return _lib.make_params(
            B=1, T=self.num_traj_states, dof=self.dof, H=1, W=1, x_lims=self.env_params['x_lims'],
            y_lims=self.env_params['y_lims'], total_time_sec=self.total_time_sec,
            r_sphere=self.robot_model.get_sphere_radii(), K_s=gp['K_s'], K_g=gp['K_g'], reg=self.optim_params['reg'],
            Q_c_inv=gp['Q_c_inv'], cost_sigma=ob['cost_sigma'], epsilon_dist=ob['epsilon_dist'],
            non_holonomic=self.non_holonomic, K_d=gp.get('K_d'), use_vel_limits=self.use_vel_limits, K_v=gp.get('K_v'),
            v_x=gp.get('v_x'), v_y=gp.get('v_y'), q_full=self.q_full if q_full is None else q_full,
            Q_c_inv_static=qc_static, w_obs_static=w_static, eps_static=eps_static)

    def static_weights(self, B, like):
        """The planner's constant covariances as (B, ...) expanded views (zero-copy); passing exactly
        these objects to ``forward`` selects the constant fast path of the kernel."""
        key = (B, like.device, like.dtype)
        if key not in self._static_views:
            T, dof = self.num_traj_states, self.dof
            qc = torch.as_tensor(self.gp_params['Q_c_inv']).to(like.device, like.dtype).reshape(1, 1, dof, dof).expand(B, T - 1, dof, dof)
            w = torch.full((1, 1, 1, 1), 1.0 / as_float(self.obs_params['cost_sigma']) ** 2, device=like.device, dtype=like.dtype).expand(B, T, 1, 1)
            eps = torch.full((1, 1, 1, 1), as_float(self.obs_params['epsilon_dist']), device=like.device, dtype=like.dtype).expand(B, T, 1, 1)
            self._static_views[key] = (qc, w, eps)
        return self._static_views[key]

    def _is_static(self, qc, w, eps):
        for views in self._static_views.values():
            if qc is views[0] and w is views[1] and eps is views[2]:
                return True
        return False

    # ------------------------------------------------------------------ the GN step
    def forward(self, thb, startb, goalb, imb, sdfb, qc_inv_trajb, obscov_inv_trajb, eps_trajb):
        """One GN iteration for the whole batch (``imb`` is accepted and ignored, as in the reference)."""
        self.start_prior.set_mean(startb)
        self.goal_prior.set_mean(goalb)
        if self.q_full:
            self.gp_prior.set_inv_cov(qc_inv_trajb)
        else:
            self.gp_prior.Q_c_inv, self.gp_prior.Q_inv = qc_inv_trajb, None     # Q^-1 is rebuilt inside the kernel
        self.obs_factor.set_inv_cov(obscov_inv_trajb)
        self.obs_factor.set_eps(eps_trajb)
        static = self._is_static(qc_inv_trajb, obscov_inv_trajb, eps_trajb)
        self._state = dict(start=startb, goal=goalb, qc=qc_inv_trajb, w=obscov_inv_trajb, eps=eps_trajb, static=static, head=None)
        # expanded (broadcast) weight tensors are passed as they are; autograd sums their gradient
        return _GNStep.apply(self, static, None, thb, startb, goalb, sdfb, qc_inv_trajb, obscov_inv_trajb, eps_trajb)

    # ------------------------------------------------------------------ the GN step with the fused covariance head
    def split_head(self, out, mode='diag_identity', learn_eps=False):
        """The learned module's output ``out`` (B,1,out_dim) as the three raw slices the reference's
        ``get_covariances`` multiplies out (diff_gpmp2_planner.py:247-283), as zero-copy views:
        q (B,T-1,n) with n = 0 / 1 / dof / state_dim values per GP factor (None for 'fix_dynamics'),
        o (B,T), e (B,T) or None."""
        B = out.shape[0]
        G, S = self.num_gp_factors, self.num_obs_factors * self.nlinks
        n = _lib.head_block(mode, self.dof)
        flat = out[:, 0] if out.dim() == 3 else out
        need = G * n + S + (S if learn_eps else 0)
        if flat.dim() != 2 or flat.shape[1] < need or (learn_eps and flat.shape[1] != need):
            raise ValueError('head output must be (B,1,%d) for dynamics_mode %r, learn_eps=%s; got %s'
                             % (need, mode, learn_eps, tuple(out.shape)))
        q = flat[:, :G * n].reshape(B, G, n) if n else None
        o = flat[:, G * n:G * n + S]
        e = flat[:, G * n + S:G * n + 2 * S] if learn_eps else None
        return q, o, e

    def forward_head(self, thb, startb, goalb, imb, sdfb, out, mode=None, learn_eps=False):
        """``forward`` fed with the learned module's raw output instead of the covariances: replaces
        ``get_covariances`` + ``forward`` of the reference (diff_gpmp2_planner.py:183-206) by ONE launch --
        the kernels square / outer-multiply the raw values while they assemble each state (DGPMP2_FLAG_HEAD),
        so the (B,T-1,dof,dof), (B,T,1,1) covariance tensors never exist.  Differentiable w.r.t. ``out``.
        Without ``learn_eps`` the constructor's epsilon_dist is used, as in the reference's ``step`` (:205).
        Statefulness: the start / goal means are installed on the prior factors and ``error_batch`` /
        ``error_ext_batch`` / ``information_band`` evaluate with this call's covariances, as after ``forward``; the
        covariance attributes of ``gp_prior`` / ``obs_factor`` are NOT rewritten (those tensors are never formed) --
        call ``get_covariances`` + ``forward`` when stand-alone factor objects must carry them."""
        mode = mode or self.dynamics_mode or 'diag_identity'
        q, o, e = self.split_head(out, mode, learn_eps)
        self.start_prior.set_mean(startb)
        self.goal_prior.set_mean(goalb)
        self._state = dict(start=startb, goal=goalb, qc=q, w=o, eps=e, static=False, head=mode)
        return _GNStep.apply(self, False, mode, thb, startb, goalb, sdfb, q, o, e)

    def _check(self, status):
        self.last_status = status
        if self.strict and status is not None:
            bad = torch.nonzero(status)
            if bad.numel() > 0:
                b = int(bad[0])
                raise RuntimeError('cholesky: the Gauss-Newton system of problem %d is not positive-definite '
                                   '(first failing state %d)' % (b, int(status[b]) - 1))

    # ------------------------------------------------------------------ errors
    def _installed(self, dt):
        """(params, weight kwargs) of the covariances installed by the last forward() / forward_head()."""
        s = self._state
        head = s.get('head')
        p = self.cparams(q_full=(head == 'q_full') if head is not None else None)
        if s['static']:
            return p, {}
        c = lambda t: to_cuda(t.detach(), dt) if t is not None else None
        return p, dict(qc_inv=c(s['qc']), w_obs=c(s['w']), eps=c(s['eps']), head=head)

    def _errors(self, thb, sdfb):
        """(err, err_ext, err_sg, err_gp, err_obs), each (B,), differentiable w.r.t. thb (err excepted)."""
        if self._state is None:
            raise RuntimeError('PlanLayer: call forward() (or set the factor means / covariances) before error_batch')
        return list(_Errors.apply(self, thb, sdfb))

    def error_batch(self, thb, sdfb):
        """Normalised weighted error 0.5 sum e^T K e / M with the covariances of the last forward() -> (B,1,1)."""
        with torch.no_grad():
            return self._errors(thb, sdfb)[0].reshape(-1, 1, 1)

    def error_ext_batch(self, thb, sdfb):
        """Same with the constructor-time covariances (reference :310-345) -> (B,1,1)."""
        return self._errors(thb, sdfb)[1].reshape(-1, 1, 1)

    def start_goal_error(self, thb):
        dummy = torch.zeros(thb.shape[0], 1, 2, 2, device=thb.device, dtype=thb.dtype)
        return self._errors(thb, dummy)[2].reshape(-1, 1)

    def gp_error(self, thb):
        dummy = torch.zeros(thb.shape[0], 1, 2, 2, device=thb.device, dtype=thb.dtype)
        return self._errors(thb, dummy)[3].reshape(-1, 1, 1)

    def obs_error(self, thb, sdfb):
        return self._errors(thb, sdfb)[4].reshape(-1, 1, 1)

    def unweighted_errors(self, thb, sdfb):
        """(err_sg (B,1), err_gp (B,1,1), err_obs (B,1,1)) from ONE factor sweep."""
        o = self._errors(thb, sdfb)
        return o[2].reshape(-1, 1), o[3].reshape(-1, 1, 1), o[4].reshape(-1, 1, 1)

    # ------------------------------------------------------------------ the information system itself
    def information_band(self, thb, sdfb):
        """Block-tridiagonal normal equations of the last-installed problem in float64:
        D (B,T,d,d), U (B,T-1,d,d), r (B,T,d) -- what the reference holds as dense A^T K A + reg I, A^T K b."""
        s = self._state
        dt = work_dtype(thb, sdfb)
        p, kw = self._installed(dt)
        return ops.band(p, to_cuda(thb, dt), to_cuda(s['start'], dt), to_cuda(s['goal'], dt), to_cuda(sdfb, dt), **kw)
