"""The planner module (API mirror of reference ``diff_gpmp2/gpmp2/diff_gpmp2_planner.py:15-299``).

``DiffGPMP2Planner(gp_params, obs_params, planner_params, optim_params, env_params, robot_model,
learn_params=None, batch_size=1, use_cuda=False)``
  ``.step(th_currb, startb, goalb, imb, sdfb, ...)`` -> 7-tuple (:176-211): one fused CUDA launch.
  ``.forward(th_initb, startb, goalb, imb, sdfb)``   -> 8-tuple (:92-174): the reference optimises
      the problems one after the other with B=1 sub-calls; here the whole batch is optimised to
      per-problem convergence in ONE persistent CUDA launch (dgpmp2_gn_solve_*), with identical
      per-problem semantics (independent problems, same stopping rule).
The learned-covariance networks (reference learning/*, cuDNN model code) are outside this
package; ``get_covariances`` (the mapping from network outputs to covariances, :247-290) is kept, and
the same mapping is available FUSED into the GN kernels: ``step_head(..., out)`` / a module installed
with ``set_learn_module`` feed the network's raw output straight to the launch (DGPMP2_FLAG_HEAD).
"""
import time

import torch
import torch.nn as nn

from .. import ops
from .._dev import as_float, back, to_cuda, work_dtype
from ..utils import mat_utils
from .plan_layer import PlanLayer


class DiffGPMP2Planner(nn.Module):
    def __init__(self, gp_params, obs_params, planner_params, optim_params, env_params, robot_model,
                 learn_params=None, batch_size=1, use_cuda=False):
        super(DiffGPMP2Planner, self).__init__()
        self.use_cuda = torch.cuda.is_available() if use_cuda else False
        self.device = torch.device('cuda') if self.use_cuda else torch.device('cpu')
        self.dof = int(planner_params['dof'])
        self.state_dim = int(planner_params['state_dim'])
        self.total_time_sec = planner_params['total_time_sec']
        self.total_time_step = int(planner_params['total_time_step'])
        self.num_traj_states = self.total_time_step + 1
        self.num_gp_factors = self.num_traj_states - 1
        self.num_obs_factors = self.num_traj_states
        self.optim_params, self.gp_params, self.obs_params = optim_params, gp_params, obs_params
        self.robot_model, self.env_params, self.learn_params = robot_model, env_params, learn_params
        self.model_type = 'feed_forward'
        self.non_holonomic = bool(planner_params.get('non_holonomic', False))
        self.use_vel_limits = bool(planner_params.get('use_vel_limits', False))
        self.batch_size = batch_size
        nl = robot_model.nlinks
        # module slots with the reference's names (:51-52, :87-88): a reference checkpoint's keys
        # (learn_module_conv.*, learn_module_fcn.*) line up once modules are assigned to them
        self.learn_module_conv = None
        self.learn_module_fcn = None
        self._module_protocol = 'none'
        qc_inv = torch.as_tensor(gp_params['Q_c_inv'])
        if learn_params is None:
            inv_cov = mat_utils.isotropic_matrix(1.0 / torch.pow(torch.as_tensor(obs_params['cost_sigma']), 2.0), nl)
            # constant per-trajectory covariances, same shapes as the reference (:41-48)
            self.qc_inv_traj = torch.zeros(self.num_gp_factors, self.dof, self.dof) + qc_inv
            self.obscov_inv_traj = torch.zeros(self.num_traj_states, nl, 1) + inv_cov
            self.eps_traj = torch.zeros(self.num_traj_states, nl, 1) + torch.as_tensor(obs_params['epsilon_dist'])
        else:
            # Reference :53-88.  The networks themselves (LearnModuleConv / LearnModuleFCN, reference learning/*) are
            # outside this package: everything the constructor derives from learn_params is derived here, the two
            # module slots stay empty until the caller assigns modules with the reference's call protocol
            # (conv(im_in) -> (features, _); fcn(th_in, features[, hidden]) -> out (B,1,out_dim)[, hidden]), and
            # step / forward raise until then.  learn_params is updated in place exactly as the reference does.
            lp = learn_params
            self.model_type = lp['model']['type'] if 'type' in lp['model'] else False
            self.learn_eps = lp['dgpmp2']['learn_eps'] if 'learn_eps' in lp['dgpmp2'] else False
            self.sdf_predict = lp['dgpmp2']['sdf_predict']
            self.use_dtheta = lp['dgpmp2']['dtheta_predict'] if 'dtheta_predict' in lp['dgpmp2'] else False
            self.res = (env_params['x_lims'][1] - env_params['x_lims'][0]) / (lp['data']['im_size'] * 1.0)
            lp['num_traj_states'] = self.num_traj_states
            lp['state_dim'] = self.state_dim
            if self.use_dtheta:
                lp['num_traj_states'] = 2 * self.num_traj_states
            self.dynamics_mode = lp['dgpmp2']['dynamics_mode']
            n_obs = self.num_obs_factors * nl
            if self.dynamics_mode == 'fix_dynamics':
                lp['out_dim'] = n_obs
                self.qc_inv_traj = torch.zeros(self.num_gp_factors, self.dof, self.dof) + qc_inv
            elif self.dynamics_mode == 'diag_identity':
                lp['out_dim'] = self.num_gp_factors + n_obs
            elif self.dynamics_mode in ('diag', 'qc_full'):
                lp['out_dim'] = self.num_gp_factors * self.dof + n_obs
            elif self.dynamics_mode == 'q_full':
                lp['out_dim'] = self.num_gp_factors * self.state_dim + n_obs
            if self.learn_eps:
                lp['out_dim'] = lp['out_dim'] + n_obs
            else:
                self.eps = obs_params['epsilon_dist']
                self.eps_traj = torch.zeros(self.num_traj_states, nl, 1) + torch.as_tensor(self.eps)
            self.fixed_conv = lp['dgpmp2']['fixed_conv'] if 'fixed_conv' in lp['dgpmp2'] else False
            self.out_dim = lp.get('out_dim')
            self._module_protocol = 'reference'
        self.plan_layer = PlanLayer(gp_params, obs_params, planner_params, optim_params, env_params, robot_model,
                                    learn_params, batch_size, self.use_cuda)

    # ------------------------------------------------------------------ optimise to convergence
    def forward(self, th_initb, startb, goalb, imb, sdfb, hiddenb=None):
        """-> (th_finalb (B,T,d), hidden, err_initb[B], err_finalb[B], err_per_iterb[B][j],
        err_ext_per_iterb[B][j], jb[B], timeb[B])."""
        start_t = time.time()
        B = th_initb.shape[0]
        plan_time = float(self.optim_params.get('plan_time', 'inf'))
        max_iters = int(self.optim_params['max_iters'])
        tol_delta = as_float(self.optim_params['tol_delta'])
        pl = self.plan_layer
        qc, w, eps = pl.static_weights(B, th_initb)
        # keep the reference's post-conditions: the factor state of the last plan_layer call is installed
        pl.start_prior.set_mean(startb)
        pl.goal_prior.set_mean(goalb)
        pl.gp_prior.Q_c_inv, pl.gp_prior.Q_inv = qc, None
        pl.obs_factor.set_inv_cov(w)
        pl.obs_factor.set_eps(eps)
        pl._state = dict(start=startb, goal=goalb, qc=qc, w=w, eps=eps, static=True)
        needs_grad = torch.is_grad_enabled() and any(
            isinstance(t, torch.Tensor) and t.requires_grad for t in (th_initb, startb, goalb, sdfb))
        if self._module_protocol != 'none':
            # covariances re-predicted at every iterate (:128-147): batched step() loop through the fused head
            return self._forward_timed(th_initb, startb, goalb, imb, sdfb, plan_time, start_t, torch.is_grad_enabled())
        if plan_time != float('inf') or needs_grad:
            # differentiable (unrolled, like the reference) or wall-clock-budgeted: batched step() loop
            return self._forward_timed(th_initb, startb, goalb, imb, sdfb, plan_time, start_t, needs_grad)
        dt = work_dtype(th_initb, sdfb)
        out = ops.gn_solve(pl.cparams(), to_cuda(th_initb, dt), to_cuda(startb, dt), to_cuda(goalb, dt),
                           to_cuda(sdfb, dt), max_iters, tol_delta)
        th_final, iters, epi, eepi, ef, eef, status = out
        pl._check(status)
        iters_h = iters.cpu().tolist()                   # one host sync for the whole batch
        epi_h, eepi_h, ef_h = epi.double().cpu(), eepi.double().cpu(), ef.double().cpu().tolist()
        err_per_iterb = [epi_h[i, :iters_h[i]].tolist() for i in range(B)]
        err_ext_per_iterb = [eepi_h[i, :iters_h[i]].tolist() for i in range(B)]
        err_initb = [e[0] for e in err_per_iterb]
        elapsed = time.time() - start_t
        return (back(th_final, th_initb).to(th_initb.dtype), None, err_initb, ef_h, err_per_iterb, err_ext_per_iterb,
                iters_h, [elapsed] * B)

    def _forward_timed(self, th_initb, startb, goalb, imb, sdfb, plan_time, start_t, needs_grad=False):
        """Batched step() loop with per-problem convergence masks: used when the result must be
        differentiable (the unrolled iterations stay on the autograd tape, as in the reference) or when
        optim_params['plan_time'] is finite (the reference's budget check, :154-156)."""
        B = th_initb.shape[0]
        max_iters = int(self.optim_params['max_iters'])
        tol_delta = as_float(self.optim_params['tol_delta'])
        th = th_initb if needs_grad else th_initb.detach().clone()
        done = torch.zeros(B, dtype=torch.bool, device=th.device)
        iters = [0] * B
        epi = [[] for _ in range(B)]
        eepi = [[] for _ in range(B)]
        for j in range(max_iters):
            dth, _, err, err_ext, _, _, _ = self.step(th, startb, goalb, imb, sdfb)
            nrm = torch.norm(dth.detach().reshape(B, -1), dim=1)
            e_h, ee_h, nrm_h, done_h = err.reshape(-1).tolist(), err_ext.reshape(-1).tolist(), nrm.tolist(), done.tolist()
            for i in range(B):
                if not done_h[i]:
                    epi[i].append(e_h[i])
                    eepi[i].append(ee_h[i])
                    iters[i] = j + 1
            th = torch.where(done.reshape(B, 1, 1), th, th + dth)
            done = done | (nrm < tol_delta)
            if bool(done.all()):
                break
            if plan_time != float('inf') and time.time() - start_t > plan_time:
                print('Plan time over')
                break
        ef = self.plan_layer.error_batch(th.detach(), sdfb.detach()).reshape(-1).tolist()
        return (th, None, [e[0] for e in epi], ef, epi, eepi, iters, [time.time() - start_t] * B)

    # ------------------------------------------------------------------ one iteration
    def set_learn_module(self, module, dynamics_mode='diag_identity', learn_eps=False):
        """Install a user-supplied learned module (the reference builds its own LearnModuleConv / LearnModuleFCN
        from learn_params, :78-88; those networks are outside this package).  ``module(th_currb, imb, sdfb)`` must
        return ``out`` (B,1,out_dim) laid out as ``get_covariances`` expects for ``dynamics_mode`` (:247-283).
        ``step`` then runs the module and ONE fused launch that forms the covariances inside the GN kernel."""
        from .. import _lib
        _lib.head_block(dynamics_mode, self.dof)          # validates the mode
        self.learn_module_fcn = module
        self.dynamics_mode = dynamics_mode
        self.learn_eps = bool(learn_eps)
        self._module_protocol = 'callable'

    def _predict(self, th_currb, imb, sdfb, conv_out=None, dtheta_currb=None, hiddenb=None):
        """Run the installed learned module(s) -> (out (B,1,out_dim), hidden).  'reference' protocol (planner built with
        learn_params; modules assigned to the learn_module_conv / learn_module_fcn slots): the reference's own call
        sequence (:186-193).  'callable' protocol (set_learn_module): out = module(th, im, sdf)."""
        if self._module_protocol == 'callable':
            return self.learn_module_fcn(th_currb, imb, sdfb), None
        if self.learn_module_fcn is None or (not self.fixed_conv and self.learn_module_conv is None):
            raise RuntimeError(
                'DiffGPMP2Planner was built with learn_params but no learned module is installed: assign modules to '
                'planner.learn_module_conv / planner.learn_module_fcn (the reference\'s LearnModuleConv / LearnModuleFCN '
                'protocol, out_dim = %s for dynamics_mode %r) or call planner.set_learn_module(module, dynamics_mode). '
                'The networks themselves are outside this package (reference diff_gpmp2/learning).'
                % (self.learn_params.get('out_dim'), self.dynamics_mode))
        if not self.fixed_conv:
            im_in = torch.cat((imb, sdfb), dim=1) if self.sdf_predict else imb
            conv_out, _ = self.learn_module_conv(im_in)
        th_in = torch.cat((th_currb, dtheta_currb), dim=-1) if self.use_dtheta else th_currb
        if self.model_type == 'feed_forward':
            return self.learn_module_fcn(th_in, conv_out), None
        return self.learn_module_fcn(th_in, conv_out, hiddenb)

    def step_head(self, th_currb, startb, goalb, imb, sdfb, out, mode=None, learn_eps=None):
        """One batched GN iteration from the learned module's raw output ``out`` (B,1,out_dim):
        -> (dthetab, err_oldb, err_ext_oldb).  Equivalent to ``get_covariances(out, mode, learn_eps)`` followed by
        ``plan_layer(...)`` (:196-206) without materialising the covariance tensors; differentiable w.r.t. ``out``."""
        mode = mode or getattr(self, 'dynamics_mode', None) or 'diag_identity'
        learn_eps = getattr(self, 'learn_eps', False) if learn_eps is None else learn_eps
        return self.plan_layer.forward_head(th_currb, startb, goalb, imb, sdfb, out, mode, learn_eps)

    def step(self, th_currb, startb, goalb, imb, sdfb, conv_out=None, dtheta_currb=None, hiddenb=None):
        """One batched GN iteration -> (dthetab, hidden, err_oldb, err_ext_oldb, qc_inv, obscov_inv, eps)."""
        B = th_currb.shape[0]
        if self._module_protocol != 'none':
            out, hidden = self._predict(th_currb, imb, sdfb, conv_out, dtheta_currb, hiddenb)
            dthetab, err_oldb, err_ext_oldb = self.step_head(th_currb, startb, goalb, imb, sdfb, out)
            with torch.no_grad():       # the covariances of the return tuple are reporting only (train_planner.py:310-311)
                cov = self.get_covariances(out, self.dynamics_mode, self.learn_eps)
                cov = list(cov) if isinstance(cov, tuple) else [cov]
                if self.dynamics_mode == 'fix_dynamics':
                    cov.insert(0, self.qc_inv_traj.to(out.device, out.dtype).unsqueeze(0).expand(B, -1, -1, -1))
                if not self.learn_eps:
                    cov.append(self.eps_traj.to(out.device, out.dtype).unsqueeze(0).expand(B, -1, -1, -1))
            return dthetab, (hidden if hiddenb is not None else None), err_oldb, err_ext_oldb, cov[0], cov[1], cov[2]
        qc, w, eps = self.plan_layer.static_weights(B, th_currb)
        dthetab, err_oldb, err_ext_oldb = self.plan_layer(th_currb, startb, goalb, imb, sdfb, qc, w, eps)
        return dthetab, None, err_oldb, err_ext_oldb, qc, w, eps

    # ------------------------------------------------------------------ errors
    def error_batch(self, thb, sdfb):
        return self.plan_layer.error_batch(thb, sdfb)

    def error_ext_batch(self, thb, sdfb):
        return self.plan_layer.error_ext_batch(thb, sdfb)

    def unweighted_errors_batch(self, thb, sdfb):
        return self.plan_layer.unweighted_errors(thb, sdfb)

    # ------------------------------------------------------------------ network output -> covariances
    def get_covariances(self, out, mode='diag_identity', learn_eps=False):
        """Map a learned module's output vector ``out`` (B,1,out_dim) to (qc_inv_traj, obscov_inv_traj[, eps_traj])
        exactly as the reference does (:247-290): every quantity is an outer product q q^T of a slice of
        ``out`` (so it is PSD), multiplied by I in 'diag_identity' mode."""
        nl = self.robot_model.nlinks
        G, S, B = self.num_gp_factors, self.num_obs_factors, out.shape[0]
        n_obs = S * nl
        if mode == 'fix_dynamics':
            n_gp, blk = 0, 0
        elif mode == 'diag_identity':
            n_gp, blk = G, 1
        elif mode == 'qc_full':
            n_gp, blk = G * self.dof, self.dof
        elif mode == 'q_full':
            n_gp, blk = G * self.state_dim, self.state_dim
        else:
            raise NotImplementedError(mode)
        qc_inv_traj = None
        if blk:
            q = out[:, 0, 0:n_gp].reshape(B, G, blk, 1)
            qc_inv_traj = q * q.transpose(2, 3)
            if mode == 'diag_identity':
                qc_inv_traj = qc_inv_traj * torch.eye(self.dof, device=out.device, dtype=out.dtype)
        o = out[:, 0, n_gp:n_gp + n_obs].reshape(B, S, nl, 1)
        obscov_inv_traj = o * o.transpose(2, 3)
        res = [] if mode == 'fix_dynamics' else [qc_inv_traj]
        res.append(obscov_inv_traj)
        if learn_eps:
            e = out[:, 0, n_gp + n_obs:].reshape(B, S, nl, 1)
            res.append(e * e.transpose(2, 3))
        return res[0] if len(res) == 1 else tuple(res)

    def get_obs_covariance(self, out):
        nl = self.robot_model.nlinks
        return torch.eye(nl, device=out.device, dtype=out.dtype).expand(self.num_obs_factors, nl, nl) * \
            (out * out).reshape(self.num_obs_factors, 1, 1)
