"""Trajectory initialisation and convergence tests (host-side glue).

API mirror of reference ``diff_gpmp2/utils/planner_utils.py``:
``check_convergence`` (:3-16), ``check_convergence_batch`` (:18-36),
``straight_line_traj`` (:38-45), ``straight_line_trajb`` (:47-56),
``path_to_traj_avg_vel`` (:60-71), ``smoothness_metrics`` (:75-90),
``collision_metrics`` (:92-102).  Vectorised; no per-step Python loops.
"""
import torch


def check_convergence(dtheta, j, err_delta, tol_err, tol_delta, max_iters, method='gauss_newton', verbose=True):
    """True when ``||dtheta||_2 < tol_delta`` or ``j >= max_iters``.  ``tol_err`` is
    accepted and ignored exactly as in the reference (its test is commented out, :7-9)."""
    nrm = float(torch.norm(dtheta))
    if nrm < tol_delta:
        if verbose:
            print('Update got too small at iter %d: %f' % (j, nrm))
        return True
    if j >= max_iters:
        if verbose:
            print('Max iters done')
        return True
    return False


def check_convergence_batch(dthetab, j, err_delta, tol_err, tol_delta, max_iters, method='gauss_newton',
                            device=torch.device('cpu')):
    """Per-problem convergence vector (B,1,1).  As in the reference (:24-27) the
    error-delta test overwrites the update-norm test, so only ``||err_delta|| < tol_err``
    (or ``j >= max_iters``) counts."""
    B = dthetab.shape[0]
    if j >= max_iters:
        print('Max iters done')
        return torch.ones(B, 1, 1, dtype=torch.uint8, device=dthetab.device)
    ed = torch.norm(err_delta.reshape(B, -1), dim=1, p=2)
    return (ed < tol_err).to(torch.int64).reshape(B, 1, 1)


def straight_line_traj(start_conf, goal_conf, traj_time, num_steps, dof, device=torch.device('cpu')):
    """(num_steps+1, 2*dof): positions interpolated linearly, constant average velocity."""
    n = int(num_steps)
    s = torch.as_tensor(start_conf, device=device).reshape(-1)[:dof]
    g = torch.as_tensor(goal_conf, device=device).reshape(-1)[:dof]
    i = torch.arange(n + 1, device=device, dtype=s.dtype).unsqueeze(1)
    pos = s.unsqueeze(0) * (num_steps - i) * 1. / num_steps * 1. + g.unsqueeze(0) * i * 1. / num_steps * 1.
    vel = ((g - s) / traj_time * 1.0).unsqueeze(0).expand(n + 1, dof)
    return torch.cat((pos, vel), dim=1).to(torch.get_default_dtype())


def straight_line_trajb(start_confb, goal_confb, traj_time, num_steps, dof, device=torch.device('cpu')):
    """Batched version: ``start_confb``/``goal_confb`` are (B,1,>=dof) -> (B, num_steps+1, 2*dof)."""
    n = int(num_steps)
    s = start_confb[:, 0, 0:dof].to(device)
    g = goal_confb[:, 0, 0:dof].to(device)
    i = torch.arange(n + 1, device=device, dtype=s.dtype).reshape(1, n + 1, 1)
    pos = s.unsqueeze(1) * (num_steps - i) * 1.0 / num_steps * 1.0 + g.unsqueeze(1) * i * 1.0 / num_steps * 1.0
    vel = ((goal_confb.to(device) - start_confb.to(device)) / traj_time * 1.0)[:, :, 0:dof].expand(-1, n + 1, -1)
    return torch.cat((pos, vel), dim=2).to(torch.get_default_dtype())


def path_to_traj_avg_vel(path, traj_time, dof, device=torch.device('cpu')):
    p = torch.as_tensor(path, device=device)
    vel = ((p[-1] - p[0]) / traj_time * 1.0).unsqueeze(0).expand(p.shape[0], dof)
    return torch.cat((p[:, :dof], vel), dim=1).to(torch.get_default_dtype())


def smoothness_metrics(traj, total_time_sec, total_time_step):
    d1 = traj[1:, :] - traj[:-1, :]
    d2 = d1[1:, :] - d1[:-1, :]
    vel = traj[:, 2:]
    acc = d1[:, 2:] / total_time_step * 1.0
    jerk = d2[:, 2:] / (total_time_step ** 2.0)
    return (torch.norm(vel, p=2, dim=1).mean(), torch.norm(acc, p=2, dim=1).mean(),
            torch.norm(jerk, p=2, dim=1).mean())


def collision_metrics(traj, obs_error, total_time_sec, total_time_step):
    inner = obs_error[1:-1, :]
    num_pen = torch.numel(torch.nonzero(inner)) / 2
    dt = total_time_sec * 1.0 / total_time_step * 1.0
    return num_pen > 0, inner.mean(), inner.max(), (num_pen * dt) / total_time_sec * 1.0
