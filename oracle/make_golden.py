"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the LIVE, unmodified
reference (/root/reference, via oracle/ref_harness.py) on seeded inputs.

    python -m oracle.make_golden            # from the repo root, in the build container

Protocol: every array input (trajectory, start, goal, SDF, per-state weights) is rounded to
float32 and the reference is fed those float32 values cast up to float64, so that the float32-
and float64-I/O kernels and the oracle all see bit-identical inputs.  Outputs are the
reference's float64 results.  The reference cannot run on the GPU box, hence the fixtures.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, 'tests', 'golden')

YAML = dict(Q_c_inv=[[1.0, 0.0], [0.0, 1.0]], K_s=0.01, K_g=0.01, cost_sigma=0.01, epsilon_dist=0.4,
            reg=0.1, total_time_sec=10.0, sphere_radius=0.4, max_iters=100, tol_delta=1e-4, tol_err=1e-3)


def r32(x):
    """round to float32, return float64 tensor"""
    return torch.as_tensor(np.asarray(x, dtype=np.float64)).to(torch.float32).to(torch.float64)


def ref_band(planner, th, sdf):
    """Dense A,b,K of the live reference -> normal equations -> band."""
    from oracle import gn_oracle
    pl = planner.plan_layer
    A, b, K = pl.construct_linear_system_batch(th, sdf)
    LAM, R = gn_oracle.normal_equations(A, b, K, float(pl.optim_params['reg']))
    D, U, r, off = gn_oracle.band_from_dense(LAM, R, pl.num_traj_states, pl.state_dim)
    return D, U, r, off


def run_step_case(name, th, start, goal, sdf, params, x_lims, y_lims, qc=None, w=None, eps=None, q_full=False,
                  k_updates=0, extra=None):
    from oracle import ref_harness
    B, T, d = th.shape
    th, start, goal, sdf = r32(th), r32(start), r32(goal), r32(sdf)
    learn = {'dgpmp2': {'dynamics_mode': 'q_full'}} if q_full else None
    if q_full:
        # DiffGPMP2Planner would build the learning nets; PlanLayer alone only reads dynamics_mode
        ref_harness.import_reference()
        planner = ref_harness.make_reference_planner(B, T, params, x_lims, y_lims)
        from diff_gpmp2.gpmp2.plan_layer import PlanLayer
        base = planner.plan_layer
        pl = PlanLayer(base.gp_params, base.obs_params, base.planner_params, base.optim_params, base.env_params,
                       base.robot_model, learn_params=learn, batch_size=B)
        planner.plan_layer = pl
    else:
        planner = ref_harness.make_reference_planner(B, T, params, x_lims, y_lims)
    im = torch.zeros_like(sdf)
    static = qc is None
    if static:
        qc_t = planner.qc_inv_traj.unsqueeze(0).repeat(B, 1, 1, 1)
        w_t = planner.obscov_inv_traj.unsqueeze(0).repeat(B, 1, 1, 1)
        eps_t = planner.eps_traj.unsqueeze(0).repeat(B, 1, 1, 1)
    else:
        qc_t, w_t, eps_t = r32(qc), r32(w), r32(eps)
    with torch.no_grad():
        for _ in range(k_updates):
            dth, _, _ = planner.plan_layer(th, start, goal, im, sdf, qc_t, w_t, eps_t)
            th = r32((th + dth).numpy())
        dth, err, err_ext = planner.plan_layer(th, start, goal, im, sdf, qc_t, w_t, eps_t)
        D, U, r, off = ref_band(planner, th, sdf)
        e_gp, _, _ = planner.plan_layer.gp_prior.get_error(th)
        c_obs, H_obs = planner.plan_layer.obs_factor.get_error(th, sdf)
        planner.plan_layer.obs_factor.set_eps(eps_t)
        e_sg, e_gpu, e_obs = planner.unweighted_errors_batch(th, sdf)
    out = dict(th=th.numpy().astype(np.float32), start=start.numpy().astype(np.float32),
               goal=goal.numpy().astype(np.float32), sdf=sdf.numpy().astype(np.float32),
               x_lims=np.array(x_lims), y_lims=np.array(y_lims), T=T, static=static, q_full=q_full,
               dth=dth.numpy(), err=err.numpy(), err_ext=err_ext.numpy(), band_D=D.numpy(), band_U=U.numpy(),
               band_r=r.numpy(), off_band_max=off, gp_err=e_gp.numpy(), obs_cost=c_obs.numpy(), obs_H=H_obs.numpy(),
               err_sg=e_sg.numpy(), err_gp=e_gpu.numpy(), err_obs=e_obs.numpy(),
               params_keys=np.array(sorted(params.keys())), params_vals=np.array([str(params[k]) for k in sorted(params.keys())]))
    if not static:
        out.update(qc=qc_t.numpy().astype(np.float32), w=w_t.numpy().astype(np.float32), eps=eps_t.numpy().astype(np.float32))
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print('%-28s B=%d T=%d |dth|max=%.3e err=%s off_band=%.1e active=%d/%d' % (
        name, B, T, float(dth.abs().max()), err.reshape(-1)[:2].numpy(), off, int((c_obs > 0).sum()), c_obs.numel()))


def synth(B, T, im_size, seed):
    from dgpmp2_b200.datasets.synthetic import make_problems
    pr = make_problems(B, T, dof=2, im_size=im_size, seed=seed, dtype=torch.float64)
    return pr['th_init'], pr['start'], pr['goal'], pr['sdf']


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    np.random.seed(0)
    rng = np.random.default_rng(7)
    from oracle import ref_harness
    ref_harness.import_reference()
    torch.set_default_dtype(torch.float64)
    lims = (-5.0, 5.0)

    # 1. config-2 shape, static covariances, straight-line iterate and a later iterate
    th, s, g, sdf = synth(4, 64, 128, seed=1)
    run_step_case('step_static_B4_T64_k0', th, s, g, sdf, YAML, lims, lims)
    run_step_case('step_static_B4_T64_k5', th, s, g, sdf, YAML, lims, lims, k_updates=5)

    # 2. per-(b,t) learned weights (Qc^-1 rank-1 + diag, w_obs, eps)
    B, T = 3, 16
    th, s, g, sdf = synth(B, T, 64, seed=2)
    th = th + 0.05 * torch.as_tensor(rng.standard_normal(th.shape))
    q = torch.as_tensor(rng.standard_normal((B, T - 1, 2, 1)))
    qc = q @ q.transpose(2, 3) + torch.eye(2) * torch.as_tensor(rng.uniform(0.2, 1.5, (B, T - 1, 1, 1)))
    w = torch.as_tensor(rng.uniform(50.0, 2.0e4, (B, T, 1, 1)))
    eps = torch.as_tensor(rng.uniform(0.1, 0.8, (B, T, 1, 1)))
    run_step_case('step_learned_B3_T16', th, s, g, sdf, YAML, lims, lims, qc=qc, w=w, eps=eps)

    # 3. q_full mode: full d x d GP inverse covariances given directly (rank-1 + diag)
    B, T = 2, 16
    th, s, g, sdf = synth(B, T, 64, seed=3)
    q = torch.as_tensor(rng.standard_normal((B, T - 1, 4, 1)))
    qf = q @ q.transpose(2, 3) + torch.eye(4) * 0.5
    w = torch.as_tensor(rng.uniform(50.0, 2.0e4, (B, T, 1, 1)))
    eps = torch.as_tensor(rng.uniform(0.1, 0.8, (B, T, 1, 1)))
    run_step_case('step_qfull_B2_T16', th, s, g, sdf, YAML, lims, lims, qc=qf, w=w, eps=eps, q_full=True)

    # 4. out-of-image states, non-square SDF (res from width only), asymmetric limits
    B, T = 2, 8
    Hh, Ww = 24, 40
    sdf = torch.as_tensor(rng.uniform(-0.5, 2.0, (B, 1, Hh, Ww)))
    xl, yl = (-4.0, 6.0), (-3.0, 3.0)
    th = torch.as_tensor(rng.uniform(-7.0, 7.0, (B, T, 4)))
    th[0, 0, 0:2] = torch.tensor([-4.0, -3.0])      # exactly on the lower limits
    th[0, 1, 0:2] = torch.tensor([6.0, 3.0])        # exactly on the upper limits
    th[0, 2, 0:2] = torch.tensor([0.0, 0.0])
    th[1, 0, 0:2] = torch.tensor([5.9, -2.9])
    s = th[:, 0:1, :].clone() + 0.1
    g = th[:, -1:, :].clone() - 0.1
    run_step_case('step_oob_B2_T8', th, s, g, sdf, YAML, xl, yl)

    # 5. non-power-of-two T (the yaml default T = 101) and T = 128
    th, s, g, sdf = synth(1, 101, 64, seed=4)
    run_step_case('step_static_B1_T101_k3', th, s, g, sdf, YAML, lims, lims, k_updates=3)
    th, s, g, sdf = synth(2, 128, 64, seed=5)
    run_step_case('step_static_B2_T128_k2', th, s, g, sdf, YAML, lims, lims, k_updates=2)
    th, s, g, sdf = synth(2, 2, 32, seed=6)
    run_step_case('step_static_B2_T2', th, s, g, sdf, YAML, lims, lims)
    th, s, g, sdf = synth(2, 3, 32, seed=8)
    run_step_case('step_static_B2_T3', th, s, g, sdf, YAML, lims, lims)

    # 6. config 1: the reference's own example flow (examples/diff_gpmp2_2d_example.py:40-67)
    from PIL import Image
    from diff_gpmp2.utils.sdf_utils import sdf_2d as ref_sdf_2d
    from diff_gpmp2.utils.planner_utils import straight_line_traj as ref_sl
    png = os.path.join(ref_harness.reference_root(), 'diff_gpmp2', 'env', 'simple_2d', '5.png')
    img = np.asarray(Image.open(png).convert('L'), dtype=np.float64) / 255.0
    cell = (lims[1] - lims[0]) / img.shape[0]
    env_sdf = r32(ref_sdf_2d(img, res=cell))                       # default padlen=1 -> 202 x 202
    T = 64
    start_conf = torch.tensor([[lims[0] + 1.0, lims[0] + 1.0]])
    goal_conf = torch.tensor([[lims[1] - 1.0, lims[1] - 1.0]])
    start = torch.cat((start_conf, torch.zeros(1, 2)), dim=1)
    goal = torch.cat((goal_conf, torch.zeros(1, 2)), dim=1)
    th_init = r32(ref_sl(start_conf, goal_conf, YAML['total_time_sec'], T - 1, 2).numpy())
    planner = ref_harness.make_reference_planner(1, T, YAML, lims, lims)
    import io
    import contextlib
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        out = planner.forward(th_init.unsqueeze(0), start.unsqueeze(0), goal.unsqueeze(0),
                              torch.zeros(1, 1, *env_sdf.shape), env_sdf.unsqueeze(0).unsqueeze(0))
    th_final, _, err_init, err_final, err_pi, err_ext_pi, k, _ = out
    run_step_case('config1_step_T64', th_init.unsqueeze(0), start.unsqueeze(0), goal.unsqueeze(0),
                  env_sdf.unsqueeze(0).unsqueeze(0), YAML, lims, lims,
                  extra=dict(fwd_th_final=th_final.numpy(), fwd_err_init=np.array(err_init),
                             fwd_err_final=np.array(err_final), fwd_err_per_iter=np.array(err_pi[0]),
                             fwd_err_ext_per_iter=np.array(err_ext_pi[0]), fwd_iters=np.array(k)))
    print('config1 forward: iters=%s err %.4f -> %.4f' % (k, err_init[0], err_final[0]))

    # 6b. forward on a problem that converges before max_iters (tol_delta raised) + a 3-problem batch run per sample
    P2 = dict(YAML)
    P2.update(tol_delta=1.0, max_iters=30)
    th, s, g, sdf = synth(3, 32, 64, seed=9)
    th, s, g, sdf = r32(th), r32(s), r32(g), r32(sdf)
    planner = ref_harness.make_reference_planner(1, 32, P2, lims, lims)
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        out = planner.forward(th, s, g, torch.zeros_like(sdf), sdf)
    th_final, _, err_init, err_final, err_pi, err_ext_pi, k, _ = out
    L = max(len(e) for e in err_pi)
    epi = np.full((3, L), np.nan)
    eepi = np.full((3, L), np.nan)
    for i in range(3):
        epi[i, :len(err_pi[i])] = err_pi[i]
        eepi[i, :len(err_ext_pi[i])] = err_ext_pi[i]
    np.savez_compressed(os.path.join(OUT, 'forward_B3_T32.npz'), th=th.numpy().astype(np.float32),
                        start=s.numpy().astype(np.float32), goal=g.numpy().astype(np.float32),
                        sdf=sdf.numpy().astype(np.float32), x_lims=np.array(lims), y_lims=np.array(lims), T=32,
                        max_iters=30, tol_delta=1.0, fwd_th_final=th_final.numpy(), fwd_err_init=np.array(err_init),
                        fwd_err_final=np.array(err_final), fwd_err_per_iter=epi, fwd_err_ext_per_iter=eepi,
                        fwd_iters=np.array(k))
    print('forward_B3_T32: iters=%s' % (k,))

    # 7. bilinear_interpolate alone (incl. points outside the image)
    from diff_gpmp2.utils.sdf_utils import bilinear_interpolate as ref_bil
    B, N, Hh, Ww = 3, 40, 20, 28
    sdf = r32(rng.uniform(-1.0, 3.0, (B, Hh, Ww)))
    pts = r32(rng.uniform(-6.5, 6.5, (B, N, 2)))
    res = (lims[1] - lims[0]) / Ww
    dist, J = ref_bil(sdf, pts, res, list(lims), list(lims))
    np.savez_compressed(os.path.join(OUT, 'bilinear_B3_N40.npz'), sdf=sdf.numpy().astype(np.float32),
                        pts=pts.numpy().astype(np.float32), res=res, x_lims=np.array(lims), y_lims=np.array(lims),
                        dist=dist.numpy(), J=J.numpy())
    print('bilinear: zero-dist points (outside) = %d / %d' % (int((dist == 0).sum()), dist.numel()))

    # 9. gradients of one GN step through the reference's own autograd graph (dense scatter + cholesky + inverse)
    B, T = 2, 16
    th, s_, g_, sdf = synth(B, T, 64, seed=11)
    th = th + 0.05 * torch.as_tensor(rng.standard_normal(th.shape))
    q = torch.as_tensor(rng.standard_normal((B, T - 1, 2, 1)))
    qc = q @ q.transpose(2, 3) + torch.eye(2) * torch.as_tensor(rng.uniform(0.2, 1.5, (B, T - 1, 1, 1)))
    w = torch.as_tensor(rng.uniform(50.0, 2.0e4, (B, T, 1, 1)))
    eps = torch.as_tensor(rng.uniform(0.3, 1.2, (B, T, 1, 1)))
    leaves = [r32(x).requires_grad_(True) for x in (th, s_, g_, sdf, qc, w, eps)]
    planner = ref_harness.make_reference_planner(B, T, YAML, lims, lims)
    dth, err, err_ext = planner.plan_layer(leaves[0], leaves[1], leaves[2], torch.zeros_like(leaves[3]), leaves[3],
                                           leaves[4], leaves[5], leaves[6])
    G = torch.as_tensor(rng.standard_normal(dth.shape))
    g2 = torch.as_tensor(rng.standard_normal(err_ext.shape))
    loss = (dth * G).sum() + (err_ext * g2).sum()
    grads = torch.autograd.grad(loss, leaves, allow_unused=True)
    names = ['th', 'start', 'goal', 'sdf', 'qc', 'w', 'eps']
    out = {n: l.detach().numpy().astype(np.float32) for n, l in zip(names, leaves)}
    out.update({'g_' + n: (gr.numpy() if gr is not None else np.zeros(tuple(l.shape))) for n, gr, l in zip(names, grads, leaves)})
    out.update(G=G.numpy(), g_err_ext=g2.numpy(), dth=dth.detach().numpy(), err_ext=err_ext.detach().numpy(), T=T,
               x_lims=np.array(lims), y_lims=np.array(lims), err_requires_grad=bool(err.requires_grad))
    np.savez_compressed(os.path.join(OUT, 'grad_B2_T16.npz'), **out)
    print('grad_B2_T16: |g_th|max=%.3e |g_sdf|max=%.3e |g_qc|max=%.3e |g_w|max=%.3e |g_eps|max=%.3e err.requires_grad=%s' % (
        float(grads[0].abs().max()), float(grads[3].abs().max()), float(grads[4].abs().max()), float(grads[5].abs().max()),
        float(grads[6].abs().max()), err.requires_grad))

    # 8. nonholonomic factor on one (T,6) trajectory (the only way the reference can run it)
    from diff_gpmp2.gpmp2.custom_factors import NonHolonomicFactor
    T = 12
    traj = r32(rng.uniform(-2.0, 2.0, (T, 6)))
    f = NonHolonomicFactor(3, torch.tensor(0.01), T)
    e, H = f.get_error_full(traj)
    np.savez_compressed(os.path.join(OUT, 'nonholonomic_T12.npz'), traj=traj.numpy().astype(np.float32),
                        err=e.numpy(), H=H.numpy(), inv_cov=f.get_inv_cov_full().numpy())
    print('nonholonomic: ok', tuple(e.shape), tuple(H.shape))


if __name__ == '__main__':
    main()
