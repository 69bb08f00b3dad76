#!/usr/bin/env python
"""Outer learning step through the differentiable planner (headless, synthetic data).

The reference trains a network that predicts the planner's covariances by back-propagating an imitation loss
through a few unrolled Gauss-Newton iterations (``diff_gpmp2/learning/train_planner.py:297-403``).  This script is
the same loop on this package, reduced to what the GN hot path needs:

  * expert labels: trajectories optimised to convergence by a planner with "expert" constants
    (``DiffGPMP2Planner.forward`` = one persistent CUDA launch for the whole batch);
  * learner: a planner with the default constants plus a small module whose RAW output goes straight into the
    fused launch (``planner.set_learn_module`` -> DGPMP2_FLAG_HEAD: the kernels form Qc^-1 = q^2 I and
    obscov_inv = o^2 themselves, the covariance tensors never exist);
  * K unrolled ``planner.step`` calls, loss = mean |th_K - th_expert|^2, ``loss.backward()`` through the CUDA
    backward kernel of every step;
  * data parallel: every rank owns a shard of the problems (no collective inside the GN loop) and the ONLY
    collective is one all-reduce of the module's flat gradient per optimiser step
    (``dgpmp2_b200.parallel.allreduce_gradients``; NCCL under torchrun, a no-op on one GPU).

    python examples/learn_covariances_headless.py [--batch 64] [--states 32] [--unroll 3] [--iters 20]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
           examples/learn_covariances_headless.py --batch 128
"""
import argparse
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from diff_gpmp2.robot_models import PointRobot2D                                  # noqa: E402
from diff_gpmp2.gpmp2.diff_gpmp2_planner import DiffGPMP2Planner                  # noqa: E402
from dgpmp2_b200 import parallel                                                  # noqa: E402
from dgpmp2_b200.datasets.synthetic import make_problems                          # noqa: E402


def make_planner(T, dtype, cost_sigma=0.01, qc=1.0, max_iters=100):
    gp = {'Q_c_inv': torch.eye(2, dtype=dtype) * qc, 'K_s': torch.tensor(0.01, dtype=dtype), 'K_g': torch.tensor(0.01, dtype=dtype)}
    ob = {'cost_sigma': torch.tensor(cost_sigma, dtype=dtype), 'epsilon_dist': torch.tensor(0.4, dtype=dtype)}
    pp = {'dof': 2, 'state_dim': 4, 'total_time_sec': 10.0, 'total_time_step': T - 1}
    op = {'method': 'gauss_newton', 'reg': 0.1, 'plan_time': 'inf', 'max_iters': max_iters, 'tol_err': 1e-3, 'tol_delta': 1e-4}
    env = {'x_lims': [-5.0, 5.0], 'y_lims': [-5.0, 5.0]}
    return DiffGPMP2Planner(gp, ob, pp, op, env, PointRobot2D(torch.tensor(0.4, dtype=dtype)))


class CovarianceHead(nn.Module):
    """th (B,T,d) -> out (B,1,(T-1)+T) = [q | o] for dynamics_mode 'diag_identity' (diff_gpmp2_planner.py:254-262):
    the planner's own constants (q = 1 -> Qc^-1 = I, o = 1/sigma -> obscov_inv = 1/sigma^2) scaled by exp(linear map of
    the current trajectory), so that it starts exactly at the hand-set planner and moves in relative steps."""

    def __init__(self, T, d, cost_sigma, dtype):
        super().__init__()
        self.lin = nn.Linear(T * d, (T - 1) + T, dtype=dtype)
        nn.init.zeros_(self.lin.weight)
        nn.init.zeros_(self.lin.bias)
        base = torch.ones((T - 1) + T, dtype=dtype)
        base[T - 1:] = 1.0 / cost_sigma
        self.register_buffer('base', base)

    def forward(self, th, im, sdf):
        # bounded relative steps: every raw output stays within exp(+-1) of the planner's constant, whatever the optimiser
        # does (an unbounded exp() of T*d inputs overflows the obstacle weight after one Adam step at T = 64)
        x = th.reshape(th.shape[0], -1) / float(th.shape[1] * th.shape[2])
        return (self.base * torch.exp(torch.tanh(self.lin(x)))).unsqueeze(1)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64, help='problems over all ranks')
    ap.add_argument('--states', type=int, default=32)
    ap.add_argument('--unroll', type=int, default=3, help='GN iterations unrolled on the autograd tape')
    ap.add_argument('--iters', type=int, default=20, help='optimiser steps')
    ap.add_argument('--lr', type=float, default=0.02)
    ap.add_argument('--device', type=str, default='cuda', help="where the tensors live ('cpu': staged to the GPU per call)")
    ap.add_argument('--f64', action='store_true')
    ap.add_argument('--ext-weight', type=float, default=1e-3,
                    help='weight of the external loss gp + sg + lambda * obs of the final iterate (train_planner.py:327-346)')
    ap.add_argument('--ext-obs-lambda', type=float, default=1.0)
    ap.add_argument('--json-out', dest='log', type=str, default=None, help='write per-step timings / checks as JSON (rank 0)')
    args = ap.parse_args(argv)
    rank, world, local_rank = parallel.init_distributed()
    dtype = torch.float64 if args.f64 else torch.float32
    dev = torch.device(args.device, local_rank) if args.device == 'cuda' else torch.device('cpu')
    torch.manual_seed(0)
    T = args.states

    pr = make_problems(args.batch, T, im_size=64, seed=0, dtype=dtype)
    th0, start, goal, sdf, im = parallel.shard_batch([pr[k] for k in ('th_init', 'start', 'goal', 'sdf', 'im')], rank, world)
    th0, start, goal, sdf, im = (t.to(dev).contiguous() for t in (th0, start, goal, sdf, im))

    # expert labels: a planner that trusts the obstacle term less and the GP prior more
    with torch.no_grad():
        th_expert = make_planner(T, dtype, cost_sigma=0.03, qc=2.0).forward(th0, start, goal, im, sdf)[0]

    planner = make_planner(T, dtype)
    head = CovarianceHead(T, 4, 0.01, dtype).to(dev)
    planner.set_learn_module(head, 'diag_identity')
    opt = torch.optim.Adam(head.parameters(), lr=args.lr)
    losses, log = [], []
    on_gpu = dev.type == 'cuda'
    import torch.distributed as dist
    n_allreduce = [0]
    if dist.is_initialized():                                   # count the collectives issued per optimiser step
        real_all_reduce = dist.all_reduce

        def counting_all_reduce(*a, **k):
            n_allreduce[0] += 1
            return real_all_reduce(*a, **k)
        dist.all_reduce = counting_all_reduce

    def ev():
        e = torch.cuda.Event(enable_timing=True) if on_gpu else None
        if e is not None:
            e.record()
        return e
    for it in range(args.iters):
        opt.zero_grad()
        th = th0
        t0 = ev()
        for _ in range(args.unroll):                       # truncated back-propagation through the planner
            dth = planner.step(th, start, goal, im, sdf)[0]
            th = th + dth
        loss = ((th - th_expert) ** 2).mean()
        if args.ext_weight > 0.0:
            # the reference's external loss on the final iterate: differentiable unweighted errors (train_planner.py:327-346)
            e_sg, e_gp, e_obs = planner.unweighted_errors_batch(th, sdf)
            loss = loss + args.ext_weight * (e_gp.mean() + e_sg.mean() + args.ext_obs_lambda * e_obs.mean())
        t1 = ev()
        loss.backward()
        t2 = ev()
        n_allreduce[0] = 0
        n = parallel.allreduce_gradients(head.parameters())          # the only collective: one flat bucket
        t3 = ev()
        opt.step()
        losses.append(float(loss.detach()))
        rec = {'iter': it, 'loss': losses[-1], 'grad_elements': n, 'all_reduce_calls': n_allreduce[0]}
        if on_gpu:
            torch.cuda.synchronize()
            rec.update(forward_ms=t0.elapsed_time(t1), backward_ms=t1.elapsed_time(t2), allreduce_ms=t2.elapsed_time(t3))
        if world > 1:
            assert n_allreduce[0] == 1, 'expected exactly one all-reduce per optimiser step, saw %d' % n_allreduce[0]
            flat = torch.cat([p.detach().reshape(-1) for p in head.parameters()])
            lo, hi = flat.clone(), flat.clone()
            real_all_reduce(lo, op=dist.ReduceOp.MIN)
            real_all_reduce(hi, op=dist.ReduceOp.MAX)
            assert torch.equal(lo, hi), 'parameters differ between ranks after the optimiser step'
            rec['params_identical_on_all_ranks'] = True
        log.append(rec)
        if rank == 0:
            print('iter %3d  loss %.6f  (%d gradient elements all-reduced over %d rank(s)%s)' % (
                it, losses[-1], n, world,
                ', fwd %.2f ms bwd %.2f ms all-reduce %.3f ms' % (rec['forward_ms'], rec['backward_ms'], rec['allreduce_ms']) if on_gpu else ''))
    if dist.is_initialized():
        dist.all_reduce = real_all_reduce
    if args.log and rank == 0:
        import json
        with open(args.log, 'w') as f:
            json.dump({'world': world, 'batch_total': args.batch, 'batch_per_rank': int(th0.shape[0]), 'states': T, 'unroll': args.unroll,
                       'backend': (dist.get_backend() if dist.is_initialized() else None), 'steps': log}, f, indent=1)
    return losses, head


if __name__ == '__main__':
    main()
