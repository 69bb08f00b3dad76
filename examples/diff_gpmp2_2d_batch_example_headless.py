#!/usr/bin/env python
"""Headless counterpart of the reference's ``examples/diff_gpmp2_2d_batch_example.py``.

Same flow and the same ``diff_gpmp2.*`` imports (resolved by this repository's alias package onto the
B200 implementation): load the YAML parameters, read a batch from a ``PlanningDataset`` through a
``DataLoader``, build straight-line initial trajectories, run ``DiffGPMP2Planner.forward`` on the
batch.  Differences, all forced by what the reference ships: the dataset folder it reads is not in
the reference tree, so a small one is synthesised first in the reference's on-disk format; nothing is
plotted; and ``forward`` returns the 8-tuple of the reference's code (the reference script unpacks 6).

    python examples/diff_gpmp2_2d_batch_example_headless.py [--batch 4] [--steps 63]
"""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch
from torch.utils.data import DataLoader

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from diff_gpmp2.robot_models import PointRobot2D                                  # noqa: E402
from diff_gpmp2.gpmp2.diff_gpmp2_planner import DiffGPMP2Planner                  # noqa: E402
from diff_gpmp2.utils.helpers import load_params                                  # noqa: E402
from diff_gpmp2.utils.planner_utils import straight_line_trajb                    # noqa: E402
from diff_gpmp2.datasets import PlanningDataset                                   # noqa: E402
from dgpmp2_b200.datasets.synthetic import make_problems                          # noqa: E402
from dgpmp2_b200.datasets.writer import write_dataset                             # noqa: E402

PLANNER_YAML = """gpmp2:
  planner_params: {dof: 2, state_dim: 4, total_time_sec: 10, total_time_step: %d, use_vel_limits: False}
  gp_params: {Q_c_inv: [[1.0, 0.0], [0.0, 1.0]], K_s: 0.01, K_g: 0.01, K_v: 0.01, v_x: 1.0, v_y: 1.0}
  obs_params: {cost_sigma: 0.01, epsilon_dist: 0.4}
  optim_params: {method: gauss_newton, reg: 0.1, plan_time: inf, max_iters: 100, tol_err: 0.001, tol_delta: 0.0001}
"""


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=4)
    ap.add_argument('--steps', type=int, default=63)
    args = ap.parse_args(argv)
    torch.set_default_dtype(torch.float64)
    np.random.seed(0)
    torch.manual_seed(0)
    device = torch.device('cpu')          # tensors live on the host as in the reference example; compute runs on the GPU
    with tempfile.TemporaryDirectory() as tmp:
        for name, text in (('gpmp2_params.yaml', PLANNER_YAML % args.steps),
                           ('robot.yaml', 'type: point_robot\ndof: 2\nsphere_radius: [0.4]\n'),
                           ('env_params.yaml', 'dim: 2\nx_lims: [-5.0, 5.0]\ny_lims: [-5.0, 5.0]\n')):
            with open(os.path.join(tmp, name), 'w') as fp:
                fp.write(text)
        n_env = 2 * args.batch
        pr = make_problems(n_env, args.steps + 1, seed=0, dtype=torch.float64)
        write_dataset(tmp, pr['im'][:, 0].numpy(), pr['sdf'][:, 0].numpy(), pr['start'][:, 0].numpy(), pr['goal'][:, 0].numpy(),
                      pr['th_init'].numpy())
        dataset = PlanningDataset(tmp, mode='train', label_subdir='opt_trajs_gpmp2')
        loader = DataLoader(dataset, batch_size=args.batch, shuffle=True, num_workers=0)
        env_data, planner_params, gp_params, obs_params, optim_params, robot_data = load_params(
            os.path.join(tmp, 'gpmp2_params.yaml'), os.path.join(tmp, 'robot.yaml'), os.path.join(tmp, 'env_params.yaml'), device)
        sample = next(iter(loader))
    im_b, sdf_b, start_b, goal_b = sample['im'], sample['sdf'], sample['start'], sample['goal']
    env_params = {'x_lims': env_data['x_lims'], 'y_lims': env_data['y_lims']}
    robot = PointRobot2D(robot_data['sphere_radius'][0])
    th_init_b = straight_line_trajb(start_b, goal_b, planner_params['total_time_sec'], planner_params['total_time_step'],
                                    planner_params['dof'], device)
    planner = DiffGPMP2Planner(gp_params, obs_params, planner_params, optim_params, env_params, robot)
    th_finalb, _, err_initb, err_finalb, err_per_iterb, err_ext_per_iterb, jb, timeb = planner.forward(
        th_init_b, start_b, goal_b, im_b, sdf_b)
    for i in range(args.batch):
        print('problem %d: %3d iterations, cost %.4f -> %.4f' % (i, jb[i], err_initb[i], err_finalb[i]))
    print('batch planning time = %f (seconds)' % timeb[-1])
    return th_finalb, err_initb, err_finalb, jb


if __name__ == '__main__':
    main()
