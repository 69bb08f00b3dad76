"""Deterministic synthetic random-obstacle 2-D planning problems (host-side input prep).

The reference ships no dataset files; its generator scripts draw axis-aligned
rectangles on a 128x128 occupancy map ("forest": many 4-5 px rectangles,
"multi_obs": a few 16-26 px rectangles; reference
``diff_gpmp2/datasets/generate_2d_dataset.py:41-66`` and
``obst_generator.py:179-221``), build the signed distance field with
``sdf_2d(map, padlen=0, res=cell)`` (``generate_2d_dataset.py:211``) and sample
start / goal uniformly in ``[lims+1, lims-1]^2`` at least 0.6 of the workspace
diagonal apart (``:138,152,172-179``).  This module reproduces that recipe (not
its exact random stream) with a private ``numpy`` generator so benchmark and
test inputs are reproducible from a seed.  Everything here runs on the host
before the hot path; nothing is timed.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np
import torch

from ..utils.sdf_utils import sdf_2d


def random_obstacle_map(rng: np.random.Generator, im_size: int, kind: str,
                        keep_free_px: Sequence[Tuple[float, float]] = (), free_radius: int = 8) -> np.ndarray:
    """Occupancy image, 1.0 = free (white), 0.0 = obstacle, row 0 = y_max."""
    occ = np.zeros((im_size, im_size), dtype=np.float64)
    if kind == 'forest':
        n_obs = int(rng.integers(23, 45))
        lo = max(1, im_size // 30)
        hi = lo + 1
        x0, x1 = 0, im_size - 1
    elif kind == 'multi_obs':
        n_obs = int(rng.integers(2, 5))
        lo = max(1, im_size // 8)
        hi = lo + 10
        x0, x1 = int(0.1 * im_size), int(0.9 * im_size)
    else:
        raise ValueError('unknown map kind %r' % (kind,))
    placed = 0
    tries = 0
    while placed < n_obs and tries < 50 * n_obs:
        tries += 1
        w = int(rng.integers(lo, hi + 1))
        h = int(rng.integers(lo, hi + 1))
        cx = int(rng.integers(x0, x1 + 1))
        cy = int(rng.integers(x0, x1 + 1))
        c0, c1 = max(0, cx - w // 2), min(im_size, cx + (w + 1) // 2)
        r0, r1 = max(0, cy - h // 2), min(im_size, cy + (h + 1) // 2)
        if c1 <= c0 or r1 <= r0:
            continue
        bad = False
        for (px, py) in keep_free_px:
            if (c0 - free_radius <= px <= c1 + free_radius) and (r0 - free_radius <= py <= r1 + free_radius):
                bad = True
                break
        if bad:
            continue
        occ[r0:r1, c0:c1] = 1.0
        placed += 1
    return 1.0 - occ


def sample_start_goal(rng: np.random.Generator, x_lims, y_lims, dist_factor: float = 0.6):
    lo_x, hi_x = x_lims[0] + 1.0, x_lims[1] - 1.0
    lo_y, hi_y = y_lims[0] + 1.0, y_lims[1] - 1.0
    min_d = dist_factor * np.hypot(x_lims[1] - x_lims[0], y_lims[1] - y_lims[0])
    while True:
        s = np.array([rng.uniform(lo_x, hi_x), rng.uniform(lo_y, hi_y)])
        g = np.array([rng.uniform(lo_x, hi_x), rng.uniform(lo_y, hi_y)])
        if np.hypot(*(g - s)) >= min_d:
            return s, g


def straight_line(start_conf: np.ndarray, goal_conf: np.ndarray, total_time_sec: float, T: int) -> np.ndarray:
    """Constant-velocity straight line (reference ``planner_utils.py:47-56``), (T, 2*dof)."""
    dof = start_conf.shape[-1]
    n = T - 1
    i = np.arange(T, dtype=np.float64)[:, None]
    pos = start_conf[None, :] * (n - i) * 1.0 / n * 1.0 + goal_conf[None, :] * i * 1.0 / n * 1.0
    vel = np.broadcast_to((goal_conf - start_conf) / total_time_sec * 1.0, (T, dof))
    return np.concatenate((pos, vel), axis=1)


def make_problems(B: int, T: int, dof: int = 2, im_size: int = 128, seed: int = 0,
                  x_lims=(-5.0, 5.0), y_lims=(-5.0, 5.0), total_time_sec: float = 10.0,
                  unique_envs: int = 0, dtype=torch.float32, heading: bool = None) -> Dict[str, torch.Tensor]:
    """B synthetic problems: ``im, sdf (B,1,H,W)``, ``start, goal (B,1,d)``, ``th_init (B,T,d)``.

    ``unique_envs`` > 0 draws only that many distinct maps and reuses them round-robin
    with fresh start/goal samples (keeps host prep time bounded for very large B).
    dof 3 adds a heading state (start heading 0, goal heading pi/2, as the reference's
    ``examples/diff_gpmp2_nonholonomic_example.py:44-46``).
    """
    rng = np.random.default_rng(seed)
    d = 2 * dof
    cell = (x_lims[1] - x_lims[0]) / im_size
    n_env = B if unique_envs <= 0 else min(B, unique_envs)
    ims = np.zeros((B, 1, im_size, im_size), dtype=np.float64)
    sdfs = np.zeros((B, 1, im_size, im_size), dtype=np.float64)
    start = np.zeros((B, 1, d), dtype=np.float64)
    goal = np.zeros((B, 1, d), dtype=np.float64)
    th = np.zeros((B, T, d), dtype=np.float64)
    env_cache = []
    for b in range(B):
        s, g = sample_start_goal(rng, x_lims, y_lims)
        if b < n_env:
            spx = ((s[0] - x_lims[0]) / cell, (y_lims[1] - s[1]) / cell)
            gpx = ((g[0] - x_lims[0]) / cell, (y_lims[1] - g[1]) / cell)
            kind = 'forest' if rng.random() < 0.5 else 'multi_obs'
            im = random_obstacle_map(rng, im_size, kind, keep_free_px=(spx, gpx))
            sdf = sdf_2d(im, padlen=0, res=cell)
            env_cache.append((im, sdf))
        else:
            im, sdf = env_cache[b % n_env]
        ims[b, 0] = im
        sdfs[b, 0] = sdf
        sc = np.zeros(dof)
        gc = np.zeros(dof)
        sc[:2] = s
        gc[:2] = g
        if dof == 3:
            gc[2] = np.pi / 2.0
        start[b, 0, :dof] = sc
        goal[b, 0, :dof] = gc
        th[b] = straight_line(sc, gc, total_time_sec, T)
    out = {'im': ims, 'sdf': sdfs, 'start': start, 'goal': goal, 'th_init': th}
    return {k: torch.from_numpy(v).to(dtype) for k, v in out.items()}
