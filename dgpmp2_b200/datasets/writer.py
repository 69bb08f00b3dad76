"""Write a planning dataset in the reference's on-disk format (the layout produced by reference
``diff_gpmp2/datasets/generate_2d_dataset.py:244-275`` and read by ``planning_dataset.py:48-70``):

    <root>/<mode>/meta.yaml                                  {num_envs, probs_per_env, im_size, env_params}
    <root>/<mode>/im_sdf/<i>_im.png, <i>_sdf.npy              occupancy image (white = free) and SDF
    <root>/<mode>/<label_subdir>/env_<i>_prob_<j>.npz         start, goal, th_opt, th_init
"""
import os

import numpy as np
import yaml


def write_dataset(root_dir, ims, sdfs, starts, goals, th_opts, th_inits=None, mode='train',
                  label_subdir='opt_trajs_gpmp2', env_params=None, probs_per_env=1):
    """ims, sdfs: (E,H,W); starts, goals: (E*P, d); th_opts: (E*P, T, d) with P = probs_per_env."""
    from PIL import Image
    sub = os.path.join(root_dir, mode)
    os.makedirs(os.path.join(sub, 'im_sdf'), exist_ok=True)
    os.makedirs(os.path.join(sub, label_subdir), exist_ok=True)
    ims, sdfs = np.asarray(ims), np.asarray(sdfs)
    E = ims.shape[0]
    for i in range(E):
        Image.fromarray((np.clip(ims[i], 0.0, 1.0) * 255).astype(np.uint8), mode='L').save(os.path.join(sub, 'im_sdf', '%d_im.png' % i))
        np.save(os.path.join(sub, 'im_sdf', '%d_sdf' % i), sdfs[i])
        for j in range(probs_per_env):
            k = i * probs_per_env + j
            np.savez(os.path.join(sub, label_subdir, 'env_%d_prob_%d' % (i, j)), start=np.asarray(starts[k]),
                     goal=np.asarray(goals[k]), th_opt=np.asarray(th_opts[k]),
                     th_init=np.asarray(th_inits[k] if th_inits is not None else th_opts[k]))
    meta = {'num_envs': int(E), 'probs_per_env': int(probs_per_env), 'im_size': int(ims.shape[-1]),
            'env_params': env_params or {'x_lims': [-5.0, 5.0], 'y_lims': [-5.0, 5.0]}}
    with open(os.path.join(sub, 'meta.yaml'), 'w') as fp:
        yaml.safe_dump(meta, fp)
    return sub
