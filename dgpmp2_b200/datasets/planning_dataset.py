"""Planning dataset reader (format of reference ``diff_gpmp2/datasets/planning_dataset.py:15-70``):
``<root>/<mode>/meta.yaml``, ``im_sdf/<i>_im.png``, ``im_sdf/<i>_sdf.npy``,
``<label_subdir>/env_<i>_prob_<j>.npz`` -> dict sample {im (1,H,W), sdf (1,H,W), start (1,d),
goal (1,d), th_opt (T,d)}."""
import os

import numpy as np
import torch
import yaml
from torch.utils.data import Dataset

from ..utils.sdf_utils import rgb2gray


class PlanningDataset(Dataset):
    def __init__(self, root_dir, mode='train', num_envs=-1, num_env_probs=-1, label_subdir='opt_trajs_gpmp2'):
        self.root_dir = os.path.abspath(root_dir)
        self.subdir = os.path.join(root_dir, mode)
        self.imsdf_dir = os.path.join(self.subdir, 'im_sdf')
        self.label_dir = os.path.join(self.subdir, label_subdir)
        with open(os.path.join(self.subdir, 'meta.yaml')) as f:
            self.meta_data = yaml.safe_load(f)
        if 0 < num_envs <= self.meta_data['num_envs'] and 0 < num_env_probs <= self.meta_data['probs_per_env']:
            self.meta_data['num_envs'] = num_envs
            self.meta_data['probs_per_env'] = num_env_probs
        self.num_files = self.meta_data['num_envs'] * self.meta_data['probs_per_env']

    def __len__(self):
        return self.num_files

    def __getitem__(self, idx):
        ppe = self.meta_data['probs_per_env']
        env_idx, prob_idx = int(idx / ppe), int(idx % ppe)
        from PIL import Image
        im = np.asarray(Image.open(os.path.join(self.imsdf_dir, '%d_im.png' % env_idx)), dtype=np.float64) / 255.0
        if im.ndim > 2:
            im = rgb2gray(im)
        im = torch.tensor(np.array([im > 0.75], dtype=np.float64))
        sdf = torch.tensor(np.load(os.path.join(self.imsdf_dir, '%d_sdf.npy' % env_idx))[None])
        npf = np.load(os.path.join(self.label_dir, 'env_%d_prob_%d.npz' % (env_idx, prob_idx)))
        return {'im': im, 'sdf': sdf, 'start': torch.tensor(npf['start'][None]), 'goal': torch.tensor(npf['goal'][None]),
                'th_opt': torch.tensor(npf['th_opt'])}
