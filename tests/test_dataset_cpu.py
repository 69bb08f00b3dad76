"""CPU: the dataset writer/reader pair reproduces the reference's on-disk format and sample dict."""
import numpy as np
import torch

from dgpmp2_b200.datasets.synthetic import make_problems
from dgpmp2_b200.datasets.writer import write_dataset


def test_dataset_roundtrip_in_reference_format(tmp_path):
    from diff_gpmp2.datasets import PlanningDataset
    pr = make_problems(3, 8, im_size=32, seed=1, dtype=torch.float64)
    sub = write_dataset(str(tmp_path), pr['im'][:, 0].numpy(), pr['sdf'][:, 0].numpy(), pr['start'][:, 0].numpy(),
                        pr['goal'][:, 0].numpy(), pr['th_init'].numpy())
    for f in ('meta.yaml', 'im_sdf/0_im.png', 'im_sdf/2_sdf.npy', 'opt_trajs_gpmp2/env_1_prob_0.npz'):
        assert (tmp_path / 'train' / f).exists()
    ds = PlanningDataset(str(tmp_path), mode='train')
    assert len(ds) == 3
    s = ds[1]
    assert s['im'].shape == (1, 32, 32) and s['sdf'].shape == (1, 32, 32)
    assert s['start'].shape == (1, 4) and s['goal'].shape == (1, 4) and s['th_opt'].shape == (8, 4)
    np.testing.assert_array_equal(s['sdf'][0].numpy(), pr['sdf'][1, 0].numpy())
    np.testing.assert_array_equal(s['im'][0].numpy(), (pr['im'][1, 0].numpy() > 0.75).astype(np.float64))
    np.testing.assert_array_equal(s['th_opt'].numpy(), pr['th_init'][1].numpy())
