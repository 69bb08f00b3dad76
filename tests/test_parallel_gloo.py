"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding, gather, and the single
gradient all-reduce.  The per-rank compute is the CPU oracle here (there is no GPU in this container);
on the GPU box the same helpers feed the CUDA path (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dgpmp2_b200.parallel import allreduce_gradients, gather_batch, shard_batch, shard_range


def test_shard_range_partitions_exactly():
    for n in [0, 1, 7, 8, 1024, 8191]:
        for w in [1, 2, 3, 8]:
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from oracle import gn_oracle
        from tests.helpers import load_golden, oracle_params, golden_weights, t64
        g = load_golden('step_learned_B3_T16')           # 3 problems over 2 ranks: ragged shards (2 + 1)
        p = oracle_params(g['T'])
        qc, w, eps, q_full = golden_weights(g, p)
        full = [t64(g['th']), t64(g['start']), t64(g['goal']), t64(g['sdf']), qc, w, eps]
        th, start, goal, sdf, qc_l, w_l, eps_l = shard_batch(full, rank, world)
        dth, err, err_ext = gn_oracle.gn_step(th, start, goal, sdf, qc_l, w_l, eps_l, p, q_full)
        dth_all = gather_batch(dth, 3)
        err_all = gather_batch(err, 3)
        # one flat all-reduce of "learning" gradients
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
        loss = (net(th.reshape(-1, 4).float()) ** 2).sum() * (rank + 1)
        loss.backward()
        local = [p_.grad.clone() for p_ in net.parameters()]
        calls = {'n': 0}
        orig = dist.all_reduce

        def counting(*a, **k):
            calls['n'] += 1
            return orig(*a, **k)
        dist.all_reduce = counting
        n = allreduce_gradients(net.parameters(), average=True)
        dist.all_reduce = orig
        torch.save({'dth': dth_all, 'err': err_all, 'grads': [p_.grad for p_ in net.parameters()], 'local': local,
                    'n': n, 'calls': calls['n']}, os.path.join(out_dir, 'rank%d.pt' % rank))
    finally:
        dist.destroy_process_group()


def test_sharded_step_and_single_gradient_allreduce(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / 'rank0.pt')
    r1 = torch.load(tmp_path / 'rank1.pt')
    from tests.helpers import load_golden
    g = load_golden('step_learned_B3_T16')
    # gathered sharded result == the unsharded reference result, in problem order, on every rank
    for r in (r0, r1):
        np.testing.assert_allclose(r['dth'].numpy(), g['dth'], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(r['err'].numpy(), g['err'], rtol=1e-12)
        assert r['calls'] == 1                       # exactly one collective for all parameters
        assert r['n'] == 4 * 3 + 3 + 3 * 2 + 2
    for ga, gb, la, lb in zip(r0['grads'], r1['grads'], r0['local'], r1['local']):
        assert torch.equal(ga, gb)
        torch.testing.assert_close(ga, (la + lb) / 2)
