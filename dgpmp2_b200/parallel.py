"""Multi-GPU plumbing: shard the planning batch, one process per GPU, no collective in the GN loop.

The problems of a batch share nothing (each has its own SDF, start, goal and trajectory; the
block-tridiagonal coupling is along t inside one problem), so rank g of G simply owns the problems
[lo, hi) of ``shard_range`` with their SDFs resident on its GPU.  The only collective of the whole
system is the all-reduce of the outer learning gradient (reference training loop,
``learning/train_planner.py:366-403``), done once per optimizer step over ONE flat bucket.
``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is the transport.
"""
import os
from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) of n items for ``rank`` of ``world`` (sizes differ by at most one)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError('bad rank/world %d/%d' % (rank, world))
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors: Sequence[Optional[torch.Tensor]], rank: int, world: int) -> List[Optional[torch.Tensor]]:
    """Slice every (B, ...) tensor to this rank's problems (views, no copies)."""
    out = []
    for t in tensors:
        if t is None:
            out.append(None)
            continue
        lo, hi = shard_range(t.shape[0], rank, world)
        out.append(t[lo:hi])
    return out


def init_distributed(backend: Optional[str] = None):
    """Initialise the default process group from the torchrun environment (RANK, WORLD_SIZE,
    LOCAL_RANK, MASTER_ADDR, MASTER_PORT).  Returns (rank, world, local_rank)."""
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device('cuda', local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def allreduce_gradients(params: Iterable[torch.nn.Parameter], average: bool = True, group=None) -> int:
    """Sum (or average) the ``.grad`` of ``params`` over all ranks with a SINGLE all-reduce of one flat
    bucket; parameters without a gradient contribute zeros.  Returns the number of elements reduced."""
    params = [p for p in params if p.requires_grad]
    if not params:
        return 0
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return sum(p.numel() for p in params)
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return off


def gather_batch(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather the per-rank slices of a batch-sharded (B_local, ...) tensor back into (total, ...)
    in problem order (statistics / iteration counts after a sharded solve; not on the hot path)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(total, r, world) for r in range(world)]
    nmax = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((nmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([bufs[r][:hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
