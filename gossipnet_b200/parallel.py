"""Multi-GPU plumbing: one process per GPU, images shard over ranks.

The forward needs no collective (images are independent); training
all-reduces ONE flat fp32 buffer (gradients + image count) per step.  The same
helpers run under gloo on CPU tensors (tests/test_dist_gloo.py) and under NCCL
on the GPUs.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Join the process group torchrun describes (RANK / WORLD_SIZE / MASTER_*);
    returns (rank, world, local_rank).  No-op for a single process."""
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        kw = {}
        if backend == 'nccl':
            torch.cuda.set_device(local)
            kw['device_id'] = torch.device('cuda', local)
        dist.init_process_group(backend, **kw)
    return rank, world, local


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard(items, rank_=None, world_=None):
    """Contiguous, balanced split of a list of images: rank r gets
    items[lo:hi] with sizes differing by at most one."""
    r = rank() if rank_ is None else rank_
    w = world() if world_ is None else world_
    n = len(items)
    lo = (n * r) // w
    hi = (n * (r + 1)) // w
    return items[lo:hi]


def allreduce_sum_(flat):
    """In-place sum over ranks of one flat buffer (no-op for a single process)."""
    if world() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def barrier():
    """All ranks wait for each other (no-op for a single process)."""
    if world() > 1:
        dist.barrier()


def max_over_ranks(value, device=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
