"""Multi-GPU plumbing: one process per GPU, envs sharded by contiguous index range, no per-step
communication. The only collective is an all-reduce(sum) of the 8-double episode-metric vector
(NCCL over NVLink on GPUs; gloo in the CPU tests).

Env i's reset key is split(PRNGKey(seed), total+1)[i+1] (the VmapGymWrapper._reset scheme,
/root/reference/po_brax/envs/wrappers.py:160-163); threefry split is counter based, so every rank
derives its own slice locally (pobrax_split_keys)."""
from typing import Tuple

import torch


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous env range [lo, hi) of `rank`; the first total % world ranks get one extra env."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError('bad shard arguments')
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_keys(env, seed: int, total: int, rank: int, world: int) -> torch.Tensor:
    """This rank's slice of split(PRNGKey(seed), total+1)[1:], computed on the device."""
    lo, hi = shard_range(total, rank, world)
    if hi - lo != env.batch_size:
        raise ValueError(f'env.batch_size {env.batch_size} != shard size {hi - lo}')
    key = ((seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF)
    return env.split_keys(key, total + 1, first=1 + lo, count=hi - lo)


def reduce_metric_vector(acc: torch.Tensor, world: int, group=None) -> torch.Tensor:
    """all-reduce(sum) of a copy of the accumulator vector; returns the global totals."""
    out = acc.clone()
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def reduce_metrics(state, world: int, group=None) -> torch.Tensor:
    acc = state.buf['acc']
    if acc is None:
        raise RuntimeError('create(..., eval_metrics=True) to maintain the device-side episode accumulators')
    return reduce_metric_vector(acc, world, group)
