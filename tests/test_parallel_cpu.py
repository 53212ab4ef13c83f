"""The N>1 path on CPU: contiguous env sharding, counter-based key slices and the episode-metric
all-reduce, with world_size 2 over gloo (the GPU path uses the same code over NCCL)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import threefry as tf
from po_brax_b200.parallel import reduce_metric_vector, shard_range


def test_shard_ranges_partition_the_env_axis():
    for total, world in ((1 << 20, 8), (1000, 3), (7, 8), (128, 1)):
        r = [shard_range(total, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == total
        assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_key_slices_are_local():
    """Env i's key is split(PRNGKey(seed), total+1)[i+1]: each rank's slice, computed independently,
    concatenates to the single-process key table (wrappers.py:160-163 scheme)."""
    total, world = 1000, 4
    full = tf.split(tf.prng_key(5), total + 1)[1:]
    parts = []
    for k in range(world):
        lo, hi = shard_range(total, k, world)
        # what pobrax_split_keys computes for (n = total + 1, first = 1 + lo, count = hi - lo)
        parts.append(tf.split(tf.prng_key(5), total + 1)[1 + lo:1 + hi])
    assert (np.concatenate(parts) == full).all()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    acc = torch.arange(8, dtype=torch.float64) * (rank + 1)  # this rank's accumulator vector
    tot = reduce_metric_vector(acc, world)
    assert torch.equal(acc, torch.arange(8, dtype=torch.float64) * (rank + 1))  # input untouched
    if rank == 0:
        torch.save(tot, out)
    dist.destroy_process_group()


def test_metric_allreduce_gloo_world2(tmp_path):
    out = str(tmp_path / 'tot.pt')
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    tot = torch.load(out)
    assert torch.equal(tot, torch.arange(8, dtype=torch.float64) * 3)
