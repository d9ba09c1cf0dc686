"""Host logic of the multi-GPU layer on CPU: shard arithmetic and the gloo all-gather of the
scalar table with world_size 2 and 3 (uneven shards)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from srsran_ce_pytorch_b200 import sharding


@pytest.mark.parametrize("n,world,align", [(256, 2, 4), (1000, 8, 4), (12, 5, 4), (8, 8, 1), (4, 8, 4), (0, 4, 4)])
def test_shard_range_partitions(n, world, align):
    parts = [sharding.shard_range(n, r, world, align) for r in range(world)]
    assert parts[0][0] == 0 and parts[-1][1] == n
    for (b0, e0), (b1, e1) in zip(parts[:-1], parts[1:]):
        assert e0 == b1
    sizes = [e - b for b, e in parts]
    assert all(s % align == 0 for s in sizes) and max(sizes) - min(sizes) <= align


def test_shard_range_rejects_misaligned():
    with pytest.raises(ValueError):
        sharding.shard_range(10, 0, 2, 4)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_units, align, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_units * 6, dtype=torch.float64).reshape(n_units, 6)
        b, e = sharding.shard_range(n_units, rank, world, align)
        got = sharding.gather_scalars(full[b:e].clone(), n_units, align)
        ret[rank] = bool(torch.equal(got, full))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_units,align", [(2, 64, 4), (3, 40, 4)])
def test_gather_scalars_gloo(world, n_units, align):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, align, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret[r] for r in range(world))


def test_gather_scalars_single_process_is_identity():
    t = torch.randn(8, 6, dtype=torch.float64)
    assert sharding.gather_scalars(t, 8, 4) is t
