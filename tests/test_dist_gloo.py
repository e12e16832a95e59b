"""CPU, world_size 2, gloo: the N>1 host logic (ray sharding + per-image gather).  The per-rank "render" is a
deterministic stand-in computed with torch on CPU -- the CUDA kernels are covered by the GPU tests; this checks that
gather(shards) reproduces the unsharded result in the original ray order, including uneven shards."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neurad_studio_b200.dist import ShardedOutputs, shard_range, shard_sizes


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 230400, 1497601):
        for world in (1, 2, 3, 8):
            edges = [shard_range(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sz = shard_sizes(n, world)
            assert max(sz) - min(sz) <= 1


def _fake_render(rays):
    return {"features": torch.stack([rays.sum(-1) * (i + 1) for i in range(5)], -1), "depth": rays[:, :1] * 2.0}


def _worker(rank, world, port, n_rays, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(0)
        rays = torch.rand(n_rays, 3, generator=gen)  # every rank knows the camera -> generates the same rays
        a, b = shard_range(n_rays, world, rank)
        so = ShardedOutputs(n_rays, {"features": 5, "depth": 1}, "cpu")
        loc = so.local()
        res = _fake_render(rays[a:b])
        for k in loc:
            loc[k].copy_(res[k])
        full = so.gather()
        ref = _fake_render(rays)
        ok = all(torch.equal(full[k], ref[k]) for k in ref)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rays", [1000, 1001])
def test_sharded_render_gather_world2(n_rays):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 500) + n_rays % 7
    mp.spawn(_worker, args=(world, port, n_rays, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)
