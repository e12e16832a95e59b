"""torchrun --nproc-per-node N tools/check_peer_gather.py -- the fused render+gather (peer stores over NVLink from the
render epilogue) must produce exactly what render + NCCL all_gather produces, and both must equal the single-GPU render
of the whole bundle (strong scaling: contiguous shards of one batch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import neurad_studio_b200 as nsb
from neurad_studio_b200 import scene
from neurad_studio_b200.backend import B200Backend
from neurad_studio_b200.dist import PeerGatherBuffers, shard_range


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = nsb.small_config(n_actors=4, log2_main=16, log2_prop=14)
    trajs = scene.make_trajectories(cfg.n_actors, cfg.duration)
    params = scene.make_params(cfg, seed=3, beta=3.0, sdf_bias=0.5, trajectories=trajs)
    n_total = 4096 * world
    rays = scene.random_rays(n_total, cfg, seed=4, trajectories=trajs)
    a, b = shard_range(n_total, world, rank)
    mine = {k: v[a:b] for k, v in rays.items()}
    be = B200Backend(dev)
    be.load_params(cfg, params)
    # reference: render locally, NCCL all_gather
    out = be.render(mine)
    ref = {}
    for k in ("features", "depth", "accumulation"):
        full = torch.empty(world, b - a, out[k].shape[1], device=dev)
        dist.all_gather_into_tensor(full.view(-1), out[k].reshape(-1).contiguous())
        ref[k] = full
    # fused: the kernel stores every row into all peers' buffers
    pg = PeerGatherBuffers(b - a, cfg.feature_dim, dev)
    pg.bind(be)
    loc = pg.local()
    loc.update({k: torch.empty(b - a, 1, device=dev) for k in ("prop_depth_0", "prop_depth_1")})
    be.render(mine, out=loc)
    pg.barrier()
    torch.cuda.synchronize()
    be.check_status()
    ok = all(torch.equal(pg.buf[k], ref[k]) for k in ref)
    be.set_peer_outputs(None)
    # strong scaling: the gathered shards ARE the single-GPU render of the whole bundle (rays are independent and the
    # kernels are batch-composition independent), bit for bit -- also through ShardedOutputs (NCCL gather into the
    # buffer the kernel wrote its slice of)
    from neurad_studio_b200.dist import ShardedOutputs

    whole = be.render(rays)
    torch.cuda.synchronize()
    for k in ref:
        ok = ok and torch.equal(pg.buf[k].reshape(n_total, -1), whole[k].reshape(n_total, -1))
    so = ShardedOutputs(n_total, {"features": cfg.feature_dim, "depth": 1, "accumulation": 1}, dev)
    loc2 = so.local()
    loc2.update({k: torch.empty(b - a, 1, device=dev) for k in ("prop_depth_0", "prop_depth_1")})
    be.render(mine, out=loc2)
    gathered = so.gather()
    for k in ref:
        ok = ok and torch.equal(gathered[k], whole[k].reshape(n_total, -1))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("PEER_GATHER_OK" if int(flag.item()) == 1 else "PEER_GATHER_MISMATCH")
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
