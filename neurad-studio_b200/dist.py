"""Multi-GPU plumbing: rays shard across ranks with no data-path collective; the only communication is the
per-image gather of the per-ray outputs (torch.distributed, NCCL over NVLink on GPUs, gloo in the CPU tests).

The reference has no inference-time multi-GPU path at all: eval/render run on one device
per process (pipelines/ad_pipeline.py:197-306).  This module is therefore new functionality with a single-GPU
result to reproduce: gather(render(shard_r) for r in ranks) == render(all rays).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_range(n_rays: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced row-major shard [start, end) of a flattened ray bundle (rays.py:300-311): the first
    n % world ranks take one extra ray.  Contiguous image rows keep the hash-grid accesses of a rank coherent."""
    base, rem = divmod(n_rays, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(n_rays: int, world: int):
    return [shard_range(n_rays, world, r)[1] - shard_range(n_rays, world, r)[0] for r in range(world)]


class ShardedOutputs:
    """Full-size output buffers of which every rank fills its own slice; `gather()` completes them everywhere.

    The render kernel writes straight into the slice (no staging copy): the buffer IS the all-gather buffer.
    Uneven shards are padded to the largest shard so that a single all_gather_into_tensor per output suffices."""

    def __init__(self, n_rays: int, widths: Dict[str, int], device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_rays = n_rays
        self.sizes = shard_sizes(n_rays, self.world)
        self.pad = max(self.sizes)
        self.buf = {k: torch.zeros(self.world, self.pad, w, device=device) for k, w in widths.items()}

    def local(self) -> Dict[str, torch.Tensor]:
        """This rank's output slices ([n_local, width], contiguous) to hand to `B200Backend.render(out=...)`."""
        n = self.sizes[self.rank]
        return {k: b[self.rank, :n] for k, b in self.buf.items()}

    def gather(self) -> Dict[str, torch.Tensor]:
        """All-gather every output; returns [n_rays, width] tensors in the original ray order."""
        out = {}
        for k, b in self.buf.items():
            if self.world > 1:
                dist.all_gather_into_tensor(b.view(-1), b[self.rank].reshape(-1).clone() if b.device.type == "cpu" else b[self.rank].reshape(-1), group=self.group)
            out[k] = torch.cat([b[r, : self.sizes[r]] for r in range(self.world)], dim=0)
        return out


class PeerGatherBuffers:
    """Gather buffers in symmetric (peer-mapped) memory for the fused render+gather path.

    Every rank allocates `[world, rows_per_rank, width]` per output through torch.distributed._symmetric_memory and
    exchanges the mappings once (`rendezvous`).  `bind(backend)` then tells the render kernel to store each finished
    row into ALL ranks' buffers (its own slice locally, the others over NVLink), so that after `barrier()` every rank
    holds the complete gather -- no NCCL collective on the data path."""

    WIDTHS = ("features", "depth", "accumulation")

    def __init__(self, rows_per_rank: int, feature_dim: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.rows = rows_per_rank
        widths = {"features": feature_dim, "depth": 1, "accumulation": 1}
        self.buf, self.hdl = {}, {}
        for k, w in widths.items():
            t = symm_mem.empty(self.world * rows_per_rank * w, dtype=torch.float32, device=device)
            self.hdl[k] = symm_mem.rendezvous(t, self.group)
            self.buf[k] = t.view(self.world, rows_per_rank, w)

    def local(self) -> Dict[str, torch.Tensor]:
        return {k: b[self.rank] for k, b in self.buf.items()}

    def bind(self, backend) -> None:
        ptrs = {k: [int(p) for p in self.hdl[k].buffer_ptrs] for k in self.buf}
        backend.set_peer_outputs(ptrs, self_rank=self.rank, row_offset=self.rank * self.rows)

    def barrier(self) -> None:
        """Device-side cross-rank barrier on the current stream (all ranks' rows have landed once it completes)."""
        self.hdl["features"].barrier()
