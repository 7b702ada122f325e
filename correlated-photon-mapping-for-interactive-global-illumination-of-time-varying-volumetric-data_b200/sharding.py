"""Multi-GPU plumbing of the photon path: one process per GPU, photons sharded by
index range, one exchange step.  torch.distributed is the transport (NCCL on GPUs, gloo in the CPU tests);
no compute lives here.

Photon i of rank r is photon r * photons_per_gpu + i of the global photon set: it uses that MWC64X stream and
that host base offset (cpmh_runtime_init / cpm_rng_host_base_offsets_range), so the union of all ranks'
photon records is bit-identical to one GPU tracing the whole set.  Every rank splats its own photons into its
own light volume; the frame's result is the SUM over ranks, produced out of place -- the per-rank volume must
stay local because the next frame updates it incrementally (-old/+new of the rank's re-traced photons).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def photon_shard(rank: int, world: int, photons_per_gpu: int):
    """(first global photon id, count) owned by `rank` under weak scaling"""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return rank * photons_per_gpu, photons_per_gpu


def strong_shard(rank: int, world: int, total: int):
    """(first, count) of a fixed photon set split as evenly as possible (ranks < total % world get one more)"""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def allreduce_light_volume(local: torch.Tensor, out: torch.Tensor | None = None, group=None) -> torch.Tensor:
    """out = sum over ranks of `local`; `local` is left untouched.  Single process: returns `local`."""
    if not is_distributed():
        return local
    if out is None or out.shape != local.shape or out.dtype != local.dtype or out.device != local.device:
        out = torch.empty_like(local)
    out.copy_(local)
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def max_over_ranks(values, device="cpu"):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def sum_over_ranks(values, device="cpu"):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]
