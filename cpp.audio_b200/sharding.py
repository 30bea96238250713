"""Multi-GPU plumbing for the hot path: one process per GPU, independent units per rank, no data-path collective.

Units and how they split:
  * independent streams (config 4)   -> contiguous stream ranges per rank
  * one long stream (config 2)       -> contiguous frame ranges per rank; a rank needs the samples of its frames, i.e. its
                                        span plus a (window - hop) halo that it simply reads from the input (no exchange)
  * voices (config 3)                -> contiguous voice ranges per rank
The only collective is the final gather of results (peak lists, or a sum of mixed buses) with torch.distributed
(NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `n_units` for `rank`; sizes differ by at most one, earlier ranks get the larger shards."""
    base, rem = divmod(n_units, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def frame_shard(n_samples: int, window: int, hop: int, rank: int, world: int) -> Tuple[int, int, int, int]:
    """Frame range [f0, f1) of one stream for `rank` and the sample span [s0, s1) those frames read (halo included)."""
    frames = 0 if n_samples < window else 1 + (n_samples - window) // hop
    f0, f1 = shard_range(frames, rank, world)
    if f1 <= f0:
        return f0, f1, 0, 0
    return f0, f1, f0 * hop, (f1 - 1) * hop + window


def gather_rows(local: torch.Tensor, counts: List[int] | None = None) -> torch.Tensor:
    """All-gather row blocks of different lengths (dim 0) into one tensor on every rank, in rank order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    if counts is None:
        n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        all_n = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(all_n, n)
        counts = [int(t.item()) for t in all_n]
    width = max(counts)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def reduce_bus(bus: torch.Tensor) -> torch.Tensor:
    """Sum of the mixed output buses of voice shards (the single exchange step of a voice-sharded mix)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(bus, op=dist.ReduceOp.SUM)
    return bus
