"""Frame sharding across the GPUs of one box.

Frames are independent units (no arithmetic couples them in inference), so each rank reconstructs a
contiguous block of frames with replicated weights and NO data-path collective; the only communication
is an optional gather of per-frame scalars (NLL / log-det) at the end, over ``torch.distributed``
(NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def frame_shard(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of the contiguous block of frames owned by ``rank``; blocks differ by at most one frame."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def frame_seed(frame_id: int) -> int:
    """Per-frame generator seed of the streaming config (seed = frame id)."""
    return int(frame_id)


def gather_frame_scores(local: torch.Tensor, n_frames: int) -> torch.Tensor:
    """All-gather per-frame scalars (shape (n_local, k)) into frame order (n_frames, k) on every rank."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [frame_shard(n_frames, r, world) for r in range(world)]
    max_n = max(b - a for a, b in sizes)
    pad = torch.zeros((max_n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([bufs[r][: b - a] for r, (a, b) in enumerate(sizes)], 0)
