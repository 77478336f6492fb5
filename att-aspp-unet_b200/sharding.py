"""Multi-GPU plumbing for sweeps: frames (or whole cases) are independent, so ranks get contiguous frame blocks and
the only exchange is a gather of per-frame integer scores.  No collective sits on the forward
path; ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests) is used for the final few-KB gather only.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(n_frames: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced block of frames for ``rank`` (first ``n % world`` ranks get one extra frame)."""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_areas(local_areas: np.ndarray, n_frames: int, device: torch.device | str = "cpu") -> np.ndarray:
    """All-gather the per-rank ``int32`` area blocks into the full ``[n_frames]`` vector (rank order == frame order).
    Falls through to the local vector when no process group is initialised."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return np.asarray(local_areas, np.int32)
    world = dist.get_world_size()
    longest = max(shard_range(n_frames, world, r)[1] - shard_range(n_frames, world, r)[0] for r in range(world))
    buf = torch.zeros(longest, dtype=torch.int32, device=device)
    buf[: len(local_areas)] = torch.as_tensor(np.asarray(local_areas, np.int32), device=device)
    parts: List[torch.Tensor] = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = []
    for r, p in enumerate(parts):
        lo, hi = shard_range(n_frames, world, r)
        out.append(p[: hi - lo].cpu().numpy())
    return np.concatenate(out).astype(np.int32)


def select_global(areas: np.ndarray) -> Tuple[int, int]:
    """``(best_idx, owner-independent area)`` with numpy's first-max tie-break; ``(-1, 0)`` for an empty sweep."""
    if areas.size == 0 or int(areas.max()) == 0:
        return -1, 0
    i = int(areas.argmax())
    return i, int(areas[i])


def owner_of(frame: int, n_frames: int, world: int) -> int:
    for r in range(world):
        lo, hi = shard_range(n_frames, world, r)
        if lo <= frame < hi:
            return r
    raise ValueError("frame out of range")
