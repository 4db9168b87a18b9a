"""Host-side logic of the range-sharded MSM: which points a rank owns and how the
per-rank window sums are exchanged and combined.  The reference's analogue is the 8-way contiguous
range partition of /root/reference/porla/Client/Client.hpp:747-787 (thread t owns
[t*n/8, (t+1)*n/8), partial sums added serially); here the partition is over GPUs/processes and
the exchange is one all-gather of nwin*128 bytes per rank (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of rank `rank`; the last rank takes the remainder, exactly as
    Client.hpp:753-754 gives the last thread `n_points - n_point_each_thread*t`."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    each = n // world
    lo = each * rank
    hi = n if rank == world - 1 else lo + each
    return lo, hi


def gather_window_sums(local, world: int, dist=None):
    """All-gather the per-rank window-sum buffer (a torch.uint8 tensor of nwin*128 bytes).
    Returns a tensor of world*nwin*128 bytes ordered by rank.  `dist` defaults to
    torch.distributed; with world == 1 nothing is exchanged."""
    if world == 1:
        return local
    import torch
    if dist is None:
        import torch.distributed as dist  # type: ignore
    out = torch.empty(local.numel() * world, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local)
    return out
