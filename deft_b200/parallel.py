"""Stream-parallel sharding of the hot path: video streams (or nuScenes cameras) are
independent recurrences, so they are dealt round-robin to ranks -- one process per GPU, weights replicated, NO
collective on the data path.  torch.distributed is only used to agree on the clock: a barrier on both sides of the
timed region and max-over-ranks of the device time."""
from __future__ import annotations

import torch
import torch.distributed as dist


def stream_assignment(n_streams: int, world_size: int, rank: int):
    """Round-robin stream ids owned by `rank` (C4: 6 cameras over 2/4/8 GPUs -> 3 / 2,2,1,1 / 1x6 + 2 idle)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    return list(range(rank, n_streams, world_size))


def aggregate_throughput(local_units: float, local_ms: float, device=None):
    """Whole-job throughput = sum of units over ranks / max of time over ranks (units per second).
    Works on any initialised backend (nccl with a CUDA device, gloo on CPU); single process if uninitialised."""
    if not (dist.is_available() and dist.is_initialized()):
        return local_units / (local_ms * 1e-3), local_ms
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    t = torch.tensor([local_ms], dtype=torch.float64, device=dev)
    u = torch.tensor([local_units], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(u.item()) / (float(t.item()) * 1e-3), float(t.item())
