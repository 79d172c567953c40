"""Multi-GPU plumbing of the hot path: replicas only.

Independent picture streams shard across GPUs, one process per GPU; the only exchange is the
"trivial broadcast of headers": rank 0 sends the sequence constants (xb200_seq) so that every rank
configures its context identically, and timings are reduced with MAX for reporting.  Works with any
torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from . import api


def broadcast_seq(seq: np.ndarray, dist, device="cpu") -> np.ndarray:
    """Rank 0's xb200_seq blob -> every rank (byte-exact)."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(seq).view(np.uint8).copy()).to(device)
    dist.broadcast(t, src=0)
    return t.cpu().numpy().view(api.SEQ).copy()


def max_over_ranks(values, dist, device="cpu"):
    """Element-wise MAX of a list of floats over all ranks (device timings are reported as max-over-ranks)."""
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def shard_frames(n_frames: int, rank: int, world: int):
    """Round-robin assignment of the pictures of one temporal layer (or of independent streams) to ranks."""
    return list(range(rank, n_frames, world))
