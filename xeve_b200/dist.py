"""Multi-GPU plumbing of the hot path.

Across streams: replicas only.  Independent picture streams shard across GPUs, one process per GPU; the only exchange is the
"trivial broadcast of headers": rank 0 sends the sequence constants (xb200_seq) so that every rank
configures its context identically, and timings are reduced with MAX for reporting.
Within one stream: the pictures of one wave of the picture DAG go to different ranks and each reference picture is broadcast once
after its wave (picture_plan / broadcast_picture below).
Works with any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).
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


# ---- within one stream: the picture DAG ("pictures of the same temporal layer are independent") -----------------------
def picture_waves(pictures):
    """pictures: [(poc, [POCs it references]), ...] in coding order.  Returns the waves of the picture DAG: wave k holds every picture
    whose references all lie in waves < k (for the default hierarchical-B GOP of 16: [0], [16], [8], [4, 12], [2, 6, 10, 14],
    [1, 3, .., 15]).  Pictures of one wave share no data, so they can be decided on different GPUs; order inside a wave is coding
    order."""
    level, waves = {}, []
    for poc, refs in pictures:
        lv = 1 + max((level[r] for r in refs), default=-1)
        level[poc] = lv
        while len(waves) <= lv:
            waves.append([])
        waves[lv].append(poc)
    return waves


def picture_plan(pictures, world):
    """(waves, owner, exchanged): owner[poc] = rank that decides the picture (round-robin inside its wave); exchanged = the POCs some
    other rank than the owner reads later, i.e. the reference pictures that have to travel (one broadcast each: the deblocked, padded
    picture and its MV map, SURVEY 8e: 9.1 + 1.3 MB at 1080p)."""
    waves = picture_waves(pictures)
    owner = {poc: i % world for wave in waves for i, poc in enumerate(wave)}
    exchanged = {r for poc, refs in pictures for r in refs if world > 1 and any(owner[p] != owner[r] for p, rf in pictures if r in rf)}
    return waves, owner, exchanged


def broadcast_picture(planes, map_mv, src, dist, device="cpu"):
    """One exchange step: rank `src` sends a reconstructed reference picture (Y, U, V s16 planes) and its MV map; every rank returns the
    same arrays.  On the receiving ranks `planes` / `map_mv` only give the shapes."""
    import torch
    out = []
    for a in list(planes) + [map_mv]:
        a = np.ascontiguousarray(a, np.int16)
        t = torch.from_numpy(a.view(np.uint8).copy()).to(device)     # bytes: gloo has no 16-bit integer type
        dist.broadcast(t, src=src)
        out.append(t.cpu().numpy().view(np.int16).reshape(a.shape))
    return out[:3], out[3]
