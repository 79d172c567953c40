"""CPU test of the N > 1 host path (gloo, world_size 2): header broadcast, max-over-ranks timing, sharding."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from xeve_b200 import api
    from xeve_b200 import dist as xd
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seq = api.make_seq(1920, 1080, "fast") if rank == 0 else api.make_seq(352, 288, "medium")
    got = xd.broadcast_seq(seq, dist)
    t = xd.max_over_ranks([1.0 + rank, 5.0 - rank], dist)
    q.put((rank, got.tobytes(), t, xd.shard_frames(8, rank, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_header_broadcast_and_timing_reduce_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    sys.path.insert(0, ROOT)
    from xeve_b200 import api
    ref = api.make_seq(1920, 1080, "fast").tobytes()
    assert res[0][1] == ref and res[1][1] == ref           # every rank holds rank 0's header
    assert res[0][2] == res[1][2] == [2.0, 5.0]            # max over ranks
    assert res[0][3] == [0, 2, 4, 6] and res[1][3] == [1, 3, 5, 7]
