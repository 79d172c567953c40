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


# ---- within one stream: pictures of one wave of the picture DAG on different ranks, reference pictures broadcast ------------------------
def _refs_of(pc):
    pp = pc["pp"]
    return sorted({int(pp["ref_poc"][l][k]) for l in range(2) for k in range(4) if int(pp["ref_pic"][l][k]) >= 0})


def _dag_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import tracedata
    from xeve_b200 import dist as xd
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seq, pics = tracedata.live_chain(frames=17)                       # every rank traces the same deterministic reference encode
    by_poc = {int(pc["pp"]["poc"]): pc for pc in pics}
    waves, owner, exchanged = xd.picture_plan([(int(pc["pp"]["poc"]), _refs_of(pc)) for pc in pics], world)
    enc = tracedata.ChainEncoder(seq, check=True)                     # asserts every picture it encodes against the reference's
    mine, received = [], []
    for wave in waves:
        for poc in wave:
            if owner[poc] == rank:
                enc.encode(by_poc[poc])
                mine.append(poc)
        for poc in wave:                                              # exchange step of the wave
            if poc not in exchanged:
                continue
            shape = [a.shape for a in by_poc[poc]["org"]]
            f = ((shape[0][1] + 3) // 4) * ((shape[0][0] + 3) // 4)
            if owner[poc] == rank:
                planes, mv = enc.done[poc]["post"], enc.done[poc]["map_mv"]
            else:
                planes, mv = [np.zeros(s, np.int16) for s in shape], np.zeros((f, 2, 2), np.int16)
            planes, mv = xd.broadcast_picture(planes, mv.reshape(f, 2, 2), owner[poc], dist)
            if owner[poc] != rank:
                enc.adopt(poc, planes, mv)
                received.append(poc)
    q.put((rank, mine, received, waves, sorted(exchanged)))
    dist.barrier()
    dist.destroy_process_group()


def test_picture_dag_sharded_over_two_ranks_matches_the_reference():
    """SURVEY 8e within one stream: the 17 pictures of a hierarchical-B GOP split over 2 ranks wave by wave (xd.picture_plan), each
    reference picture broadcast once after its wave (xd.broadcast_picture: deblocked planes + MV map); every rank's pictures --
    decided from references it partly received from the other rank -- equal the single-thread reference's (asserted per picture
    inside ChainEncoder).  The per-picture work is the oracle chain here (CPU); on the GPU box the same plan drives one device each."""
    import pytest
    sys.path.insert(0, ROOT)
    from oracle import refharness as rh
    if not rh.available():
        pytest.skip("oracle/_ref (compiled reference) not built here")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dag_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    import queue
    res = []
    for _ in range(600):
        try:
            res.append(q.get(timeout=0.5))
        except queue.Empty:
            assert all(p.exitcode in (None, 0) for p in procs), "a rank died"
        if len(res) == len(procs):
            break
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, mine0, recv0, waves, exch), (_, mine1, recv1, _, _) = sorted(res)
    assert waves == [[0], [16], [8], [4, 12], [2, 6, 10, 14], [1, 3, 5, 7, 9, 11, 13, 15]]
    assert sorted(mine0 + mine1) == list(range(17)) and not set(mine0) & set(mine1)
    assert len(mine1) == 7 and set(recv0) == set(mine1) & set(exch) and set(recv1) == set(mine0) & set(exch)
    assert not set(exch) & {1, 3, 5, 7, 9, 11, 13, 15}               # the deepest layer is never a reference: nothing to send


def test_picture_plan_host_logic():
    sys.path.insert(0, ROOT)
    from xeve_b200 import dist as xd
    gop = [(0, []), (16, [0]), (8, [0, 16]), (4, [0, 8]), (12, [8, 16]), (2, [0, 4]), (6, [4, 8]), (10, [8, 12]), (14, [12, 16]),
           (1, [0, 2]), (3, [2, 4])]
    waves, owner, exch = xd.picture_plan(gop, 4)
    assert waves == [[0], [16], [8], [4, 12], [2, 6, 10, 14], [1, 3]]
    assert [owner[p] for p in (2, 6, 10, 14)] == [0, 1, 2, 3] and owner[0] == owner[16] == owner[8] == 0
    assert exch == {2, 4, 8, 12, 16}                                  # POC 0 is only read on its owner's rank; POC 2 is read by POC 3 on rank 1
