"""How many host cores does the reference arm really get?  Replays the ME work list of 2 CTU rows on 1..N threads."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import refharness as rh  # noqa: E402
from xeve_b200 import api  # noqa: E402
from xeve_b200.worklist import FrameWork  # noqa: E402

print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    if os.path.exists(f):
        print(f, open(f).read().strip())
clip, fr = bench.frames_for_bench()
seq = api.make_seq(bench.W, bench.H, bench.PRESET)
keep, planes = bench._padded_planes_struct(fr, [bench.REF_POCS[0], bench.REF_POCS[1], bench.POC], clip.depth)
fw = FrameWork(bench.W, bench.H, bench.POC, bench.REF_POCS, bench.PAN, 2, [0, 1], rows=2, me_range=32)
for th in (1, 2, 4, 8, 16, 32, 64, 128):
    t0 = time.perf_counter()
    _, s = rh.replay_me_raw(seq, planes, None, fw.me_uni.astype(rh.ME_REC), th)
    print(f"threads {th:3d}: harness {s:.3f} s  wall {time.perf_counter() - t0:.3f} s  items/s {len(fw.me_uni) / s:.0f}")
