"""Profiling driver: xb200_analyze_intra over the 32/16/8/4 quad-tree of one 1080p picture (run under ncu).
usage: python tools/prof_intra.py [reps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from xeve_b200 import api  # noqa: E402
from xeve_b200.clips import to_internal10  # noqa: E402
from xeve_b200.worklist import synth_intra  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
hp = api.Hotpath(api.make_seq(bench.W, bench.H, bench.PRESET))
clip, fr = bench.frames_for_bench()
cur = hp.pic_create(padded=False)
hp.pic_upload(cur, *fr[bench.POC], clip.depth)
items, states, rates, side, elems = synth_intra(bench.W, bench.H, [to_internal10(p, clip.depth) for p in fr[bench.POC]], cur, hp.rdoq_rates)
for _ in range(reps):
    out, st, coef, rec = hp.analyze_intra(items, rates, states, side, elems)
    print("analyze_intra kernel ms", round(hp.last_kernel_ms, 3), "CUs", len(items), "modes", np.bincount(out["ipm"][:, 0], minlength=5).tolist(),
          "nnz CUs", int((out["nnz"] != 0).any(1).sum()))
hp.close()
