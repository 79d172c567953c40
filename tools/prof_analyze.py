"""Profiling driver: one 1080p picture through xb200_analyze_cu (run under ncu / compute-sanitizer).
usage: python tools/prof_analyze.py [reps] [rows]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from xeve_b200 import api  # noqa: E402
from xeve_b200.worklist import FrameWork  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rows = int(sys.argv[2]) if len(sys.argv) > 2 else None
seq = api.make_seq(bench.W, bench.H, bench.PRESET)
hp = api.Hotpath(seq)
clip, fr = bench.frames_for_bench()
refs = []
for poc in bench.REF_POCS:
    hd = hp.pic_create(padded=True)
    hp.pic_upload(hd, *fr[poc], clip.depth)
    refs.append(hd)
cur = hp.pic_create(padded=False)
hp.pic_upload(cur, *fr[bench.POC], clip.depth)
fw = FrameWork(bench.W, bench.H, bench.POC, bench.REF_POCS, bench.PAN, cur, refs, me_range=int(seq["me_range"][0]), rows=rows)
cu = fw.build_cu(hp.rdoq_rates, max_search_range=int(seq["me_range"][0]))
for _ in range(reps):
    out, st, coef, rec = hp.analyze_cu(cu, fw.cu_rates, fw.cu_states, fw.cu_elems)
    print("analyze_cu kernel ms", round(hp.last_kernel_ms, 3), "CUs", len(cu), "modes", np.bincount(out["best_idx"], minlength=5).tolist())
hp.close()
