"""Profiling driver: xb200_deblock of one synthetic 1080p (or WxH) picture (run under ncu).
usage: python tools/prof_deblock.py [reps] [w h]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xeve_b200 import api  # noqa: E402
from xeve_b200.worklist import synth_deblock  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
hp = api.Hotpath(api.make_seq(w, h))
d = synth_deblock(w, h, 0)
pic = hp.pic_create(padded=True)
for _ in range(reps):
    hp.pic_upload_s16(pic, *(np.ascontiguousarray(a) for a in d["pre"]))
    hp.deblock(pic, d["cus"], d["pp"], d["map_scu"], d["map_refi"], d["map_mv"])
    print("deblock kernel ms", round(hp.last_kernel_ms, 4), "CUs", len(d["cus"]))
hp.close()
