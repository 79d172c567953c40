"""Generates tests/golden/df_golden.npz from the compiled reference (oracle/_ref, this container).

Run:  python tests/golden/make_df_golden.py
Deblocking fixture: the first three pictures in coding order (POC 0 = intra picture with 4x4 CUs, POC 16,
POC 8) of a Baseline/fast encode of the seeded 176x144 clip -- the unfiltered reconstruction ctx->fn_loop_filter receives,
the frame maps it reads (map_scu / map_refi / map_mv), the leaf CUs xeve_deblock_tree enumerates, and the picture it leaves
behind (the reference's in-situ result).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tracedata  # noqa: E402

pics = tracedata.live_df(pic_hi=2)
out = dict(n=np.int32(len(pics)))
for i, d in enumerate(pics):
    for k in ("cus", "pp", "map_scu", "map_refi", "map_mv"):
        out[f"{k}{i}"] = d[k]
    for k, a in zip("yuv", d["pre"]):
        out[f"pre{i}_{k}"] = a
    for k, a in zip("yuv", d["post"]):
        out[f"post{i}_{k}"] = a
path = os.path.join(ROOT, "tests", "golden", "df_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB;", [len(d["cus"]) for d in pics], "CUs")
