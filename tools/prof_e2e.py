"""Wall time of every host-buffer C-ABI call of one bench step (where does e2e go?).  usage: python tools/prof_e2e.py [reps]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from xeve_b200 import api  # noqa: E402
from xeve_b200.worklist import FrameWork  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
seq = api.make_seq(bench.W, bench.H, bench.PRESET)
hp = api.Hotpath(seq)
L, ctx = hp.L, hp.h
clip, fr = bench.frames_for_bench()
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
refs = []
for poc in bench.REF_POCS:
    hd = hp.pic_create(padded=True)
    hp.pic_upload(hd, *fr[poc], clip.depth)
    refs.append(hd)
cur = hp.pic_create(padded=False)
cur_planes = [pin(p) for p in fr[bench.POC]]
hp.pic_upload(cur, *[p.numpy() for p in cur_planes], clip.depth)
fw = FrameWork(bench.W, bench.H, bench.POC, bench.REF_POCS, bench.PAN, cur, refs, me_range=int(seq["me_range"][0]))
cnt = fw.counts()
me_uni = hp.me(fw.me_uni)
bi_mc, me_bi_in = fw.build_bi(me_uni)
side = hp.bi_org(bi_mc, fw.bi_cur, fw.side_off, fw.side_elems)
me_bi = hp.me(me_bi_in, side)
res_in = fw.build_residue(me_bi)
h_me_uni, h_me_bi, h_res = pin(fw.me_uni.view(np.uint8)), pin(me_bi_in.view(np.uint8)), pin(res_in.view(np.uint8))
h_side, h_coef = pin(np.zeros(fw.side_elems, np.int16)), pin(np.zeros(fw.res_elems, np.int16))
HP = lambda t: C.c_void_p(t.data_ptr())
planes = (C.c_void_p * 3)(*[p.data_ptr() for p in cur_planes])
strides = (C.c_int32 * 3)(bench.W, bench.W // 2, bench.W // 2)
calls = [
    ("pic_upload", lambda: L.xb200_pic_upload(ctx, cur, planes, strides, clip.depth, api.MEM_HOST), 3 * bench.W * bench.H // 2, 0),
    ("me_uni", lambda: L.xb200_me(ctx, HP(h_me_uni), cnt["me_uni"], None, 0, api.MEM_HOST), fw.me_uni.nbytes, fw.me_uni.nbytes),
    ("bi_org", lambda: L.xb200_bi_org(ctx, bi_mc.ctypes.data_as(C.c_void_p), cnt["bi_org"], fw.bi_cur.ctypes.data_as(C.c_void_p),
                                      fw.side_off.ctypes.data_as(C.c_void_p), HP(h_side), fw.side_elems, api.MEM_HOST),
     bi_mc.nbytes + fw.bi_cur.nbytes + fw.side_off.nbytes, 2 * fw.side_elems),
    ("me_bi", lambda: L.xb200_me(ctx, HP(h_me_bi), cnt["me_bi"], HP(h_side), fw.side_elems, api.MEM_HOST), me_bi_in.nbytes + 2 * fw.side_elems,
     me_bi_in.nbytes),
    ("residue", lambda: L.xb200_residue(ctx, HP(h_res), cnt["residue"], fw.rates.ctypes.data_as(C.c_void_p), 1, HP(h_coef), None, fw.res_elems,
                                        api.MEM_HOST), res_in.nbytes, res_in.nbytes + 2 * fw.res_elems),
]
tot = {k: 0.0 for k, *_ in calls}
kms = {k: 0.0 for k, *_ in calls}
for it in range(reps + 1):
    for name, fn, h2d, d2h in calls:
        t0 = time.perf_counter()
        assert fn() == 0
        if it:
            tot[name] += time.perf_counter() - t0
            kms[name] += hp.last_kernel_ms
for name, fn, h2d, d2h in calls:
    ms = tot[name] / reps * 1e3
    print(f"{name:10s} {ms:7.3f} ms wall  kernels {kms[name] / reps:6.3f} ms  h2d {h2d / 1e6:6.1f} MB  d2h {d2h / 1e6:6.1f} MB  "
          f"-> {(h2d + d2h) / 1e6 / max(ms - kms[name] / reps, 1e-3):6.1f} GB/s for the copies")
print("step", round(sum(tot.values()) / reps * 1e3, 3), "ms")
ro = np.frombuffer(h_res.numpy().tobytes(), api.RESIDUE_ITEM)
wsq = ro["mc"]["w"].astype(np.int64) ** 2
sizes = np.stack([wsq, wsq // 4, wsq // 4], 1)
print("non-zero planes", round(float((ro["nnz"] != 0).mean()), 4), "of the planes,", round(float((sizes * (ro["nnz"] != 0)).sum() / sizes.sum()), 4),
      "of the coefficient bytes =", round(2 * float((sizes * (ro["nnz"] != 0)).sum()) / 1e6, 2), "MB of", round(2 * fw.res_elems / 1e6, 1))
hp.close()
