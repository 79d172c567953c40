"""Generates tests/golden/intra_golden.npz from the compiled reference (oracle/_ref, this container).

Run:  python tests/golden/make_intra_golden.py
Intra-analysis fixture: 400 ctx->fn_pintra_analyze_cu calls sampled from the first three pictures in coding
order (the intra picture POC 0 plus two B pictures) of a Baseline/fast encode of the seeded 176x144 clip -- every CU size that
occurs (4x4 .. 64x64), with the reference samples xeve_get_nbr assembled, the coder states, the RDOQ rate tables and the
reference's in-situ results (cost, modes, nnz, coefficient / reconstruction hashes, output coder state).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tracedata  # noqa: E402

td = tracedata.live_intra(pic_hi=2)
tr = td.live
rng = np.random.default_rng(11)
allit = td.intra
big = np.nonzero(allit["log2_cuw"] >= 4)[0]                       # keep every large CU, sample the small ones
small = rng.permutation(np.nonzero(allit["log2_cuw"] < 4)[0])[:400 - min(len(big), 150)]
idx = np.sort(np.concatenate([big[:150], small]))
it = allit[idx].copy()
used_rates = np.unique(it["rate_idx"])
rmap = {int(r): i for i, r in enumerate(used_rates)}
used_pics = np.unique(it["cur_pic"])
pmap = {int(p): i for i, p in enumerate(used_pics)}
sbac = np.zeros(2 * len(it), tr.cu_sbac.dtype)
side, pos = [], 0
for i, r in enumerate(it):
    sbac[2 * i], sbac[2 * i + 1] = tr.cu_sbac[int(r["state_in"])], tr.cu_sbac[int(r["state_out"])]
    n = 8 * (1 << int(r["log2_cuw"])) + 6
    side.append(tr.samp[int(r["nb_off"]):int(r["nb_off"]) + n])
    r["state_in"], r["state_out"], r["rate_idx"], r["cur_pic"], r["nb_off"] = 2 * i, 2 * i + 1, rmap[int(r["rate_idx"])], pmap[int(r["cur_pic"])], pos
    pos += n
out = dict(intra=it, sbac=sbac, rates=tr.rates[used_rates], side=np.concatenate(side), seq=tr.const, pics=tr.pics[used_pics])
for i, p in enumerate(used_pics):
    out[f"p{i}_y"], out[f"p{i}_u"], out[f"p{i}_v"] = td.planes[int(p)]
path = os.path.join(ROOT, "tests", "golden", "intra_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(it), "CUs, sizes", np.bincount(it["log2_cuw"]).tolist(), "slice types",
      np.bincount(it["slice_type"]).tolist())
