"""Generates tests/golden/qcif_trace.npz from the compiled reference (oracle/_ref, this container).

Run:  python tests/golden/make_golden.py
The fixture holds work lists traced from a real Baseline/fast encode of a seeded 176x144 clip
(pictures 1..2 in coding order: POC 16 and POC 8) with the reference's in-situ results, so the
oracle and the CUDA path can be pinned on machines that have neither /root/reference nor
oracle/_ref.  Everything is bounded to keep the file small.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tracedata  # noqa: E402
from oracle import refharness as rh  # noqa: E402

td = tracedata.live_trace("cif", frames=20, pic_lo=1, pic_hi=2, mask=15, **tracedata.QCIF)
tr = td.live
rng = np.random.default_rng(5)

# ME: everything from the two pictures, org_bi blocks compacted into a side buffer
me = tr.me.copy()
side = []
pos = 0
for r in me:
    if r["bi"]:
        n = 1 << (int(r["log2_cuw"]) + int(r["log2_cuh"]))
        side.append(tr.samp[int(r["org_bi_off"]):int(r["org_bi_off"]) + n])
        r["org_bi_off"] = pos
        pos += n
side = np.concatenate(side) if side else np.zeros(0, np.int16)

# MC: first 1500 calls with expected predictions
mc = tr.mc[:1500].copy()
pred, off, hsh, _ = rh.replay_mc(tr, recs=mc, nthreads=4)
assert np.array_equal(hsh, mc["out_hash"])

# TQ: 700 calls, inputs compacted, expected outputs from the reference
idx = np.sort(rng.permutation(len(tr.tq))[:700])
tq = tr.tq[idx].copy()
coef_ref, nnz_ref, resi_ref, _ = rh.replay_tq(tr, nthreads=4)
tq_in, tq_out, tq_resi = [], [], []
pos = 0
used_rates = np.unique(tq["rate_idx"])
remap = {int(r): i for i, r in enumerate(used_rates)}
for r in tq:
    n = (3 << (2 * int(r["log2_cuw"]))) >> 1
    o = int(r["in_off"])
    tq_in.append(tr.samp[o:o + n]); tq_out.append(coef_ref[o:o + n]); tq_resi.append(resi_ref[o:o + n])
    r["in_off"] = pos
    r["rate_idx"] = remap[int(r["rate_idx"])]
    pos += n
assert np.array_equal(nnz_ref[idx], tq["nnz"])

# CU decisions: 500 xeve_pinter_analyze_cu calls with the reference's results, coder states and rate tables compacted
cidx = np.sort(rng.permutation(len(tr.cu))[:500])
cu = tr.cu[cidx].copy()
cu_rates_used = np.unique(cu["rate_idx"])
rmap = {int(r): i for i, r in enumerate(cu_rates_used)}
cu_sbac = np.zeros(2 * len(cu), tr.cu_sbac.dtype)
for i, r in enumerate(cu):
    cu_sbac[2 * i], cu_sbac[2 * i + 1] = tr.cu_sbac[int(r["state_in"])], tr.cu_sbac[int(r["state_out"])]
    r["state_in"], r["state_out"], r["rate_idx"] = 2 * i, 2 * i + 1, rmap[int(r["rate_idx"])]
    r["me_first"] = r["me_cnt"] = 0

out = dict(cu=cu, cu_sbac=cu_sbac, cu_rates=tr.rates[cu_rates_used], seq=tr.const, pics=tr.pics, me=me, side=side, mc=mc, mc_pred=pred, mc_off=off, tq=tq, rates=tr.rates[used_rates],
           tq_in=np.concatenate(tq_in), tq_coef_out=np.concatenate(tq_out), tq_resi_out=np.concatenate(tq_resi), tq_nnz=nnz_ref[idx])
for i in range(len(tr.pics)):
    out[f"p{i}_y"], out[f"p{i}_u"], out[f"p{i}_v"] = td.planes[i]
path = os.path.join(ROOT, "tests", "golden", "qcif_trace.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path) // 1024, "KiB;", len(me), "ME,", len(mc), "MC,", len(tq), "TQ,", len(cu), "CU items")
