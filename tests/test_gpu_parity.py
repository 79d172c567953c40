"""GPU parity tests: the CUDA path, called through the C ABI, against (a) the results the
UNMODIFIED reference produced in situ for the very same calls (traced work lists) and (b) the
oracle restatement.  Bit-exact: everything here is integer work."""
import numpy as np
import pytest

import tracedata
from oracle import oracle as xo
from oracle import refharness as rh
from xeve_b200 import api

pytestmark = pytest.mark.gpu


def _remap(arr, field, handles):
    out = arr.copy()
    v = out[field]
    out[field] = np.where(v >= 0, handles[np.clip(v, 0, len(handles) - 1)], -1)
    return out


def test_padding_matches_reference(trace, gpu_ctx):
    """device border replication == xeve_picbuf_expand (reference src_base/xeve_util.c:190-248)"""
    for i, p in enumerate(trace.pics):
        if int(p["kind"]) != 1:
            continue
        got = gpu_ctx.pic_download(int(gpu_ctx.handles[i]), with_padding=True)
        exp = trace.padded_planes(i)
        for g, e in zip(got, exp):
            assert np.array_equal(g, e)
        if trace.source == "live":  # the reference's own padded buffer
            full = trace.live.plane_views(i)
            assert np.array_equal(got[0], full[0][:, : got[0].shape[1]])
            assert np.array_equal(got[1], full[1][:, : got[1].shape[1]])


def test_upload_8bit_converts_and_pads(gpu_ctx):
    """xb200_pic_upload: 8-bit input -> internal 10-bit (<< 2) + padding"""
    from xeve_b200.clips import Clip, to_internal10
    w, h = gpu_ctx.w, gpu_ctx.hgt
    c = Clip("cif", w=w, h=h)
    y, u, v = c.frame(3)
    hd = gpu_ctx.pic_create(padded=True)
    gpu_ctx.pic_upload(hd, y, u, v, 8)
    gy, gu, gv = gpu_ctx.pic_download(hd, with_padding=False)
    assert np.array_equal(gy, to_internal10(y, 8)) and np.array_equal(gu, to_internal10(u, 8)) and np.array_equal(gv, to_internal10(v, 8))
    py = gpu_ctx.pic_download(hd, with_padding=True)[0]
    assert np.array_equal(py[144:-144, 144:-144], gy)
    assert (py[:144, :144] == gy[0, 0]).all() and (py[-144:, -144:] == gy[-1, -1]).all()
    assert np.array_equal(py[10, 144:-144], gy[0]) and np.array_equal(py[144:-144, 3], gy[:, 0])
    gpu_ctx.pic_destroy(hd)


def test_distortion_probes(trace, gpu_ctx):
    """SAD / SSD / SATD kernels vs the oracle on random blocks reaching into the padding"""
    import ctypes as C
    rng = np.random.default_rng(7)
    refs = [i for i, p in enumerate(trace.pics) if int(p["kind"]) == 1]
    orgs = [i for i, p in enumerate(trace.pics) if int(p["kind"]) == 0]
    n = 600
    items = np.zeros(n, api.BLK_ITEM)
    w, h = gpu_ctx.w, gpu_ctx.hgt
    exp_sad, exp_ssd, exp_satd = np.zeros(n, np.int32), np.zeros(n, np.int64), np.zeros(n, np.int32)
    L = xo.lib()
    for k in range(n):
        l2 = int(rng.integers(2, 7))
        pl = int(rng.integers(0, 3))
        sc = 1 if pl == 0 else 2
        bs = 1 << l2
        o, r = int(rng.choice(orgs)), int(rng.choice(refs))
        x1 = int(rng.integers(0, (w // sc - bs) // 4 + 1)) * 4
        y1 = int(rng.integers(0, (h // sc - bs) // 4 + 1)) * 4
        x2 = int(rng.integers(-127 // sc, w // sc - 1))
        y2 = int(rng.integers(-127 // sc, h // sc - 1))
        items[k] = (gpu_ctx.handles[o], gpu_ctx.handles[r], x1, y1, x2, y2, pl, pl, l2, l2)
        a = trace.planes[o][pl]
        b = trace.padded_planes(r)[pl]
        pad = 144 // sc
        ap = a.ctypes.data + 2 * (y1 * a.shape[1] + x1)
        bp = b.ctypes.data + 2 * ((y2 + pad) * b.shape[1] + x2 + pad)
        exp_sad[k] = L.xo_sad(bs, bs, ap, a.shape[1], bp, b.shape[1], 10)
        exp_ssd[k] = L.xo_ssd(bs, bs, ap, a.shape[1], bp, b.shape[1], 10)
        exp_satd[k] = L.xo_satd(bs, bs, ap, a.shape[1], bp, b.shape[1], 10)
    assert np.array_equal(gpu_ctx.sad(items), exp_sad)
    assert np.array_equal(gpu_ctx.ssd(items), exp_ssd)
    assert np.array_equal(gpu_ctx.satd(items), exp_satd)


def _me_items(trace, gpu_ctx):
    it = np.ascontiguousarray(trace.me).astype(api.ME_ITEM)
    it = _remap(it, "cur_pic", gpu_ctx.handles)
    it = _remap(it, "ref_pic", gpu_ctx.handles)
    return it


def test_me_matches_reference_in_situ(trace, gpu_ctx):
    """xb200_me vs what pinter_me_epzs returned inside the real encode (mv, cost, mot_bits)"""
    items = _me_items(trace, gpu_ctx)
    items["mv_out"] = 0
    items["cost"] = 0
    items["mot_bits_out"] = 0
    got = gpu_ctx.me(items, trace.side)
    exp = trace.me
    bad = np.where((got["mv_out"] != exp["mv_out"]).any(1) | (got["cost"] != exp["cost"]) |
                   (got["mot_bits_out"] != exp["mot_bits_out"]).any(1))[0]
    assert len(bad) == 0, (len(bad), len(exp), exp[bad[:3]], got[bad[:3]])
    assert len(exp) > 1000 and (exp["bi"] == 1).sum() > 100 and set(np.unique(exp["log2_cuw"])) == {3, 4, 5, 6}


def test_me_matches_oracle(trace, gpu_ctx):
    items = _me_items(trace, gpu_ctx)
    got = gpu_ctx.me(items, trace.side)
    opl = trace.oracle_planes()
    exp = xo.me_batch(trace.seq, opl, trace.side, np.ascontiguousarray(trace.me))
    for f in ("mv_out", "cost", "mot_bits_out"):
        assert np.array_equal(got[f], exp[f]), f


def test_mc_matches_reference(trace, gpu_ctx):
    """xb200_mc vs xeve_mc (replayed through the reference) and the oracle: Y, U, V, uni + bi"""
    recs = np.ascontiguousarray(trace.mc)
    off, total = rh.mc_offsets(recs)
    items = _remap(recs.astype(api.MC_ITEM), "ref_pic", gpu_ctx.handles)
    got = gpu_ctx.mc(items, off, total)
    exp = xo.mc_batch(trace.seq, trace.oracle_planes(), recs, off, total)
    assert np.array_equal(got, exp)
    if trace.source == "live":
        pred_ref, _, hsh, _ = rh.replay_mc(trace.live, nthreads=4)
        assert np.array_equal(hsh, recs["out_hash"])  # replay == in situ
        assert np.array_equal(got, pred_ref)
    else:
        n = len(trace.expect["mc_off"])
        end = int(trace.expect["mc_off"][-1] + recs[n - 1]["w"] * recs[n - 1]["h"] * 3 // 2)
        assert np.array_equal(got[:end], trace.expect["mc_pred"][:end])
    assert ((recs["refi"] >= 0).all(1)).sum() > 100  # bi-prediction is exercised


def _tq_expected(trace):
    if trace.source == "live":
        coef, nnz, resi, _ = rh.replay_tq(trace.live, nthreads=4)
        return coef, nnz, resi
    items, coef, resi = xo.tq_batch(trace.seq, np.ascontiguousarray(trace.tq), trace.rates, trace.tq_coef)
    return coef, items["nnz"], resi


def _item_mask(tq, n):
    m = np.zeros(n, bool)
    for r in tq:
        o = int(r["in_off"])
        m[o:o + ((3 << (2 * int(r["log2_cuw"]))) >> 1)] = True
    return m


def test_tq_itdq_recon_match_reference(trace, gpu_ctx):
    """xb200_tq (DCT + RDOQ), xb200_itdq, xb200_recon vs xeve_sub_block_tq / xeve_itdq / xeve_recon_blk"""
    tq = np.ascontiguousarray(trace.tq).astype(api.TQ_ITEM)
    exp_coef, exp_nnz, exp_resi = _tq_expected(trace)
    if trace.source == "live":
        assert np.array_equal(exp_nnz, trace.tq["nnz"])  # replay == in situ
    tq["nnz"] = 0
    items, coef = gpu_ctx.tq(tq, trace.rates, trace.tq_coef)
    m = _item_mask(tq, len(coef))
    assert np.array_equal(items["nnz"], exp_nnz)
    assert np.array_equal(coef[m], exp_coef[m])
    resi = gpu_ctx.itdq(items, coef)
    # planes with nnz == 0 are left untouched by fn_itdp: compare only where the reference wrote
    for r, nz in zip(items, exp_nnz):
        o, ny = int(r["in_off"]), 1 << (2 * int(r["log2_cuw"]))
        for c, (a, b) in enumerate(((0, ny), (ny, ny + ny // 4), (ny + ny // 4, ny + ny // 2))):
            if nz[c]:
                assert np.array_equal(resi[o + a:o + b], exp_resi[o + a:o + b])
    # recon against the oracle on a synthetic prediction
    rng = np.random.default_rng(3)
    pred = rng.integers(0, 1024, len(coef)).astype(np.int16)
    rec = gpu_ctx.recon(items, resi, pred)
    L = xo.lib()
    import ctypes as C
    for r in items[:: max(1, len(items) // 300)]:
        o, ny = int(r["in_off"]), 1 << (2 * int(r["log2_cuw"]))
        for c, (a, b) in enumerate(((0, ny), (ny, ny + ny // 4), (ny + ny // 4, ny + ny // 2))):
            e = np.zeros(b - a, np.int16)
            rs, pr = np.ascontiguousarray(resi[o + a:o + b]), np.ascontiguousarray(pred[o + a:o + b])
            L.xo_recon(rs.ctypes.data_as(C.c_void_p), pr.ctypes.data_as(C.c_void_p), int(r["nnz"][c] != 0), b - a,
                       e.ctypes.data_as(C.c_void_p), 10)
            assert np.array_equal(rec[o + a:o + b], e)
    assert (exp_nnz.sum(1) > 0).sum() > 100 and set(np.unique(tq["log2_cuw"])) == {3, 4, 5, 6}


def test_fused_residue_matches_oracle(trace, gpu_ctx):
    """xb200_residue = MC -> diff -> SSD -> DCT+RDOQ -> dequant+IDCT -> recon -> SSD in one kernel"""
    rng = np.random.default_rng(11)
    mc = np.ascontiguousarray(trace.mc)
    pocs = {int(p["poc"]): i for i, p in enumerate(trace.pics) if int(p["kind"]) == 0}
    sel = [i for i in range(len(mc)) if int(mc[i]["w"]) >= 8 and int(mc[i]["poc"]) in pocs]
    sel = np.array(sel)[rng.permutation(len(sel))[:1500]]
    items = np.zeros(len(sel), api.RESIDUE_ITEM)
    off = 0
    tq = trace.tq
    for k, i in enumerate(sel):
        t = tq[int(rng.integers(0, len(tq)))]
        items[k]["mc"] = mc[i]
        items[k]["cur_pic"] = pocs[int(mc[i]["poc"])]
        items[k]["slice_type"], items[k]["run_stats"], items[k]["qp"] = 0, 7, t["qp"]
        items[k]["rate_idx"], items[k]["lambda"], items[k]["out_off"] = t["rate_idx"], t["lambda"], off
        off += int(mc[i]["w"]) * int(mc[i]["h"]) * 3 // 2
    exp_items, exp_coef, exp_rec = xo.residue_batch(trace.seq, trace.oracle_planes(), trace.rates, items, off)
    dev = items.copy()
    dev["cur_pic"] = gpu_ctx.handles[items["cur_pic"]]
    dmc = dev["mc"].copy()
    dmc = _remap(dmc, "ref_pic", gpu_ctx.handles)
    dev["mc"] = dmc
    got_items, got_coef, got_rec = gpu_ctx.residue(dev, trace.rates, off)
    for f in ("nnz", "dist_pred", "dist_rec"):
        assert np.array_equal(got_items[f], exp_items[f]), f
    assert np.array_equal(got_coef, exp_coef)
    assert np.array_equal(got_rec, exp_rec)
    assert (exp_items["nnz"].sum(1) > 0).sum() > 50


def test_bad_arguments_are_rejected(gpu_ctx):
    """error convention of the boundary: reference codes (inc/xeve.h:50-74)"""
    it = np.zeros(1, api.ME_ITEM)
    it["cur_pic"], it["ref_pic"] = 9999, 9999
    it["log2_cuw"] = it["log2_cuh"] = 4
    it["gop_size"] = 16
    with pytest.raises(api.Xb200Error) as e:
        gpu_ctx.me(it)
    assert e.value.code == api.ERR_INVALID_ARGUMENT
    it["cur_pic"] = it["ref_pic"] = int(gpu_ctx.handles[0])
    it["log2_cuw"], it["log2_cuh"] = 4, 3  # non-square CUs do not exist in Baseline
    with pytest.raises(api.Xb200Error) as e:
        gpu_ctx.me(it)
    assert e.value.code in (api.ERR_UNSUPPORTED, api.ERR_INVALID_ARGUMENT)
    assert len(gpu_ctx.me(np.zeros(0, api.ME_ITEM))) == 0


# ---- other BASELINE.json configurations: preset medium (range 64, 4 half-pel points), 10-bit input --------------
def _upload_trace(tr):
    hp = api.Hotpath(tr.seq)
    handles = []
    for i, p in enumerate(tr.pics):
        h = hp.pic_create(padded=int(p["kind"]) == 1)
        hp.pic_upload_s16(h, *(np.ascontiguousarray(a) for a in tr.planes[i]))
        handles.append(h)
    hp.handles = np.array(handles, np.int32)
    return hp


@pytest.mark.skipif(not rh.available(), reason="needs oracle/_ref to trace a live encode")
@pytest.mark.parametrize("name,preset,frames,extra,override", [
    ("cif", "medium", 20, "", dict(w=176, h=144, squares=[(32, 20, 30, 3, 2)])),           # configs 3/4 search settings
    ("2160p10", "fast", 18, "", dict(w=256, h=192, squares=[(48, 60, 40, 5, 2)], pan=(6, 2))),  # 10-bit input, faster motion
    ("cif", "fast", 20, "me_sub=3;me_sub_pos=8", dict(w=176, h=144, squares=[(32, 20, 30, 3, 2)])),  # quarter-pel stage (slow preset)
    ("cif", "fast", 20, "me_sub=1", dict(w=176, h=144, squares=[(32, 20, 30, 3, 2)])),      # integer-pel only: me_ipel_refinement
    ("cif", "fast", 20, "rdoq=0;qp=27", dict(w=176, h=144, squares=[(32, 20, 30, 3, 2)])),  # plain quantiser, lower QP
    ("cif", "fast", 20, "ref=2;me_ref_num=2;qp=24", dict(w=176, h=144, squares=[(32, 20, 30, 3, 2)])),  # two references per list
    ("cif", "fast", 12, "bframes=0", dict(w=176, h=144, squares=[(32, 20, 30, 3, 2)])),     # low delay: static range 64
    ("cif", "fast", 12, "bframes=0;inter_slice_type=1", dict(w=176, h=144, squares=[(32, 20, 30, 3, 2)])),  # P slices
])
def test_me_mc_tq_other_configs(name, preset, frames, extra, override):
    tr = tracedata.live_trace(name, frames=frames, pic_lo=1, pic_hi=2, preset=preset, extra=extra, **override)
    hp = _upload_trace(tr)
    try:
        items = np.ascontiguousarray(tr.me).astype(api.ME_ITEM)
        items = _remap(_remap(items, "cur_pic", hp.handles), "ref_pic", hp.handles)
        got = hp.me(items, tr.side)
        for f in ("mv_out", "cost", "mot_bits_out"):
            assert np.array_equal(got[f], tr.me[f]), (name, preset, f)
        recs = np.ascontiguousarray(tr.mc)
        off, total = rh.mc_offsets(recs)
        pred_ref, _, hsh, _ = rh.replay_mc(tr.live, nthreads=4)
        assert np.array_equal(hp.mc(_remap(recs.astype(api.MC_ITEM), "ref_pic", hp.handles), off, total), pred_ref)
        coef_ref, nnz_ref, _, _ = rh.replay_tq(tr.live, nthreads=4)
        tq = np.ascontiguousarray(tr.tq).astype(api.TQ_ITEM)
        it2, coef = hp.tq(tq, tr.rates, tr.tq_coef)
        m = _item_mask(tq, len(coef))
        assert np.array_equal(nnz_ref, tr.tq["nnz"])  # replay == in situ
        assert np.array_equal(it2["nnz"], nnz_ref) and np.array_equal(coef[m], coef_ref[m])
        assert len(tr.me) > 300
        # the whole CU decision under this configuration
        cu, sz, elems = tracedata.cu_slots(tr.cu)
        got, st, coef, rec = hp.analyze_cu(_cu_to_device(cu, hp.handles), tr.cu_rates, tr.cu_sbac, elems)
        tracedata.check_cu_results(got, tr.cu, coef, rec, sz, st, tr.cu_sbac)
        assert len(cu) > 300
    finally:
        hp.close()


def test_full_size_1080p_properties():
    """BASELINE config 2 size (1920x1080), properties that need no oracle run at this size:
    (1) a CU whose content is an exact integer translation of the reference is found with SAD 0 when the
        true vector lies inside the first search window; (2) residue of a perfect prediction is all zero:
        nnz = 0, dist_pred = dist_rec = 0, rec == org; (3) ME cost is invariant to which picture slot holds
        the reference; (4) bi_org of a perfect prediction equals org."""
    from xeve_b200.worklist import FrameWork, cu_grid
    w, h = 1920, 1080
    rng = np.random.default_rng(5)
    ref = [rng.integers(0, 1024, (h, w)).astype(np.int16), rng.integers(0, 1024, (h // 2, w // 2)).astype(np.int16),
           rng.integers(0, 1024, (h // 2, w // 2)).astype(np.int16)]
    dx, dy = 2, -1  # integer-pel shift (chroma shifts by 1, -0.5 -> not exact; only luma properties are asserted for ME)
    cur = [np.roll(ref[0], (-dy, -dx), (0, 1)), ref[1].copy(), ref[2].copy()]
    hp = api.Hotpath(api.make_seq(w, h, "fast"))
    try:
        r0, r1 = hp.pic_create(True), hp.pic_create(True)
        hp.pic_upload_s16(r0, *ref)
        hp.pic_upload_s16(r1, *ref)
        c = hp.pic_create(False)
        hp.pic_upload_s16(c, *cur)
        fw = FrameWork(w, h, 8, (0, 16), (0, 0), c, [r0, r1], seed=1)
        inner = (fw.x >= 64) & (fw.y >= 64) & (fw.x + (1 << fw.l2.astype(int)) <= w - 64) & (fw.y + (1 << fw.l2.astype(int)) <= h - 64)
        me = fw.me_uni.copy()
        me["mvp"] = 0
        out = hp.me(me)
        sel = np.repeat(inner, 2)
        assert (out["mv_out"][sel] == [dx * 4, dy * 4]).all()      # cur(x) = ref(x + d)
        lam = int(me["lambda_mv"][0])
        from xeve_b200 import worklist  # noqa: F401
        assert np.array_equal(out["cost"][0::2], out["cost"][1::2])  # same content in both reference slots
        # residue with the found vectors: luma residual is exactly zero inside
        bi_mc, me_bi_in = fw.build_bi(out)
        side = hp.bi_org(bi_mc, fw.bi_cur, fw.side_off, fw.side_elems)
        k = int(np.where(inner & (fw.l2 == 6))[0][0])
        s = 64
        blk = cur[0][fw.y[k]:fw.y[k] + s, fw.x[k]:fw.x[k] + s]
        assert np.array_equal(side[fw.side_off[k]:fw.side_off[k] + s * s].reshape(s, s), blk)  # 2*org - org
        me_bi = hp.me(me_bi_in, side)
        res = fw.build_residue(me_bi)
        res["run_stats"] = 1  # luma only (chroma is not translation-exact for odd shifts)
        ro, coef, rec = hp.residue(res, fw.rates, fw.res_elems)
        sel3 = np.repeat(inner, 3)
        assert (ro["nnz"][sel3, 0] == 0).all() and (ro["dist_pred"][sel3, 0] == 0).all() and (ro["dist_rec"][sel3, 0] == 0).all()
        o = int(ro["out_off"][3 * k])
        assert np.array_equal(rec[o:o + s * s].reshape(s, s), blk)
        assert len(out) == 2 * len(cu_grid(w, h)[0]) == 85800
    finally:
        hp.close()


@pytest.mark.parametrize("log2n", [5, 6])
def test_tensor_core_dct_is_bit_exact(gpu_ctx, log2n):
    """tcgen05 (fp16 operands, fp32 TMEM accumulators, hi/lo split) forward DCT == the integer transform of
    the reference (xeve_trans, src_base/xeve_tq.c:396-404), incl. saturated +-1023 residuals"""
    import ctypes as C
    n = 1 << log2n
    rng = np.random.default_rng(log2n)
    blocks = rng.integers(-1023, 1024, (300, n, n)).astype(np.int16)
    blocks[0] = 1023
    blocks[1] = -1023
    blocks[2] = np.where((np.add.outer(np.arange(n), np.arange(n)) & 1) == 0, 1023, -1023)
    blocks[3] = np.where(rng.random((n, n)) < 0.5, 1023, -1023)
    blocks[4] = 0
    blocks[5, :, :] = (np.sign(np.cos(np.pi * (2 * np.arange(n) + 1) / (2 * n)))[None, :] * 1023).astype(np.int16)
    blocks[6:40] = rng.integers(-8, 9, (34, n, n)).astype(np.int16)
    got = gpu_ctx.fwd_dct_tc(blocks, log2n)
    L = xo.lib()
    exp = blocks.copy()
    for b in exp:
        L.xo_fwd_transform(b.ctypes.data_as(C.c_void_p), log2n, log2n, 10)
    assert np.array_equal(got, exp)


def test_fused_residue_with_tensor_core_dct(trace, monkeypatch):
    """xb200_residue with XB200_TC_DCT=1: the 32/64-point forward stages run on tcgen05 inside the fused
    kernel and every output (coefficients, nnz, distortions, reconstruction) stays bit-identical"""
    monkeypatch.setenv("XB200_TC_DCT", "1")
    hp = _upload_trace(trace)
    try:
        rng = np.random.default_rng(21)
        mc = np.ascontiguousarray(trace.mc)
        pocs = {int(p["poc"]): i for i, p in enumerate(trace.pics) if int(p["kind"]) == 0}
        sel = np.array([i for i in range(len(mc)) if int(mc[i]["w"]) >= 32 and int(mc[i]["poc"]) in pocs])
        sel = sel[rng.permutation(len(sel))[:600]]
        items = np.zeros(len(sel), api.RESIDUE_ITEM)
        off = 0
        for k, i in enumerate(sel):
            t = trace.tq[int(rng.integers(0, len(trace.tq)))]
            items[k]["mc"] = mc[i]
            items[k]["cur_pic"] = pocs[int(mc[i]["poc"])]
            items[k]["slice_type"], items[k]["run_stats"], items[k]["qp"] = 0, 7, [34, 34, 34]  # low QP: many non-zero blocks
            items[k]["rate_idx"], items[k]["lambda"], items[k]["out_off"] = t["rate_idx"], t["lambda"], off
            off += int(mc[i]["w"]) * int(mc[i]["h"]) * 3 // 2
        exp_items, exp_coef, exp_rec = xo.residue_batch(trace.seq, trace.oracle_planes(), trace.rates, items, off)
        dev = items.copy()
        dev["cur_pic"] = hp.handles[items["cur_pic"]]
        dev["mc"] = _remap(dev["mc"].copy(), "ref_pic", hp.handles)
        got_items, got_coef, got_rec = hp.residue(dev, trace.rates, off)
        for f in ("nnz", "dist_pred", "dist_rec"):
            assert np.array_equal(got_items[f], exp_items[f]), f
        assert np.array_equal(got_coef, exp_coef) and np.array_equal(got_rec, exp_rec)
        assert (exp_items["nnz"][:, 0] > 0).sum() > 100
    finally:
        hp.close()


@pytest.mark.skipif(not rh.available(), reason="needs oracle/_ref (reference xeve_get_motion / xeve_get_avail_inter / xeve_get_mv_dir)")
def test_mvp_inputs_match_reference(gpu_ctx):
    """xb200_mvp vs the reference's neighbour availability, MVP candidates (incl. the (1,1) fallback, quirk q7)
    and POC-scaled temporal-direct MVs (C integer division) on random SCU maps"""
    import ctypes as C
    rng = np.random.default_rng(17)
    w_scu, h_scu = 88, 72
    f = w_scu * h_scu
    map_scu = ((rng.random(f) < 0.8).astype(np.uint32) << 31) | ((rng.random(f) < 0.2).astype(np.uint32) << 15) | \
        rng.integers(0, 1 << 15, f).astype(np.uint32)
    maps = [rng.integers(-600, 600, (f, 2, 2)).astype(np.int16) for _ in range(3)]
    n = 4000
    items = np.zeros(n, api.MVP_ITEM)
    l2 = rng.integers(3, 7, n)
    items["log2_cuw"] = items["log2_cuh"] = l2
    s = (1 << l2) >> 2
    items["x_scu"] = (rng.integers(0, w_scu, n) // s * s).clip(0, w_scu - s)
    items["y_scu"] = (rng.integers(0, h_scu, n) // s * s).clip(0, h_scu - s)
    items["lidx"] = rng.integers(0, 2, n)
    for poc, rp, lp0 in ((8, (0, 16), 0), (4, (0, 8), 8), (12, (8, 16), 0)):
        pic = np.zeros(1, api.MVP_PIC)
        pic["w_scu"], pic["h_scu"], pic["poc"], pic["ref_poc"], pic["col_list_poc0"] = w_scu, h_scu, poc, rp, lp0
        got = gpu_ctx.mvp(items, pic, map_scu, maps[0], maps[1], maps[2])
        exp = items.copy()
        L = rh.lib()
        assert L.rh_sizeof_mvp() == api.MVP_ITEM.itemsize
        L.rh_mvp.restype = None
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        L.rh_mvp.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
        L.rh_mvp(p(exp), n, p(pic), p(map_scu), p(maps[0]), p(maps[1]), p(maps[2]))
        for fld in ("avail", "refi", "mvp", "mv_dir"):
            assert np.array_equal(got[fld], exp[fld]), (fld, poc)
    assert len(np.unique(exp["avail"])) > 10 and (exp["mvp"][:, :3] == 1).all(-1).any()


def test_rdo_bit_counter_matches_oracle_and_reference(gpu_ctx):
    """xb200_rdo_bits: bits and output coder states identical to the oracle (and to the reference's counters when built)."""
    import ratedata
    it, st, coef = ratedata.work(seed=11, n=4000)
    g, sg = gpu_ctx.rdo_bits(it, st, coef)
    o, so = xo.rdo_bits_batch(it, st, coef)
    assert np.array_equal(g["bits"], o["bits"])
    assert sg.tobytes() == so.tobytes()
    if rh.available():
        r, sr = rh.rdo_bits(it, st, coef)
        assert np.array_equal(g["bits"], r["bits"]) and sg.tobytes() == sr.tobytes()
    bad = it[:1].copy()
    bad["state_in"] = len(st)
    with pytest.raises(api.Xb200Error):
        gpu_ctx.rdo_bits(bad, st, coef)


def test_rdoq_rate_tables_match_reference(trace, gpu_ctx):
    """xb200_rdoq_rates on traced coder states == the rate tables the reference derived from them."""
    if trace.source != "live":
        rng = np.random.default_rng(3)
        import ratedata
        st = ratedata.rand_states(rng, 500)
        assert gpu_ctx.rdoq_rates(st).tobytes() == xo.rdoq_rates(st, api.RATES).tobytes()
        return
    tr = trace.live
    assert gpu_ctx.rdoq_rates(tr.sbac).tobytes() == tr.rates.tobytes()
    assert xo.rdoq_rates(tr.sbac, api.RATES).tobytes() == tr.rates.tobytes()


def _cu_to_device(cu, handles):
    dev = cu.copy()
    dev["cur_pic"] = handles[cu["cur_pic"]]
    rp = cu["ref_pic"]
    dev["ref_pic"] = np.where(rp >= 0, handles[np.clip(rp, 0, len(handles) - 1)], -1)
    return dev


def test_analyze_cu_matches_reference_in_situ(trace, gpu_ctx):
    """xb200_analyze_cu == xeve_pinter_analyze_cu as it ran inside the reference encoder: best mode, IEEE-double RD cost,
    refi / mv / mvd / mvp_idx, nnz, coefficient and reconstruction hashes, output coder state -- every traced CU."""
    cu, sz, elems = tracedata.cu_slots(trace.cu)
    got, st, coef, rec = gpu_ctx.analyze_cu(_cu_to_device(cu, gpu_ctx.handles), trace.cu_rates, trace.cu_sbac, elems)
    assert len(np.unique(trace.cu["best_idx"])) == 5 and len(np.unique(trace.cu["log2_cuw"])) == 4
    tracedata.check_cu_results(got, trace.cu, coef, rec, sz, st, trace.cu_sbac)
    # and the oracle, element for element (covers the coefficients of CUs the hash check skips)
    o, so, ocoef, orec = xo.analyze_cu_batch(trace.seq, trace.oracle_planes(), trace.cu_rates, cu, trace.cu_sbac, elems)
    assert np.array_equal(coef, ocoef) and np.array_equal(rec, orec)
    assert np.array_equal(got["cost"], o["cost"]) and np.array_equal(got["mvp_idx"], o["mvp_idx"])
    assert st[cu["state_out"]].tobytes() == so[cu["state_out"]].tobytes()


# ---- deblocking ---------------------------------------------------------------------------------------
def _gpu_deblock(hp, d, expand=True):
    h = hp.pic_create(padded=True)
    hp.pic_upload_s16(h, *(np.ascontiguousarray(a) for a in d["pre"]))
    hp.deblock(h, d["cus"], d["pp"], d["map_scu"], d["map_refi"], d["map_mv"], expand=expand)
    act, full = hp.pic_download(h, with_padding=False), hp.pic_download(h, with_padding=True)
    hp.pic_destroy(h)
    return act, full


def test_deblock_matches_reference_in_situ():
    """xb200_deblock == the picture xeve_loop_filter left behind (golden fixture always; live pictures incl. another preset
    and QP when oracle/_ref is here), and the expanded borders == xeve_picbuf_expand of that picture"""
    pics = tracedata.golden_df()
    if rh.available():
        pics += tracedata.live_df(pic_hi=5) + tracedata.live_df(pic_hi=2, preset="medium", extra="qp=22")
    h, w = pics[0]["pre"][0].shape
    hp = api.Hotpath(api.make_seq(w, h))
    for d in pics:
        act, full = _gpu_deblock(hp, d)
        for k, (g, e) in enumerate(zip(act, d["post"])):
            assert np.array_equal(g, e), k
        for k, (g, e) in enumerate(zip(full, d["post"])):
            pad = 144 if k == 0 else 72
            assert np.array_equal(g, np.pad(e, pad, mode="edge")), k
    hp.close()


@pytest.mark.parametrize("w,h,seed", [(64, 64, 0), (200, 136, 1), (1920, 1080, 2), (3840, 2160, 3)])
def test_deblock_matches_oracle_synthetic(w, h, seed):
    """random quad-trees down to 4x4 with 30 % intra CUs (long chroma runs), random strengths, up to full 1080p"""
    d = tracedata.synth_df(w, h, seed, intra_frac=0.3)
    exp = xo.deblock(d["pre"], d["cus"], d["pp"], d["map_scu"], d["map_refi"], d["map_mv"])
    hp = api.Hotpath(api.make_seq(w, h))
    act, _ = _gpu_deblock(hp, d, expand=False)
    for k, (g, e) in enumerate(zip(act, exp)):
        assert np.array_equal(g, e), k
    # idempotence-style property at full size: a picture whose maps say "no edge is active" (class 3) comes back untouched
    quiet = dict(d, map_scu=(d["map_scu"] & ~np.uint32((1 << 15) | (1 << 24))), map_refi=np.zeros_like(d["map_refi"]),
                 map_mv=np.zeros_like(d["map_mv"]))
    act, _ = _gpu_deblock(hp, quiet, expand=False)
    assert all(np.array_equal(g, e) for g, e in zip(act, d["pre"]))
    # argument checks
    bad = d["cus"].copy()
    bad["x"][0] = w
    hb = hp.pic_create(padded=True)
    with pytest.raises(api.Xb200Error):
        hp.deblock(hb, bad, d["pp"], d["map_scu"], d["map_refi"], d["map_mv"])
    hp.close()


# ---- intra analysis -----------------------------------------------------------------------------------
def _gpu_intra(td):
    hp = api.Hotpath(td.seq)
    handles = []
    for i, p in enumerate(td.pics):
        h = hp.pic_create(padded=int(p["kind"]) == 1)
        hp.pic_upload_s16(h, *(np.ascontiguousarray(a) for a in td.planes[i]))
        handles.append(h)
    items, sz, elems = tracedata.intra_slots(td.intra)
    dev = items.copy()
    dev["cur_pic"] = np.array(handles, np.int32)[items["cur_pic"]]
    got, st, coef, rec = hp.analyze_intra(dev, td.cu_rates, td.cu_sbac, td.side, elems)
    tracedata.check_intra_results(got, td.intra, coef, rec, sz, st, td.cu_sbac)      # the reference's in-situ results
    exp, est, ecoef, erec = xo.analyze_intra_batch(td.seq, td.oracle_planes(), td.cu_rates, items, td.cu_sbac, td.side, elems)
    assert np.array_equal(coef, ecoef) and np.array_equal(rec, erec) and st.tobytes() == est.tobytes()  # and the oracle, byte for byte
    bad = dev[:4].copy()
    bad["log2_cuw"][0] = 7
    with pytest.raises(api.Xb200Error):
        hp.analyze_intra(bad, td.cu_rates, td.cu_sbac, td.side, elems)
    hp.close()
    return got


def test_intra_matches_reference_in_situ(monkeypatch):
    """xb200_analyze_intra == pintra_analyze_cu: golden fixture always (through the thread-per-CU kernels and, with
    XB200_INTRA_SMALL=team, through the team kernels for 4x4 / 8x8 as well), live traces (another preset / QP) with oracle/_ref"""
    _gpu_intra(tracedata.golden_intra())
    monkeypatch.setenv("XB200_INTRA_SMALL", "team")
    _gpu_intra(tracedata.golden_intra())
    monkeypatch.delenv("XB200_INTRA_SMALL")
    if rh.available():
        for kw in (dict(pic_hi=3), dict(pic_hi=1, preset="medium", extra="qp=24")):
            td = tracedata.live_intra(**kw)
            got = _gpu_intra(td)
            assert len(got) > 500


@pytest.mark.parametrize("w,h,seed,cip", [(64, 64, 0, 0), (176, 144, 1, 1), (1920, 1080, 2, 0)])
def test_intra_neighbours_match_oracle_and_reference(w, h, seed, cip):
    """xb200_intra_nbr == xeve_get_avail_intra + xeve_get_nbr + xeve_get_mpm: random pictures / COD / IF maps, with and without
    constrained intra prediction, CUs 4x4 .. 64x64 incl. the picture corners; the samples feed xb200_analyze_intra unchanged"""
    planes, items, ms, mi, ws, hs, elems = tracedata.synth_nbr(w, h, seed, n=2000)
    hp = api.Hotpath(api.make_seq(w, h))
    pic = hp.pic_create(padded=True)
    hp.pic_upload_s16(pic, *planes)
    got_it, got_side = hp.intra_nbr(pic, items, ms, mi, ws, hs, cip, elems)
    exp_it, exp_side = xo.intra_nbr(planes, items, ms, mi, ws, hs, cip, elems)
    assert np.array_equal(got_side, exp_side)
    assert np.array_equal(got_it["avail"], exp_it["avail"]) and np.array_equal(got_it["mpm"], exp_it["mpm"])
    if rh.available():
        ref_it, ref_side = rh.intra_nbr(planes, items.astype(rh.NBR_REC), ms, mi, ws, hs, cip, elems)
        assert np.array_equal(got_side, ref_side) and np.array_equal(got_it["mpm"], ref_it["mpm"]) and np.array_equal(got_it["avail"], ref_it["avail"])
    bad = items[:2].copy()
    bad["x"][0] = 2
    with pytest.raises(api.Xb200Error):
        hp.intra_nbr(pic, bad, ms, mi, ws, hs, cip, elems)
    hp.close()


@pytest.mark.skipif(not rh.available(), reason="needs oracle/_ref to trace a live encode")
@pytest.mark.parametrize("name,preset,frames,extra,override,pic_hi", [
    ("2160p10", "fast", 6, "", dict(w=256, h=192, squares=[(48, 60, 40, 5, 2)], pan=(6, 2)), 2),   # 10-bit input
    ("cif", "fast", 8, "rdoq=0;qp=27", dict(w=176, h=144, squares=[(32, 20, 30, 3, 2)]), 2),        # plain quantiser, lower QP
    ("cif", "fast", 8, "bframes=0;inter_slice_type=1", dict(w=176, h=144, squares=[(32, 20, 30, 3, 2)]), 3),  # low delay, P slices
    ("cif", "medium", 8, "qp=40", dict(w=176, h=144, squares=[(32, 20, 30, 3, 2)]), 2),              # high QP: many all-zero blocks
])
def test_intra_and_deblock_other_configs(name, preset, frames, extra, override, pic_hi):
    """intra analysis and the loop filter on other encoder configurations, against the reference's in-situ results"""
    td = tracedata.live_intra(name, frames, 0, pic_hi, preset, extra, **override)
    assert len(td.intra) > 100
    _gpu_intra(td)
    pics = tracedata.live_df(name, frames, 0, pic_hi, preset, extra, **override)
    h, w = pics[0]["pre"][0].shape
    hp = api.Hotpath(api.make_seq(w, h))
    for d in pics:
        act, _ = _gpu_deblock(hp, d)
        assert all(np.array_equal(g, e) for g, e in zip(act, d["post"]))
    hp.close()


@pytest.mark.parametrize("clip,w,h", [("1080p", 1920, 1080), ("2160p10", 3840, 2160)])
def test_intra_full_size_sample_matches_oracle(clip, w, h):
    """BASELINE.json sizes (1080p 8-bit, 2160p 10-bit input): a random sample of the frame-parallel intra work list (reference
    samples from the original picture, every CU size of the 32/16/8/4 quad-tree, positions all over the picture) -- CUDA path ==
    oracle, byte for byte"""
    from xeve_b200.clips import Clip, to_internal10
    from xeve_b200.worklist import synth_intra
    c = Clip(clip)
    planes = [to_internal10(p, c.depth) for p in c.frame(8)]
    hp = api.Hotpath(api.make_seq(w, h))
    cur = hp.pic_create(padded=False)
    hp.pic_upload_s16(cur, *planes)
    items, states, rates, side, _ = synth_intra(w, h, planes, cur, hp.rdoq_rates, seed=3)
    rng = np.random.default_rng(9)
    pick = np.concatenate([rng.permutation(np.nonzero(items["log2_cuw"] == l2)[0])[:n] for l2, n in ((2, 1500), (3, 1000), (4, 500), (5, 200))])
    sel = items[np.sort(pick)].copy()
    sel["state_out"] = 1 + np.arange(len(sel))
    sel, sz, elems = tracedata.intra_slots(sel)
    st = states[:len(sel) + 1]
    got, gst, gcoef, grec = hp.analyze_intra(sel, rates, st, side, elems)
    ora = sel.copy()
    ora["cur_pic"] = 0
    pl = (xo.PLANES * 1)()
    keep = [np.ascontiguousarray(p) for p in planes]
    pl[0].y, pl[0].u, pl[0].v = (k.ctypes.data for k in keep)
    pl[0].s_l, pl[0].s_c, pl[0].w_l, pl[0].h_l, pl[0].poc = w, w // 2, w, h, 0
    exp, est, ecoef, erec = xo.analyze_intra_batch(api.make_seq(w, h), pl, rates, ora, st, side, elems)
    for f in ("cost", "dist_cu", "ipm", "nnz", "cm_ipm_out"):
        assert np.array_equal(got[f], exp[f]), f
    assert np.array_equal(gcoef, ecoef) and np.array_equal(grec, erec) and gst.tobytes() == est.tobytes()
    assert len(np.unique(got["ipm"][:, 0])) == 5 and (got["nnz"] == 0).all(1).any() and (got["nnz"] != 0).all(1).any()
    hp.close()


def test_new_operators_edge_cases():
    """empty work lists are accepted and leave everything untouched; a deblock call without CUs still expands the borders;
    bad handles / sizes are rejected with the reference's status codes"""
    w, h = 64, 64
    hp = api.Hotpath(api.make_seq(w, h))
    d = tracedata.synth_df(w, h, 5)
    pic = hp.pic_create(padded=True)
    hp.pic_upload_s16(pic, *(np.ascontiguousarray(a) for a in d["pre"]))
    hp.deblock(pic, d["cus"][:0], d["pp"], d["map_scu"], d["map_refi"], d["map_mv"], expand=True)
    full = hp.pic_download(pic, with_padding=True)
    for k, (g, e) in enumerate(zip(full, d["pre"])):
        assert np.array_equal(g, np.pad(e, 144 if k == 0 else 72, mode="edge"))
    z = np.zeros(0, api.INTRA_ITEM)
    got, st, coef, rec = hp.analyze_intra(z, np.zeros(1, api.RATES), np.zeros(1, api.SBAC), np.zeros(8, np.int16), 0)
    assert len(got) == 0
    it, side = hp.intra_nbr(pic, np.zeros(0, api.NBR_ITEM), d["map_scu"], np.zeros(len(d["map_scu"]), np.int8), w // 4, h // 4, 0, 0)
    assert len(it) == 0
    with pytest.raises(api.Xb200Error) as e:
        hp.deblock(99, d["cus"], d["pp"], d["map_scu"], d["map_refi"], d["map_mv"])
    assert e.value.code == api.ERR_INVALID_ARGUMENT
    bad_pp = d["pp"].copy()
    bad_pp["w_scu"] = 3
    with pytest.raises(api.Xb200Error):
        hp.deblock(pic, d["cus"], bad_pp, d["map_scu"], d["map_refi"], d["map_mv"])
    hp.close()


@pytest.mark.skipif(not rh.available(), reason="needs oracle/_ref to trace a live encode")
def test_full_size_1080p_in_situ():
    """BASELINE.json configs[1] size, straight against the reference: the first two pictures of the 1080p clip encoded by the
    reference (low delay) -- all 172 k pintra_analyze_cu calls, all 42 k xeve_pinter_analyze_cu calls and both loop-filter passes,
    recomputed on the device from the traced inputs: costs (IEEE doubles), modes, coder states, coefficient / reconstruction hashes,
    deblocked pictures"""
    td = tracedata.live_trace("1080p", frames=2, pic_lo=0, pic_hi=1, preset="fast", mask=rh.TRACE_INTRA | rh.TRACE_DF | rh.TRACE_CU,
                              extra="bframes=0")
    td.intra = td.live.intra.copy()
    assert len(td.intra) > 150000 and len(td.cu) > 40000
    hp = _upload_trace(td)
    items, sz, elems = tracedata.intra_slots(td.intra)
    dev = items.copy()
    dev["cur_pic"] = hp.handles[items["cur_pic"]]
    got, st, coef, rec = hp.analyze_intra(dev, td.cu_rates, td.cu_sbac, td.side, elems)
    tracedata.check_intra_results(got, td.intra, coef, rec, sz, st, td.cu_sbac)
    cu, sz, elems = tracedata.cu_slots(td.cu)
    dcu = cu.copy()
    dcu["cur_pic"] = hp.handles[cu["cur_pic"]]
    dcu["ref_pic"] = np.where(cu["ref_pic"] >= 0, hp.handles[np.clip(cu["ref_pic"], 0, len(hp.handles) - 1)], -1)
    gcu, gst, gcoef, grec = hp.analyze_cu(dcu, td.cu_rates, td.cu_sbac, elems)
    tracedata.check_cu_results(gcu, td.cu, gcoef, grec, sz, gst, td.cu_sbac)
    for d in tracedata.df_from_trace(td.live):
        act, _ = _gpu_deblock(hp, d)
        assert all(np.array_equal(g, e) for g, e in zip(act, d["post"]))
    hp.close()
