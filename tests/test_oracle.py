"""CPU tests (no GPU): the oracle restatement against (a) the golden fixture traced from the
reference and (b) the compiled reference itself when oracle/_ref is present."""
import ctypes as C

import numpy as np
import pytest

import tracedata
from oracle import oracle as xo
from oracle import refharness as rh

needs_ref = pytest.mark.skipif(not rh.available(), reason="oracle/_ref (compiled reference) not built here")
p = lambda a: a.ctypes.data_as(C.c_void_p)


def test_oracle_me_matches_golden(golden):
    out = xo.me_batch(golden.seq, golden.oracle_planes(), golden.side, np.ascontiguousarray(golden.me))
    for f in ("mv_out", "cost", "mot_bits_out"):
        assert np.array_equal(out[f], golden.me[f]), f
    assert (golden.me["bi"] == 1).sum() > 100


def test_oracle_mc_matches_golden(golden):
    mc = np.ascontiguousarray(golden.mc)
    off, total = rh.mc_offsets(mc)
    assert np.array_equal(off, golden.expect["mc_off"])
    pred = xo.mc_batch(golden.seq, golden.oracle_planes(), mc, off, total)
    assert np.array_equal(pred, golden.expect["mc_pred"])


def test_oracle_tq_itdq_match_golden(golden):
    items, coef, resi = xo.tq_batch(golden.seq, np.ascontiguousarray(golden.tq), golden.rates, golden.tq_coef)
    assert np.array_equal(items["nnz"], golden.expect["tq_nnz"])
    assert np.array_equal(coef, golden.expect["tq_coef_out"])
    for r, nz in zip(items, items["nnz"]):  # the reference leaves nnz == 0 planes untouched
        o, ny = int(r["in_off"]), 1 << (2 * int(r["log2_cuw"]))
        for c, (a, b) in enumerate(((0, ny), (ny, ny + ny // 4), (ny + ny // 4, ny + ny // 2))):
            if nz[c]:
                assert np.array_equal(resi[o + a:o + b], golden.expect["tq_resi_out"][o + a:o + b])


@needs_ref
def test_kernels_match_compiled_reference():
    """SAD / SSD / SATD / transforms / MC vs the reference's C, SSE and AVX2 tables"""
    L, R = xo.lib(), rh.lib()
    rng = np.random.default_rng(1)
    for l2 in (2, 3, 4, 5, 6):
        n = 1 << l2
        for trial in range(12):
            a = rng.integers(0, 1024, (n + 8, n + 40)).astype(np.int16)
            b = rng.integers(-1023, 2047, (n + 8, n + 40)).astype(np.int16)
            if trial == 0:
                a[:], b[:] = 1023, 0
            sa, sb = a.shape[1], b.shape[1]
            assert L.xo_sad(n, n, p(a), sa, p(b), sb, 10) == R.rh_sad(0, l2, l2, p(a), p(b), sa, sb, 10) == R.rh_sad(2, l2, l2, p(a), p(b), sa, sb, 10)
            assert L.xo_ssd(n, n, p(a), sa, p(b), sb, 10) == R.rh_ssd(1, l2, l2, p(a), p(b), sa, sb, 10) == R.rh_ssd(0, l2, l2, p(a), p(b), sa, sb, 10)
            assert L.xo_satd(n, n, p(a), sa, p(b), sb, 10) == R.rh_satd(0, n, n, p(a), p(b), sa, sb, 10) == R.rh_satd(1, n, n, p(a), p(b), sa, sb, 10)
            blk = rng.integers(-1023, 1024, (n, n)).astype(np.int16)
            if trial == 1:
                blk[:] = 1023
            if trial == 2:
                blk[:] = -1023
            x1, x2, x3 = blk.copy(), blk.copy(), blk.copy()
            L.xo_fwd_transform(p(x1), l2, l2, 10)
            R.rh_fwd_transform(0, p(x2), l2, l2, 10)
            R.rh_fwd_transform(2, p(x3), l2, l2, 10)
            assert np.array_equal(x1, x2) and np.array_equal(x1, x3)
            # inverse: encoder-plausible coefficient ranges (the reference's C butterflies form partial
            # sums in 32-bit int and its AVX2 path multiplies in 32 bits; all three agree here)
            cf = (rng.integers(-2000, 2000, (n, n)) * (rng.random((n, n)) < 0.15)).astype(np.int16)
            y1, y2, y3 = cf.copy(), cf.copy(), cf.copy()
            L.xo_inv_transform(p(y1), l2, l2, 10)
            R.rh_inv_transform(0, p(y2), l2, l2, 10)
            R.rh_inv_transform(2, p(y3), l2, l2, 10)
            assert np.array_equal(y1, y2) and np.array_equal(y1, y3)
    ref = rng.integers(0, 1024, (200, 200)).astype(np.int16)
    for trial in range(200):
        w = int(rng.choice([4, 8, 16, 32, 64]))
        gx, gy = int(rng.integers(160, 400)), int(rng.integers(160, 400))
        if trial % 3 == 0:
            gx &= ~3
        if trial % 5 == 0:
            gy &= ~3
        o1, o2, o3 = (np.zeros((w, w), np.int16) for _ in range(3))
        L.xo_mc_luma(p(ref), 200, gx, gy, gx, gy, p(o1), w, w, w, 10)
        R.rh_mc_l(0, p(ref), gx << 2, gy << 2, 200, w, p(o2), w, w, 10)
        R.rh_mc_l(2, p(ref), gx << 2, gy << 2, 200, w, p(o3), w, w, 10)
        assert np.array_equal(o1, o2) and np.array_equal(o1, o3)
        L.xo_mc_chroma(p(ref), 200, gx, gy, gx, gy, p(o1), w, w, w, 10)
        R.rh_mc_c(0, p(ref), gx << 2, gy << 2, 200, w, p(o2), w, w, 10)
        assert np.array_equal(o1, o2)


@needs_ref
def test_oracle_matches_reference_in_situ_cif(trace):
    """every pi->fn_me / pi->fn_mc / ctx->fn_tq call of 3 CIF pictures, as the reference ran them"""
    assert trace.source == "live"
    opl = trace.oracle_planes()
    out = xo.me_batch(trace.seq, opl, trace.side, np.ascontiguousarray(trace.me))
    for f in ("mv_out", "cost", "mot_bits_out"):
        assert np.array_equal(out[f], trace.me[f]), f
    mc = np.ascontiguousarray(trace.mc)
    off, total = rh.mc_offsets(mc)
    pred_ref, _, hsh, _ = rh.replay_mc(trace.live, nthreads=4)
    assert np.array_equal(hsh, mc["out_hash"])
    assert np.array_equal(xo.mc_batch(trace.seq, opl, mc, off, total), pred_ref)
    coef_ref, nnz_ref, resi_ref, _ = rh.replay_tq(trace.live, nthreads=4)
    items, coef, resi = xo.tq_batch(trace.seq, np.ascontiguousarray(trace.tq), trace.rates, trace.tq_coef)
    assert np.array_equal(nnz_ref, trace.tq["nnz"]) and np.array_equal(items["nnz"], nnz_ref)
    m = np.zeros(len(coef), bool)
    for r in trace.tq:
        o = int(r["in_off"])
        m[o:o + ((3 << (2 * int(r["log2_cuw"]))) >> 1)] = True
    assert np.array_equal(coef[m], coef_ref[m])
    for r, nz in zip(items, nnz_ref):
        o, ny = int(r["in_off"]), 1 << (2 * int(r["log2_cuw"]))
        for c, (a, b) in enumerate(((0, ny), (ny, ny + ny // 4), (ny + ny // 4, ny + ny // 2))):
            if nz[c]:
                assert np.array_equal(resi[o + a:o + b], resi_ref[o + a:o + b])


@needs_ref
def test_padding_matches_reference(trace):
    for i, pc in enumerate(trace.pics):
        if int(pc["kind"]) != 1:
            continue
        full = trace.live.plane_views(i)
        mine = trace.padded_planes(i)
        for a, b in zip(mine, full):
            assert np.array_equal(a, b[:, : a.shape[1]])


@pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built")
def test_rdo_bit_counter_matches_reference():
    """xo_rdo_bits == the reference's xeve_rdo_bit_cnt_* + xeve_get_bit_number, bits and resulting coder states."""
    import ratedata
    it, st, coef = ratedata.work()
    a, sa = rh.rdo_bits(it, st, coef)
    b, sb = xo.rdo_bits_batch(it, st, coef)
    assert a["bits"].max() > 1000 and len(np.unique(it["kind"])) == 4
    assert np.array_equal(a["bits"], b["bits"])
    assert sa.tobytes() == sb.tobytes()


@pytest.mark.skipif(not rh.available(), reason="oracle/_ref not built")
def test_rdoq_rate_tables_match_reference():
    """xo_rdoq_rates(state) == the rdoq_est_* tables the reference derived from the same state (xeve_rdoq_bit_est)."""
    import tracedata
    td = tracedata.live_trace(name="cif", frames=12, pic_lo=1, pic_hi=2, mask=4)
    tr = td.live
    assert len(tr.sbac) == len(tr.rates) > 100
    o = xo.rdoq_rates(tr.sbac, rh.RATES)
    assert o.tobytes() == tr.rates.tobytes()


def test_rate_estimation_matches_golden():
    """The oracle's bit counter and RDOQ rate tables against reference outputs committed in tests/golden/rate_golden.npz."""
    import os
    import ratedata
    from xeve_b200 import api
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "rate_golden.npz"))
    it, st, coef = ratedata.work()
    b, sb = xo.rdo_bits_batch(it, st, coef)
    assert np.array_equal(b["bits"], z["bits"])
    assert sb.tobytes() == z["states"].tobytes()
    assert xo.rdoq_rates(z["sbac"], api.RATES).tobytes() == z["rates"].tobytes()


def test_oracle_analyze_cu_matches_golden(golden):
    """xo_analyze_cu (SKIP / DIRECT / L0 / L1 / BI decision with cbf RDO) == xeve_pinter_analyze_cu in situ: mode, double
    cost, motion, nnz, coefficient and reconstruction hashes and the output coder state."""
    td = golden
    cu, sz, elems = tracedata.cu_slots(td.cu)
    o, st, coef, rec = xo.analyze_cu_batch(td.seq, td.oracle_planes(), td.cu_rates, cu, td.cu_sbac, elems)
    assert len(np.unique(td.cu["best_idx"])) == 5
    tracedata.check_cu_results(o, td.cu, coef, rec, sz, st, td.cu_sbac)


@needs_ref
def test_oracle_analyze_cu_matches_reference_in_situ_cif(trace):
    td = trace
    cu, sz, elems = tracedata.cu_slots(td.cu)
    o, st, coef, rec = xo.analyze_cu_batch(td.seq, td.oracle_planes(), td.cu_rates, cu, td.cu_sbac, elems)
    assert len(cu) > 5000
    tracedata.check_cu_results(o, td.cu, coef, rec, sz, st, td.cu_sbac)


# ---- deblocking ---------------------------------------------------------------------------------------
def test_oracle_deblock_matches_golden():
    """xo_deblock == the picture the reference's xeve_loop_filter left behind, on the committed fixture (incl. the intra
    picture whose 4x4 CUs make chroma edges 2 samples apart)"""
    pics = tracedata.golden_df()
    assert len(pics) == 3 and (pics[0]["cus"]["log2_cuw"] == 2).sum() >= 8
    for d in pics:
        got = xo.deblock(d["pre"], d["cus"], d["pp"], d["map_scu"], d["map_refi"], d["map_mv"])
        assert any(not np.array_equal(a, b) for a, b in zip(d["pre"], d["post"]))  # the filter did something
        for g, e in zip(got, d["post"]):
            assert np.array_equal(g, e)


@needs_ref
def test_oracle_deblock_matches_reference_live_and_synthetic():
    """in-situ results of a live encode (two presets), then the reference's exported edge filters replayed over random
    well-formed inputs, in coding order and in shuffled order (exercises the right-edge / COD branches)"""
    for kw in (dict(pic_hi=4), dict(pic_hi=2, preset="medium", extra="qp=22")):
        for d in tracedata.live_df(**kw):
            got = xo.deblock(d["pre"], d["cus"], d["pp"], d["map_scu"], d["map_refi"], d["map_mv"])
            assert all(np.array_equal(g, e) for g, e in zip(got, d["post"]))
    rng = np.random.default_rng(3)
    for seed, (w, h) in enumerate(((64, 64), (200, 136), (352, 288))):
        d = tracedata.synth_df(w, h, seed, intra_frac=0.3)
        for order in (np.arange(len(d["cus"])), rng.permutation(len(d["cus"]))):
            exp, _ = rh.deblock(d["pre"], d["cus"][order], d["pp"], d["map_scu"], d["map_refi"], d["map_mv"])
            got = xo.deblock(d["pre"], d["cus"][order], d["pp"], d["map_scu"], d["map_refi"], d["map_mv"])
            assert all(np.array_equal(g, e) for g, e in zip(got, exp))
            assert not np.array_equal(exp[1], d["pre"][1])


# ---- intra analysis -----------------------------------------------------------------------------------
def _oracle_intra(td):
    items, sz, elems = tracedata.intra_slots(td.intra)
    got, st, coef, rec = xo.analyze_intra_batch(td.seq, td.oracle_planes(), td.cu_rates, items, td.cu_sbac, td.side, elems)
    tracedata.check_intra_results(got, td.intra, coef, rec, sz, st, td.cu_sbac)
    return got


def test_oracle_intra_matches_golden():
    """xo_analyze_intra == pintra_analyze_cu in situ on the committed fixture: cost (IEEE double), modes, nnz, coefficient /
    reconstruction hashes, output coder state incl. the intra_dir models; CU sizes 4x4 .. 64x64, I and B slices"""
    td = tracedata.golden_intra()
    got = _oracle_intra(td)
    assert set(np.unique(td.intra["log2_cuw"])) == {2, 3, 4, 5, 6} and len(np.unique(got["ipm"][:, 0])) == 5


@needs_ref
def test_oracle_intra_matches_reference_live():
    for kw in (dict(pic_hi=3), dict(pic_hi=1, preset="medium", extra="qp=24")):
        td = tracedata.live_intra(**kw)
        assert len(td.intra) > 500
        _oracle_intra(td)


@needs_ref
def test_oracle_intra_neighbours_match_reference():
    """xo_intra_nbr == xeve_get_avail_intra + xeve_get_nbr (Y, U, V) + xeve_get_mpm on random pictures / maps, with and without
    constrained intra prediction; every availability pattern incl. picture corners"""
    for seed, (w, h), cip in ((0, (64, 64), 0), (1, (176, 144), 1), (2, (352, 288), 0)):
        planes, items, ms, mi, ws, hs, elems = tracedata.synth_nbr(w, h, seed)
        exp_it, exp_side = rh.intra_nbr(planes, items.astype(rh.NBR_REC), ms, mi, ws, hs, cip, elems)
        got_it, got_side = xo.intra_nbr(planes, items, ms, mi, ws, hs, cip, elems)
        assert np.array_equal(got_side, exp_side)
        assert np.array_equal(got_it["avail"], exp_it["avail"]) and np.array_equal(got_it["mpm"], exp_it["mpm"])
        assert len(np.unique(got_it["avail"])) > 8


OTHER_CONFIGS = [
    ("2160p10", "fast", 6, "", dict(w=256, h=192, squares=[(48, 60, 40, 5, 2)], pan=(6, 2)), 2),   # 10-bit input
    ("cif", "fast", 8, "rdoq=0;qp=27", dict(tracedata.QCIF), 2),                                     # plain quantiser, lower QP
    ("cif", "fast", 8, "bframes=0;inter_slice_type=1", dict(tracedata.QCIF), 3),                     # low delay, P slices
    ("cif", "medium", 8, "qp=40", dict(tracedata.QCIF), 2),                                          # high QP: many all-zero blocks
]


@needs_ref
@pytest.mark.parametrize("name,preset,frames,extra,override,pic_hi", OTHER_CONFIGS)
def test_oracle_intra_and_deblock_other_configs(name, preset, frames, extra, override, pic_hi):
    override = {k: v for k, v in override.items() if k != "n"}
    td = tracedata.live_intra(name, frames, 0, pic_hi, preset, extra, **override)
    assert len(td.intra) > 100
    _oracle_intra(td)
    for d in tracedata.live_df(name, frames, 0, pic_hi, preset, extra, **override):
        got = xo.deblock(d["pre"], d["cus"], d["pp"], d["map_scu"], d["map_refi"], d["map_mv"])
        assert all(np.array_equal(g, e) for g, e in zip(got, d["post"]))


import os  # noqa: E402

FULL_SIZE = [("1080p", "fast", 150000, 40000)]
if os.environ.get("XB200_SLOW_TESTS"):  # two more minutes of CPU: BASELINE.json configs[2] (688 961 intra calls, 170 196 inter CU decisions)
    FULL_SIZE.append(("2160p10", "medium", 600000, 160000))


@needs_ref
@pytest.mark.parametrize("clip,preset,min_intra,min_cu", FULL_SIZE)
def test_oracle_full_size_in_situ(clip, preset, min_intra, min_cu):
    """BASELINE.json configs[1] size (and configs[2] = 2160p 10-bit medium with XB200_SLOW_TESTS=1): the first two pictures of the clip
    encoded by the reference (low delay, so the second is an inter picture) -- every pintra_analyze_cu call (172 k at 1080p, CU 4x4 ..
    64x64), every xeve_pinter_analyze_cu call (42 k) and both loop-filter passes, reproduced by the oracle from the traced inputs:
    costs as IEEE doubles, coder states, coefficient and reconstruction hashes, deblocked pictures"""
    td = tracedata.live_trace(clip, frames=2, pic_lo=0, pic_hi=1, preset=preset,
                              mask=rh.TRACE_INTRA | rh.TRACE_DF | rh.TRACE_CU | rh.TRACE_LCU, extra="bframes=0")
    td.intra = td.live.intra.copy()
    assert len(td.intra) > min_intra and len(td.cu) > min_cu
    _oracle_intra(td)
    cu, sz, elems = tracedata.cu_slots(td.cu)
    ocu, ost, ocoef, orec = xo.analyze_cu_batch(td.seq, td.oracle_planes(), td.cu_rates, cu, td.cu_sbac, elems)
    tracedata.check_cu_results(ocu, td.cu, ocoef, orec, sz, ost, td.cu_sbac)
    for d in tracedata.df_from_trace(td.live):
        got = xo.deblock(d["pre"], d["cus"], d["pp"], d["map_scu"], d["map_refi"], d["map_mv"])
        assert all(np.array_equal(g, e) for g, e in zip(got, d["post"]))
    # the same two pictures through the decision chain: the oracle picks every CU itself (3 174 + 1 512 leaf CUs at 1080p)
    out = tracedata.chain_sequence(*tracedata.chain_inputs_from_trace(td.live))
    assert len(out[0]["intra_log"]) + len(out[1]["intra_log"]) == len(td.intra) and len(out[1]["cu_log"]) == len(td.cu)


@needs_ref
def test_oracle_mvp_inputs_match_reference():
    """xo_mvp == xeve_get_avail_inter + xeve_get_motion + xeve_get_mv_dir on random SCU maps (same generator as the GPU test):
    availability masks, the (1,1) fallback (quirk q7), the unscaled colocated candidate, POC-scaled temporal-direct MVs"""
    from xeve_b200 import api
    rng = np.random.default_rng(17)
    w_scu, h_scu = 88, 72
    f = w_scu * h_scu
    map_scu = ((rng.random(f) < 0.8).astype(np.uint32) << 31) | ((rng.random(f) < 0.2).astype(np.uint32) << 15) | \
        rng.integers(0, 1 << 15, f).astype(np.uint32) | ((rng.random(f) < 0.05).astype(np.uint32) << 26)
    maps = [rng.integers(-600, 600, (f, 2, 2)).astype(np.int16) for _ in range(3)]
    n = 4000
    items = np.zeros(n, api.MVP_ITEM)
    l2 = rng.integers(3, 7, n)
    items["log2_cuw"] = items["log2_cuh"] = l2
    s = (1 << l2) >> 2
    items["x_scu"] = (rng.integers(0, w_scu, n) // s * s).clip(0, w_scu - s)
    items["y_scu"] = (rng.integers(0, h_scu, n) // s * s).clip(0, h_scu - s)
    items["lidx"] = rng.integers(0, 2, n)
    L = rh.lib()
    assert L.rh_sizeof_mvp() == api.MVP_ITEM.itemsize
    L.rh_mvp.restype = None
    L.rh_mvp.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
    for poc, rp, lp0 in ((8, (0, 16), 0), (4, (0, 8), 8), (12, (8, 16), 0), (16, (0, 0), 0)):
        pic = np.zeros(1, api.MVP_PIC)
        pic["w_scu"], pic["h_scu"], pic["poc"], pic["ref_poc"], pic["col_list_poc0"] = w_scu, h_scu, poc, rp, lp0
        exp = items.copy()
        L.rh_mvp(p(exp), n, p(pic), p(map_scu), p(maps[0]), p(maps[1]), p(maps[2]))
        got = xo.mvp_batch(items, pic, map_scu, maps[0], maps[1], maps[2])
        for fld in ("avail", "refi", "mvp", "mv_dir"):
            assert np.array_equal(got[fld], exp[fld]), (fld, poc)
    assert len(np.unique(exp["avail"])) > 10 and (exp["mvp"][:, :3] == 1).all(-1).any()


# ---- CU decision chain: mode_coding_tree over whole pictures (SURVEY 8a' q15 / q16) ---------------------------------------------
def test_oracle_decision_chain_matches_golden():
    """three pictures (intra, then two bi-predicted ones referencing the earlier two) encoded by the oracle alone from the original
    pictures and the picture-level parameters: xo_chain_picture -> xo_deblock -> xo_pad_plane, each result feeding the next picture.
    Coder state before / after every CTU, frame maps, leaf CUs and the deblocked pictures equal what the reference produced."""
    seq, pics = tracedata.chain_golden()
    out = tracedata.chain_sequence(seq, pics)
    assert [r["slice_type"] for r in out] == [2, 0, 0]
    assert len(out[0]["intra_log"]) > 500 and len(out[2]["cu_log"]) > 100 and len(out[2]["intra_log"]) > 0
    modes = np.concatenate([r["cu_log"]["best_idx"] for r in out[1:]])
    assert len(np.unique(modes)) >= 4                                   # SKIP, DIRECT, uni- and bi-prediction all won somewhere
    sizes = np.concatenate([r["cus"]["log2_cuw"] for r in out])
    assert set(np.unique(sizes)) >= {2, 3, 4, 5}                        # leaf CUs from 4x4 to 32x32 (64x64 ones: the live tests)


@needs_ref
def test_oracle_decision_chain_whole_sequence():
    """all 20 pictures of the default hierarchical-B GOP (intra picture, anchors, four B layers incl. the odd POCs whose early CU
    termination threshold differs): the oracle's own reconstruction is the only reference data later pictures see, and every
    picture still equals the reference's -- a whole-sequence statement of rows a1 - a14, f-1, f-2, f-3 and the tree around them."""
    seq, pics = tracedata.live_chain(frames=20)
    out = tracedata.chain_sequence(seq, pics)
    assert len(out) == 20 and sorted(r["poc"] for r in out) == list(range(20))
    calls = [len(r["cu_log"]) for r in out[1:]]
    assert min(calls) < max(calls)                                      # early termination skipped sub-trees in some pictures


CHAIN_CONFIGS = [
    ("2160p10", "fast", 6, "", dict(w=256, h=192, squares=[(48, 60, 40, 5, 2)], pan=(6, 2))),        # 10-bit input
    ("cif", "fast", 8, "rdoq=0;qp=27", dict(tracedata.QCIF)),                                          # plain quantiser, lower QP
    ("cif", "fast", 8, "bframes=0;inter_slice_type=1", dict(tracedata.QCIF)),                          # low delay, P slices
    ("cif", "medium", 8, "qp=40", dict(tracedata.QCIF)),                                               # high QP: many all-zero blocks
    ("cif", "medium", 6, "qp=22", dict(w=352, h=288)),   # CIF 352x288, low QP
    ("cif", "fast", 10, "ref=2;me_ref_num=2;qp=24", dict(tracedata.QCIF)),                             # two references per list
    ("cif", "fast", 8, "me_sub=3;me_sub_pos=8", dict(tracedata.QCIF)),                                 # quarter-pel stage
    ("cif", "fast", 8, "qp=12", dict(tracedata.QCIF)),                                                 # QP extremes
    ("cif", "fast", 8, "qp=50", dict(tracedata.QCIF)),
    ("cif", "fast", 6, "", dict(w=200, h=136, squares=[(32, 20, 30, 3, 2)])),                          # CTUs cut by both picture edges
    ("cif", "fast", 6, "closed_gop=1;keyint=4", dict(tracedata.QCIF)),                                 # several intra pictures
    ("cif", "fast", 6, "bframes=0;ref=3;me_ref_num=3", dict(tracedata.QCIF)),                          # low delay, three references
    ("cif", "medium", 6, "merge_num=4", dict(tracedata.QCIF)),                                         # all four skip candidates
]


@needs_ref
@pytest.mark.parametrize("name,preset,frames,extra,override", CHAIN_CONFIGS)
def test_oracle_decision_chain_other_configs(name, preset, frames, extra, override):
    """P slices need the bitstream-order state walk (chain_eco): their RDO counter codes a direct_mode_flag the bitstream lacks"""
    seq, pics = tracedata.live_chain(name, frames, preset, extra, **override)
    out = tracedata.chain_sequence(seq, pics)
    assert len(out) == frames


INJECT_CONFIGS = [
    ("cif", "fast", 20, "", dict(tracedata.QCIF)),                                                      # default GOP, 20 pictures
    ("2160p10", "medium", 5, "", dict(w=256, h=192, squares=[(48, 60, 40, 5, 2)], pan=(6, 2))),        # 10-bit, medium
    ("cif", "fast", 6, "bframes=0;inter_slice_type=1", dict(tracedata.QCIF)),                          # P slices
    ("cif", "fast", 4, "qp=24", dict(w=352, h=288)),     # CIF 352x288 (BASELINE configs[0] size)
    ("1080p", "fast", 2, "bframes=0", dict(w=1920, h=1080)),                                           # BASELINE configs[1] size: 1 020 CTUs, 387 KB
]

if os.environ.get("XB200_SLOW_TESTS"):  # two more minutes: BASELINE configs[2] size (3840x2160 10-bit medium): 4 080 CTUs, 1.5 MB, byte-identical
    INJECT_CONFIGS.append(("2160p10", "medium", 2, "bframes=0", dict(w=3840, h=2160)))


@needs_ref
@pytest.mark.parametrize("name,preset,frames,extra,override", INJECT_CONFIGS)
def test_bitstream_is_bit_exact_with_the_decision_pass_replaced(name, preset, frames, extra, override):
    """the picture-level boundary end to end: the reference encoder with ctx->fn_mode_analyze_lcu replaced by "take the decisions
    from outside" (per 4x4 unit: mode, CU size, motion, MVP index, MVD, intra mode, nnz; coefficient planes; the picture before
    deblocking), fed with what the oracle chain decided from the original pictures alone, writes the byte-identical bitstream --
    its own entropy coder, loop filter and picture management run unchanged and none of its inter / intra analyses run"""
    ref_bs, got_bs, n_ctu, ref_calls, out = tracedata.chain_inject_roundtrip(name, frames, preset, extra, **override)
    assert n_ctu == sum(len(r["ctu"]) for r in out) and ref_calls == 0
    assert len(ref_bs) > 1000 and np.array_equal(ref_bs, got_bs)


@needs_ref
@pytest.mark.parametrize("threads,override", [(2, dict(tracedata.QCIF)), (4, dict(w=352, h=288))])
def test_oracle_decision_chain_threaded_reference(threads, override):
    """the reference with threads = n decides CTU rows y, y + n, ... as one coder-state chain per thread (each reset at its first row,
    a CTU waiting for its upper-right neighbour only); the oracle reproduces those decisions too -- n independent chains per picture"""
    seq, pics = tracedata.live_chain("cif", 6, "fast", "", threads=threads, **override)
    assert int(pics[0]["pp"]["parallel_rows"]) == threads
    first_of_rows = pics[1]["expect"]["state_in"][::(int(np.asarray(seq).reshape(-1)[0]["w"]) + 63) // 64]
    assert all(int(s["range"]) == 16384 and (s["m"] == 512).all() for s in first_of_rows[:threads])   # every chain starts from reset
    tracedata.chain_sequence(seq, pics)


# ---- Main profile, first piece: the two-stage 16-bit "IQT" transforms ---------------------------------------------
MAIN_REF = os.path.join(os.path.dirname(rh._LIB_PATH), "libxeve_main_ref.so")


@pytest.mark.skipif(not os.path.exists(MAIN_REF), reason="oracle/_ref/libxeve_main_ref.so (Main-profile reference) not built here")
def test_oracle_iqt_transforms_match_main_reference():
    """xo_iqt_fwd / xo_iqt_inv == the Main-profile reference's two-stage transforms (xeve_trans / xeve_itrans with tool_iqt: tx_pbN then
    tx_pbM with an s16 between the stages) for every block shape the tree can produce (2..64 per side, aspect ratio <= 4), 8- and
    10-bit: its C table for any s16 input, its AVX2 table (what it runs with) for residuals of the coded bit depth"""
    M, L = C.CDLL(MAIN_REF), xo.lib()
    fn_t = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_int)

    def aligned(n):
        raw = np.zeros(n + 64, np.int16)
        off = (-raw.ctypes.data % 128) // 2
        return raw[off:off + n]
    rng = np.random.default_rng(5)
    big_a, big_t = aligned(4096 + 256), aligned(4096 + 256)
    for names in (("xeve_tbl_tx", "xeve_tbl_itx"), ("xeve_tbl_tx_avx", "xeve_tbl_itx_avx")):
        tx = [fn_t(a) for a in (C.c_void_p * 6).in_dll(M, names[0])]
        itx = [fn_t(a) for a in (C.c_void_p * 6).in_dll(M, names[1])]
        n = 0
        for bd in (8, 10):
            for lw in range(1, 7):
                for lh in range(max(1, lw - 2), min(6, lw + 2) + 1):
                    w, h = 1 << lw, 1 << lh
                    for amp in ((64, (1 << bd) - 1) if "avx" in names[0] else (64, (1 << bd) - 1, 32767)):
                        blk = rng.integers(-amp, amp + 1, w * h).astype(np.int16)
                        a, t = big_a[:w * h], big_t[:w * h]
                        a[:] = blk
                        tx[lw - 1](p(a), p(t), lw - 1 + bd - 8, h)
                        tx[lh - 1](p(t), p(a), lh + 6, w)
                        got = blk.copy()
                        L.xo_iqt_fwd(p(got), lw, lh, bd)
                        assert np.array_equal(a, got), ("fwd", names[0], bd, lw, lh, amp)
                        if w == 64:
                            assert not got.reshape(h, w)[:, 32:].any()          # the 64-point stage keeps 32 outputs
                        a[:] = got
                        itx[lh - 1](p(a), p(t), 7, w)
                        itx[lw - 1](p(t), p(a), 12 - (bd - 8), h)
                        back = got.copy()
                        L.xo_iqt_inv(p(back), lw, lh, bd)
                        assert np.array_equal(a, back), ("inv", names[1], bd, lw, lh, amp)
                        n += 1
        assert n >= 96


@pytest.mark.skipif(not os.path.exists(MAIN_REF), reason="oracle/_ref/libxeve_main_ref.so (Main-profile reference) not built here")
def test_oracle_ats_transforms_match_main_reference():
    """xo_ats_fwd / xo_ats_inv == xeve_t_MxN_ats_intra / xeve_it_MxN_ats_intra of the Main-profile reference (DST-VII / DCT-VIII per
    direction, 4..32 points, aspect ratio <= 4, 8 / 10 bit; C and SSE stage tables), and the generated 8-bit matrices equal its table"""
    M, L = C.CDLL(MAIN_REF), xo.lib()
    tbl = np.ctypeslib.as_array((C.c_int8 * (2 * 4 * 1024)).in_dll(M, "xevem_tbl_tr")).reshape(2, 4, 1024)
    for typ in range(2):                                   # 0 DCT-VIII, 1 DST-VII (src_base/xeve_def.h:557)
        for l2 in range(2, 6):
            m = np.zeros(1 << (2 * l2), np.int8)
            L.xo_ats_matrix(typ, l2, p(m))
            assert np.array_equal(m, tbl[typ, l2 - 2, :1 << (2 * l2)]), (typ, l2)
    rng = np.random.default_rng(8)
    func_itrans = C.c_void_p.in_dll(M, "xeve_func_itrans")
    n = 0
    for inv_tbl in ("xeve_itrans_map_tbl", "xeve_itrans_map_tbl_sse"):
        func_itrans.value = C.addressof((C.c_void_p * 80).in_dll(M, inv_tbl))
        for bd in (8, 10):
            for lw in range(2, 6):
                for lh in range(max(2, lw - 2), min(5, lw + 2) + 1):
                    for tridx in range(4):
                        w, h = 1 << lw, 1 << lh
                        blk = rng.integers(-(1 << bd) + 1, 1 << bd, w * h).astype(np.int16)
                        a = blk.copy()
                        M.xeve_t_MxN_ats_intra(p(a), w, h, bd, 1, tridx)
                        got = blk.copy()
                        L.xo_ats_fwd(p(got), lw, lh, bd, tridx)
                        assert np.array_equal(a, got), ("fwd", bd, lw, lh, tridx)
                        M.xeve_it_MxN_ats_intra(p(a), w, h, bd, 15, tridx, 0, 0)
                        L.xo_ats_inv(p(got), lw, lh, bd, tridx)
                        assert np.array_equal(a, got), ("inv", inv_tbl, bd, lw, lh, tridx)
                        n += 1
    assert n >= 200
