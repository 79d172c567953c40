"""Work lists for the parity tests.

Two sources, same shape:
  * live  -- oracle/_ref is built (this container, and the GPU box where the prebuilt binary
             travels): run the UNMODIFIED reference encoder on a seeded clip with logging hooks
             on pi->fn_me / pi->fn_mc / ctx->fn_tq and take the recorded calls + in-situ results;
  * golden -- tests/golden/qcif_trace.npz, produced by tests/golden/make_golden.py from the
             same harness in this container, for machines without oracle/_ref.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as xo  # noqa: E402
from oracle import refharness as rh  # noqa: E402
from xeve_b200.clips import Clip  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "qcif_trace.npz")
QCIF = dict(w=176, h=144, n=20, squares=[(32, 20, 30, 3, 2)])


def clip_yuv(name, frames, **override):
    c = Clip(name, **override)
    dt = np.uint8 if c.depth == 8 else np.dtype("<u2")
    return c, np.frombuffer(b"".join(c.frame_bytes(i) for i in range(frames)), dt)


class TraceData:
    """Uniform view over a live rh.Trace or the golden fixture."""

    def __init__(self, seq, pics, planes, me, mc, tq, rates, side, tq_coef, source, expect=None):
        self.seq, self.pics, self.planes = seq, pics, planes  # planes[i] = (Y, U, V) ACTIVE areas (s16 2-D)
        self.me, self.mc, self.tq, self.rates = me, mc, tq, rates
        self.side = side          # s16 buffer addressed by me.org_bi_off
        self.tq_coef = tq_coef    # s16 buffer addressed by tq.in_off (inputs)
        self.source = source
        self.expect = expect or {}
        self._padded = {}

    def padded_planes(self, i):
        """(Y, U, V) with the 144/72 border replicated by the ORACLE (reference src_base/xeve_util.c:190-248)."""
        if i not in self._padded:
            out = []
            for k, a in enumerate(self.planes[i]):
                pad = 144 if k == 0 else 72
                h, w = a.shape
                buf = np.zeros((h + 2 * pad, w + 2 * pad), np.int16)
                buf[pad:pad + h, pad:pad + w] = a
                xo.lib().xo_pad_plane(buf.ctypes.data_as(C.c_void_p), buf.shape[1], w, h, pad)
                out.append(buf)
            self._padded[i] = out
        return self._padded[i]

    def oracle_planes(self):
        """ctypes xo_planes array (active origins inside oracle-padded copies); keep self alive."""
        arr = (xo.PLANES * len(self.pics))()
        for i, p in enumerate(self.pics):
            if int(p["kind"]) == 1:
                bufs = self.padded_planes(i)
                pads = (144, 72, 72)
            else:
                bufs = [np.ascontiguousarray(a) for a in self.planes[i]]
                self._padded[i] = bufs
                pads = (0, 0, 0)
            ptrs = [b.ctypes.data + 2 * (pd * b.shape[1] + pd) for b, pd in zip(bufs, pads)]
            arr[i].y, arr[i].u, arr[i].v = ptrs
            arr[i].s_l, arr[i].s_c = bufs[0].shape[1], bufs[1].shape[1]
            arr[i].w_l, arr[i].h_l, arr[i].poc = int(p["w_l"]), int(p["h_l"]), int(p["poc"])
        return arr


def from_live(tr: "rh.Trace") -> TraceData:
    planes = []
    for i, p in enumerate(tr.pics):
        full = tr.plane_views(i)
        pl, pc = int(p["pad_l"]), int(p["pad_c"])
        w, h = int(p["w_l"]), int(p["h_l"])
        planes.append((full[0][pl:pl + h, pl:pl + w].copy(), full[1][pc:pc + h // 2, pc:pc + w // 2].copy(),
                       full[2][pc:pc + h // 2, pc:pc + w // 2].copy()))
    td = TraceData(tr.const.copy(), tr.pics.copy(), planes, tr.me.copy(), tr.mc.copy(), tr.tq.copy(), tr.rates.copy(), tr.samp,
                   tr.samp, "live")
    td.live = tr
    td.cu, td.cu_sbac, td.cu_rates = tr.cu.copy(), tr.cu_sbac.copy(), tr.rates.copy()
    return td


def live_trace(name="cif", frames=30, pic_lo=1, pic_hi=3, preset="fast", mask=15, extra="", **override) -> TraceData:
    c, yuv = clip_yuv(name, frames, **override)
    tr = rh.encode_clip(yuv, frames, c.w, c.h, in_depth=c.depth, preset=preset, extra=extra, trace_mask=mask, pic_lo=pic_lo, pic_hi=pic_hi)
    return from_live(tr)


def golden_trace() -> TraceData:
    z = np.load(GOLDEN)
    pics = z["pics"]
    planes = [(z[f"p{i}_y"], z[f"p{i}_u"], z[f"p{i}_v"]) for i in range(len(pics))]
    expect = dict(mc_pred=z["mc_pred"], mc_off=z["mc_off"], tq_coef_out=z["tq_coef_out"], tq_resi_out=z["tq_resi_out"],
                  tq_nnz=z["tq_nnz"])
    td = TraceData(z["seq"], pics, planes, z["me"], z["mc"], z["tq"], z["rates"], z["side"], z["tq_in"], "golden", expect)
    td.cu, td.cu_sbac, td.cu_rates = z["cu"], z["cu_sbac"], z["cu_rates"]
    return td


INTRA_GOLDEN = os.path.join(ROOT, "tests", "golden", "intra_golden.npz")


def intra_slots(items):
    """(items with out_off assigned, slot sizes, total elements) for the coef / rec outputs of an intra CU list"""
    items = items.copy()
    sz = (3 << (2 * items["log2_cuw"].astype(np.int64))) >> 1
    items["out_off"] = np.concatenate([[0], np.cumsum(sz)[:-1]])
    return items, sz, int(sz.sum())


def live_intra(name="cif", frames=20, pic_lo=0, pic_hi=2, preset="fast", extra="", **override) -> TraceData:
    """ctx->fn_pintra_analyze_cu calls of the traced pictures with the reference's in-situ results (td.intra, td.cu_sbac,
    td.cu_rates, td.side = neighbour samples addressed by nb_off)"""
    override = override or QCIF
    td = live_trace(name, frames, pic_lo, pic_hi, preset, rh.TRACE_INTRA, extra, **override)
    td.intra = td.live.intra.copy()
    return td


def golden_intra() -> TraceData:
    z = np.load(INTRA_GOLDEN)
    pics = z["pics"]
    planes = [(z[f"p{i}_y"], z[f"p{i}_u"], z[f"p{i}_v"]) for i in range(len(pics))]
    e = np.empty(0)
    td = TraceData(z["seq"], pics, planes, e, e, e, z["rates"], z["side"], e, "golden")
    td.intra, td.cu_sbac, td.cu_rates = z["intra"], z["sbac"], z["rates"]
    return td


def check_intra_results(got, ref, coef, rec, sz, st_got, st_ref):
    """analyze_intra outputs against the reference's in-situ results"""
    for f in ("cost", "dist_cu", "ipm", "nnz", "cm_ipm_out"):
        assert np.array_equal(got[f], ref[f]), f
    hc, hr = xo.hash_slots(coef, got["out_off"], sz), xo.hash_slots(rec, got["out_off"], sz)
    assert np.array_equal(hc, ref["coef_hash"]) and np.array_equal(hr, ref["rec_hash"])
    assert st_got[ref["state_out"]].tobytes() == st_ref[ref["state_out"]].tobytes()


def synth_nbr(w, h, seed, n=600):
    """Random intra-neighbourhood queries: a picture of random samples, random COD / IF bits and intra modes per SCU, random
    CUs (4x4 .. 64x64) inside the picture.  Returns (planes, items, map_scu, map_ipm, w_scu, h_scu, side_elems)."""
    from xeve_b200 import api
    rng = np.random.default_rng(seed)
    ws, hs = w // 4, h // 4
    planes = [rng.integers(0, 1024, (h, w)).astype(np.int16), rng.integers(0, 1024, (h // 2, w // 2)).astype(np.int16),
              rng.integers(0, 1024, (h // 2, w // 2)).astype(np.int16)]
    map_scu = ((rng.random(ws * hs) < 0.7).astype(np.uint32) << 31) | ((rng.random(ws * hs) < 0.5).astype(np.uint32) << 15)
    map_ipm = rng.integers(0, 5, ws * hs).astype(np.int8)
    items = np.zeros(n, api.NBR_ITEM)
    l2 = rng.integers(2, 7, n)
    l2 = np.minimum(l2, int(np.log2(min(w, h))))
    s = 1 << l2
    items["log2_cuw"] = items["log2_cuh"] = l2
    items["x"] = (rng.integers(0, 1 << 16, n) % ((w - s) // 4 + 1)) * 4
    items["y"] = (rng.integers(0, 1 << 16, n) % ((h - s) // 4 + 1)) * 4
    items["x"][:8] = [0, w - 4, 0, w - 4, 0, 4, w - 8, 0][:8]   # corners and borders
    items["y"][:8] = [0, 0, h - 4, h - 4, 4, 0, h - 8, h - 8][:8]
    items["log2_cuw"][:8] = items["log2_cuh"][:8] = [2, 2, 2, 2, 2, 2, 3, 3]
    sz = 8 * (1 << items["log2_cuw"].astype(np.int64)) + 6
    items["nb_off"] = np.concatenate([[0], np.cumsum(sz)[:-1]])
    return planes, items, map_scu, map_ipm, ws, hs, int(sz.sum())


DF_GOLDEN = os.path.join(ROOT, "tests", "golden", "df_golden.npz")


def df_from_trace(tr):
    """Deblocking inputs / in-situ results of a live trace recorded with rh.TRACE_DF: list of dicts {pre, post: (Y, U, V) active
    areas, cus, pp, map_scu, map_refi, map_mv}."""
    def act(i):
        p, full = tr.pics[i], tr.plane_views(i)
        pl, pc, w, h = int(p["pad_l"]), int(p["pad_c"]), int(p["w_l"]), int(p["h_l"])
        return [full[0][pl:pl + h, pl:pl + w].copy(), full[1][pc:pc + h // 2, pc:pc + w // 2].copy(),
                full[2][pc:pc + h // 2, pc:pc + w // 2].copy()]
    out = []
    for r in tr.df:
        assert int(r["on"]) == 1
        f = int(r["pp"]["w_scu"]) * int(r["pp"]["h_scu"])
        ms, mr, mm = rh.df_maps(tr.df_maps, int(r["maps_off"]), f)
        out.append(dict(pre=act(int(r["pre_pic"])), post=act(int(r["post_pic"])), pp=r["pp"].copy(), map_scu=ms, map_refi=mr,
                        map_mv=mm, cus=tr.df_cu[int(r["cu_first"]):int(r["cu_first"] + r["cu_cnt"])].copy(), poc=int(r["poc"])))
    return out


def live_df(name="cif", frames=20, pic_lo=0, pic_hi=2, preset="fast", extra="", **override):
    """Deblocking inputs / in-situ results of the traced pictures (see df_from_trace)."""
    override = override or QCIF
    c, yuv = clip_yuv(name, frames, **override)
    tr = rh.encode_clip(yuv, frames, c.w, c.h, in_depth=c.depth, preset=preset, extra=extra, trace_mask=rh.TRACE_DF, pic_lo=pic_lo,
                        pic_hi=pic_hi)
    return df_from_trace(tr)


def golden_df():
    z = np.load(DF_GOLDEN)
    return [dict(pre=[z[f"pre{i}_{k}"] for k in "yuv"], post=[z[f"post{i}_{k}"] for k in "yuv"], cus=z[f"cus{i}"], pp=z[f"pp{i}"],
                 map_scu=z[f"map_scu{i}"], map_refi=z[f"map_refi{i}"], map_mv=z[f"map_mv{i}"]) for i in range(int(z["n"]))]


def synth_df(w, h, seed, intra_frac=0.1):
    from xeve_b200.worklist import synth_deblock
    return synth_deblock(w, h, seed, intra_frac)


def cu_slots(cu):
    """(items with out_off assigned, per-item slot sizes, total elements) for the coef / rec outputs of a CU list."""
    cu = cu.copy()
    sz = (3 << (2 * cu["log2_cuw"].astype(np.int64))) >> 1
    cu["out_off"] = np.concatenate([[0], np.cumsum(sz)[:-1]])
    return cu, sz, int(sz.sum())


def check_cu_results(got, ref, coef, rec, sz, st_got, st_ref):
    """Decision-level comparison of analyze_cu outputs with the reference's in-situ results (fields the reference leaves
    stale -- MVs of unused lists, mvp_idx of SKIP/DIRECT, coefficients of all-zero CUs -- are not compared)."""
    assert np.array_equal(got["best_idx"], ref["best_idx"])
    assert np.array_equal(got["cost"], ref["cost"])          # IEEE doubles, bit for bit
    assert np.array_equal(got["nnz"], ref["nnz"]) and np.array_equal(got["refi"], ref["refi"])
    used = ref["refi"] >= 0
    for f in ("mv", "mvd"):
        assert not ((got[f] != ref[f]).any(2) & used & (ref["best_idx"] != 4)[:, None]).any(), f
    assert not ((got["mvp_idx"] != ref["mvp_idx"]) & used & (ref["best_idx"] < 3)[:, None]).any()
    assert st_got[ref["state_out"]].tobytes() == st_ref[ref["state_out"]].tobytes()
    hc, hr = xo.hash_slots(coef, got["out_off"], sz), xo.hash_slots(rec, got["out_off"], sz)
    nz = ref["nnz"].any(1)
    assert np.array_equal(hc[nz], ref["coef_hash"][nz]) and np.array_equal(hr, ref["rec_hash"])


def get_trace() -> TraceData:
    if rh.available():
        return live_trace()
    return golden_trace()


# ---- CU decision chain (mode_coding_tree over whole pictures / a whole sequence) --------------------------------------
CHAIN_GOLDEN = os.path.join(ROOT, "tests", "golden", "chain_golden.npz")


def chain_dtypes():
    from xeve_b200 import api
    return rh.LCU_REC, rh.DF_CU, api.CU_ITEM, api.INTRA_ITEM


def chain_inputs_from_trace(tr):
    """(seq, pictures in coding order) of a live trace recorded with rh.TRACE_LCU | rh.TRACE_DF from picture 0.  Per picture:
    pp = the picture-level parameters a rate controller / reference-list manager supplies (LCU_REC of CTU 0: slice type, QP,
    lambdas, reference POCs, CU size limits), df_pp (DF_PIC), org = the original picture, expect = what the reference produced
    (per-CTU coder states, frame maps, leaf CUs, the picture before / after deblocking)."""
    td = from_live(tr)
    dfs = {d["poc"]: d for d in df_from_trace(tr)}
    pics = []
    for idx in np.flatnonzero(tr.lcu["lcu_num"] == 0):
        pp = tr.lcu[idx].copy()
        poc, d = int(pp["poc"]), dfs[int(pp["poc"])]
        recs = tr.lcu[tr.lcu["poc"] == poc]
        recs = recs[np.argsort(recs["lcu_num"], kind="stable")]     # threads > 1: CTU rows are recorded as they finish
        pics.append(dict(pp=pp, df_pp=d["pp"], org=td.planes[int(pp["cur_pic"])],
                         expect=dict(state_in=recs["state_in"].copy(), state_out=recs["state_out"].copy(), map_scu=d["map_scu"],
                                     map_refi=np.asarray(d["map_refi"]), map_mv=np.asarray(d["map_mv"]), cus=d["cus"], pre=d["pre"],
                                     post=d["post"])))
    return td.seq, pics


def chain_golden():
    z = np.load(CHAIN_GOLDEN)
    pics = []
    for i in range(int(z["n"])):
        g = lambda k: z[f"{k}{i}"]  # noqa: E731
        pics.append(dict(pp=g("pp")[0], df_pp=g("df_pp")[0], org=[g("org_y"), g("org_u"), g("org_v")],
                         expect=dict(state_in=g("state_in"), state_out=g("state_out"), map_scu=g("map_scu"), map_refi=g("map_refi"),
                                     map_mv=g("map_mv"), cus=g("cus"), pre=None, post=[g("post_y"), g("post_u"), g("post_v")])))
    return z["seq"], pics


class ChainEncoder:
    """Picture-by-picture encoder built from the oracle chain: encode(pic) runs xo_chain_picture -> deblock -> pad and keeps the result
    as a reference picture (+ colocated MV map) for later pictures; adopt() takes a reference picture produced somewhere else (another
    rank).  Nothing of the reference's own reconstruction is read; with check=True every encoded picture's coder states, frame maps,
    leaf CUs and pictures before / after deblocking are asserted equal to pic["expect"].
    hp (an xeve_b200.api.Hotpath, used together with chain_with): picture handles are the device's, original pictures are uploaded,
    and every reconstructed picture is deblocked and border-expanded ON THE DEVICE (xb200_deblock), where it stays as the reference
    picture of later xb200_analyze_cu / xb200_mc calls; the host copy is only downloaded for the comparison."""

    def __init__(self, seq, check=True, hp=None):
        self.seq, self.check, self.hp = seq, check, hp
        self.bd = int(np.asarray(seq).reshape(-1)[0]["bit_depth"])
        self.planes = (xo.PLANES * 4096)()
        self.done, self.keep, self.next_handle = {}, [], 0

    def _handle(self):
        self.next_handle += 1
        assert self.next_handle <= 4096
        return self.next_handle - 1

    def _bind_reference(self, poc, post, map_mv, h_rec, extra=None):
        padded = []
        for a, pad in zip(post, (144, 72, 72)):
            hh, ww = a.shape
            buf = np.zeros((hh + 2 * pad, ww + 2 * pad), np.int16)
            buf[pad:pad + hh, pad:pad + ww] = a
            xo.lib().xo_pad_plane(buf.ctypes.data_as(C.c_void_p), buf.shape[1], ww, hh, pad)
            padded.append(buf)
        pl = self.planes[h_rec]
        pl.y, pl.u, pl.v = [b.ctypes.data + 2 * (pd * b.shape[1] + pd) for b, pd in zip(padded, (144, 72, 72))]
        pl.s_l, pl.s_c = padded[0].shape[1], padded[1].shape[1]
        pl.w_l, pl.h_l, pl.poc = post[0].shape[1], post[0].shape[0], poc
        r = dict(extra or {})
        r.update(poc=poc, post=post, padded=padded, handle=h_rec, map_mv=np.ascontiguousarray(map_mv, np.int16))
        self.done[poc] = r
        return r

    def adopt(self, poc, post, map_mv):
        post = [np.ascontiguousarray(a, np.int16) for a in post]
        if self.hp is not None:
            h_rec = self.hp.pic_create(padded=True)
            self.hp.pic_upload_s16(h_rec, *post)
        else:
            h_rec = self._handle()
        return self._bind_reference(poc, post, map_mv, h_rec)

    def encode(self, pc):
        hp, planes = self.hp, self.planes
        pp = np.array(pc["pp"]).reshape(1).copy()
        poc = int(pp["poc"][0])
        org = [np.ascontiguousarray(a) for a in pc["org"]]
        self.keep.append(org)
        if hp is not None:
            h_org = hp.pic_create(padded=False)
            hp.pic_upload_s16(h_org, *org)
            assert 0 <= h_org < 4096
        else:
            h_org = self._handle()
        planes[h_org].y, planes[h_org].u, planes[h_org].v = [a.ctypes.data for a in org]
        planes[h_org].s_l, planes[h_org].s_c = org[0].shape[1], org[1].shape[1]
        planes[h_org].w_l, planes[h_org].h_l, planes[h_org].poc = org[0].shape[1], org[0].shape[0], poc
        pp["cur_pic"] = h_org
        col = [None, None]
        for l in range(2):
            for k in range(4):
                if int(pp["ref_pic"][0][l][k]) < 0:
                    continue
                rp = self.done[int(pp["ref_poc"][0][l][k])]     # our own (or an adopted) reconstruction of that POC
                pp["ref_pic"][0][l][k] = rp["handle"]
                if k == 0:
                    col[l] = rp["map_mv"]
        r = xo.chain_picture(self.seq, planes, pp, col[0], col[1], chain_dtypes())
        if hp is not None:
            h_rec = hp.pic_create(padded=True)
            hp.pic_upload_s16(h_rec, *r["rec"])
            hp.deblock(h_rec, r["cus"], pc["df_pp"], r["map_scu"], r["map_refi"], r["map_mv"], expand=True)
            post = [np.ascontiguousarray(a) for a in hp.pic_download(h_rec, False)]
            assert 0 <= h_rec < 4096
        else:
            h_rec = self._handle()
            post = xo.deblock(r["rec"], r["cus"], pc["df_pp"], r["map_scu"], r["map_refi"], r["map_mv"], bit_depth=self.bd)
        r["slice_type"] = int(pp["slice_type"][0])
        r = self._bind_reference(poc, post, r["map_mv"], h_rec, extra=r)
        if self.check:
            e = pc["expect"]
            assert len(e["state_in"]) == len(r["ctu"])
            assert r["ctu"]["state_in"].tobytes() == e["state_in"].tobytes() and r["ctu"]["state_out"].tobytes() == e["state_out"].tobytes(), poc
            assert np.array_equal(r["map_scu"] & 0x81FF8000, e["map_scu"] & 0x81FF8000), poc   # coded, luma cbf, skip, QP, intra
            assert np.array_equal(r["map_refi"].reshape(-1), e["map_refi"].reshape(-1)), poc
            assert np.array_equal(r["map_mv"].reshape(-1), e["map_mv"].reshape(-1)), poc
            assert len(r["cus"]) == len(e["cus"]) and all(np.array_equal(r["cus"][k], e["cus"][k]) for k in ("x", "y", "log2_cuw", "log2_cuh")), poc
            assert e["pre"] is None or all(np.array_equal(a, b) for a, b in zip(r["rec"], e["pre"])), poc
            assert all(np.array_equal(a, b) for a, b in zip(post, e["post"])), poc
        return r


def chain_sequence(seq, pics, check=True, hp=None):
    """Encode the pictures in the order given with one ChainEncoder (see there)."""
    enc = ChainEncoder(seq, check=check, hp=hp)
    return [enc.encode(pc) for pc in pics]


def live_chain(name="cif", frames=20, preset="fast", extra="", threads=1, **override):
    override = override or QCIF
    override = {k: v for k, v in override.items() if k != "n"}
    c, yuv = clip_yuv(name, frames, **override)
    tr = rh.encode_clip(yuv, frames, c.w, c.h, in_depth=c.depth, preset=preset, extra=extra, threads=threads,
                        trace_mask=rh.TRACE_LCU | rh.TRACE_DF, pic_lo=0, pic_hi=1 << 20, want_bitstream=False)
    return chain_inputs_from_trace(tr)


def chain_inject_roundtrip(name="cif", frames=20, preset="fast", extra="", **override):
    """(reference bitstream, bitstream of the reference with its mode decision replaced by the oracle chain's decisions, CTUs injected,
    analyses the reference still ran) -- see rh.encode_clip_injected"""
    override = override or QCIF
    override = {k: v for k, v in override.items() if k != "n"}
    c, yuv = clip_yuv(name, frames, **override)
    tr = rh.encode_clip(yuv, frames, c.w, c.h, in_depth=c.depth, preset=preset, extra=extra, trace_mask=rh.TRACE_LCU | rh.TRACE_DF,
                        pic_lo=0, pic_hi=1 << 20)
    out = chain_sequence(*chain_inputs_from_trace(tr))
    dec = [dict(poc=r["poc"], scu=r["scu"], coef=r["coef"], rec=r["rec"]) for r in out]
    tr2, n, ncu, nintra = rh.encode_clip_injected(yuv, frames, c.w, c.h, dec, in_depth=c.depth, preset=preset, extra=extra)
    return tr.bitstream, tr2.bitstream, n, ncu + nintra, out


# ---- the decision chain with the per-CU analyses done by somebody else (the CUDA library, or the oracle again as a plumbing check)
_CB6 = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)


def _view(ptr, dtype, n):
    dtype = np.dtype(dtype)
    return np.frombuffer((C.c_char * (dtype.itemsize * n)).from_address(ptr), dtype, n)


class chain_with:
    """Context manager: xo_chain_picture calls analyze_cu(items, rates, states, elems) -> (items, states, coef, rec),
    mc(items, off, total) -> pred and analyze_intra(items, rates, states, side, elems) -> (items, states, coef, rec) -- the signatures
    of xeve_b200.api.Hotpath -- for every CU instead of its own restatements.  Errors raised inside a callback are re-raised on exit."""

    def __init__(self, analyze_cu, mc, analyze_intra, mvp=None, intra_nbr=None, upload=None):
        """mvp(items, pic, map_scu, map_mv, col0, col1) -> items and intra_nbr(handle, items, map_scu, map_ipm, w_scu, h_scu, cip,
        side_elems) -> (items, side) (optional) also replace the chain's xo_mvp / xo_intra_nbr; upload(rec) -> handle of a device
        picture holding the reconstruction as it stands (what xb200_intra_nbr reads)."""
        from xeve_b200 import api
        self.err, self.n_cu, self.n_intra, self.n_mvp, self.n_nbr = None, 0, 0, 0, 0
        self.in_cbs = (None, None)
        if mvp is not None:
            def mvp_cb(it_p, pic_p):
                if self.err:
                    return
                try:
                    cur = xo.chain_current
                    it = _view(it_p, api.MVP_ITEM, 1)
                    it[:] = mvp(it.copy(), _view(pic_p, api.MVP_PIC, 1).copy(), cur["map_scu"], cur["map_mv"], cur["col0"], cur["col1"])
                    self.n_mvp += 1
                except Exception as e:  # noqa: BLE001
                    self.err = e

            def nbr_cb(it_p, side_p):
                if self.err:
                    return
                try:
                    cur = xo.chain_current
                    it = _view(it_p, api.NBR_ITEM, 1)
                    n = 8 * (1 << int(it["log2_cuw"][0])) + 6
                    items, side = intra_nbr(upload(cur["rec"]), it.copy(), cur["map_scu"], cur["map_ipm"], cur["w_scu"], cur["h_scu"],
                                            cur["cip"], n)
                    it[:] = items
                    _view(side_p, np.int16, n)[:] = side
                    self.n_nbr += 1
                except Exception as e:  # noqa: BLE001
                    self.err = e
            self.in_cbs = (C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)(mvp_cb), C.CFUNCTYPE(None, C.c_void_p, C.c_void_p)(nbr_cb))

        def cu_cb(cu_p, st_p, rates_p, coef_p, rec_p, pred_p):
            if self.err:
                return
            try:
                cu, st, rates = _view(cu_p, api.CU_ITEM, 1), _view(st_p, api.SBAC, 2), _view(rates_p, api.RATES, 1)
                n = 1 << (2 * int(cu["log2_cuw"][0]))
                elems = 3 * n // 2
                items, states, coef, rec = analyze_cu(cu.copy(), rates.copy(), st.copy(), elems)
                cu[:], st[:] = items, states
                _view(coef_p, np.int16, elems)[:], _view(rec_p, np.int16, elems)[:] = coef, rec
                m = np.zeros(1, api.MC_ITEM)                       # the winner's prediction (mi->pred_y_best): one more pi->fn_mc call
                m["poc"], m["x"], m["y"], m["w"], m["h"] = cu["poc"], cu["x"], cu["y"], 1 << int(cu["log2_cuw"][0]), 1 << int(cu["log2_cuw"][0])
                for l in range(2):
                    r = int(items["refi"][0][l])
                    m["refi"][0][l], m["mv"][0][l] = r, items["mv"][0][l]
                    m["ref_pic"][0][l] = int(cu["ref_pic"][0][l][r]) if r >= 0 else -1
                    m["ref_poc"][0][l] = int(cu["ref_poc"][0][l][r]) if r >= 0 else -1
                _view(pred_p, np.int16, n)[:] = mc(m, np.zeros(1, np.int64), elems)[:n]
                self.n_cu += 1
            except Exception as e:  # noqa: BLE001 -- must not propagate through the C frame
                self.err = e

        def intra_cb(it_p, st_p, rates_p, side_p, coef_p, rec_p):
            if self.err:
                return
            try:
                it, st, rates = _view(it_p, api.INTRA_ITEM, 1), _view(st_p, api.SBAC, 2), _view(rates_p, api.RATES, 1)
                w = 1 << int(it["log2_cuw"][0])
                elems = 3 * w * w // 2
                items, states, coef, rec = analyze_intra(it.copy(), rates.copy(), st.copy(), _view(side_p, np.int16, 8 * w + 6).copy(), elems)
                it[:], st[:] = items, states
                _view(coef_p, np.int16, elems)[:], _view(rec_p, np.int16, elems)[:] = coef, rec
                self.n_intra += 1
            except Exception as e:  # noqa: BLE001
                self.err = e
        self.cbs = (_CB6(cu_cb), _CB6(intra_cb))

    def __enter__(self):
        L = xo.lib()
        L.xo_chain_set_callbacks.restype = None
        L.xo_chain_set_callbacks.argtypes = [C.c_void_p, C.c_void_p]
        L.xo_chain_set_callbacks(*self.cbs)
        L.xo_chain_set_input_callbacks.restype = None
        L.xo_chain_set_input_callbacks.argtypes = [C.c_void_p, C.c_void_p]
        L.xo_chain_set_input_callbacks(*self.in_cbs)
        return self

    def __exit__(self, *exc):
        xo.lib().xo_chain_set_callbacks(None, None)
        xo.lib().xo_chain_set_input_callbacks(None, None)
        if self.err and exc[0] is None:
            raise self.err
        return False


# ---- the whole decision pass ON THE DEVICE (xb200_analyze_picture): the persistent chain kernel decides every CU of the picture ------
def picture_record(pp, df_pp, cur_pic, rec_pic, ref_handles, unfiltered=-1, deblock=1, threads=None):
    """xb200_picture from the picture-level fields of an LCU_REC (what the reference's control plane supplies) + the loop-filter inputs"""
    from xeve_b200 import api
    pp = np.asarray(pp).reshape(-1)[0]
    p = np.zeros(1, api.PICTURE)
    for k in ("poc", "slice_type", "tile_qp", "num_refp", "ref_poc", "col_list_poc0", "max_cu_inter", "min_cu_inter", "max_cu_intra",
              "min_cu_intra", "cip", "qp", "lambda_mv", "max_search_range", "lambda", "sqrt_lambda0", "dist_chroma_weight"):
        p[k] = pp[k]
    p["parallel_rows"] = int(pp["parallel_rows"]) if threads is None else threads
    p["cur_pic"], p["rec_pic"], p["unfiltered_pic"], p["deblock"] = cur_pic, rec_pic, unfiltered, deblock
    p["ref_pic"] = -1
    for l in range(2):
        for k in range(4):
            if int(pp["ref_pic"][l][k]) >= 0:
                p["ref_pic"][0][l][k] = ref_handles[int(pp["ref_poc"][l][k])]
    p["df"] = np.asarray(df_pp).reshape(-1)[0]
    return p


def leaf_cus_of(scu, w, h):
    """leaf CUs in coding order (z-scan) from the per-CTU unit records -- what xeve_deblock_tree enumerates"""
    w_lcu, out = (w + 63) // 64, []

    def walk(lcu, xc, yc, x, y, L):
        if x >= w or y >= h:
            return
        u = scu[lcu][((y - yc) >> 2) * 16 + ((x - xc) >> 2)]
        if int(u["log2"]) == L + 2 or L == 0:
            out.append((x, y, L + 2))
            return
        half = 2 << L
        for part in range(4):
            walk(lcu, xc, yc, x + (part & 1) * half, y + (part >> 1) * half, L - 1)
    for lcu in range(len(scu)):
        xc, yc = (lcu % w_lcu) << 6, (lcu // w_lcu) << 6
        walk(lcu, xc, yc, xc, yc, 4)
    return np.array(out, np.int64).reshape(-1, 3)


class DevicePictureEncoder:
    """Picture-by-picture encoder on the device: encode(pic) = xb200_analyze_picture (decision pass + loop filter + border expansion in
    one enqueue); the reconstruction stays on the device as the reference of later pictures.  With check=True everything the reference
    left behind for the picture (per-CTU coder states, frame maps, leaf CUs, the picture before / after deblocking) is compared.
    compare_oracle=True additionally runs the oracle chain on the same inputs and compares every CU analysis in call order (debug)."""

    def __init__(self, seq, hp, check=True, compare_oracle=False, threads=None, sync=True):
        self.seq, self.hp, self.check, self.compare_oracle, self.threads, self.sync = seq, hp, check, compare_oracle, threads, sync
        self.w, self.h = int(np.asarray(seq).reshape(-1)[0]["w"]), int(np.asarray(seq).reshape(-1)[0]["h"])
        self.handles, self.pending, self.results = {}, [], []
        self.oracle = ChainEncoder(seq, check=False, hp=None) if compare_oracle else None
        if compare_oracle:
            hp.picture_log_enable(self.hp.f_scu * 2, self.hp.f_scu * 2)

    def enqueue(self, pc):
        hp = self.hp
        pp = np.asarray(pc["pp"]).reshape(-1)[0]
        poc = int(pp["poc"])
        org = [np.ascontiguousarray(a) for a in pc["org"]]
        h_org, h_rec, h_pre = hp.pic_create(padded=False), hp.pic_create(padded=True), hp.pic_create(padded=True)
        hp.pic_upload_s16(h_org, *org)
        rec = picture_record(pp, pc["df_pp"], h_org, h_rec, self.handles, unfiltered=h_pre, threads=self.threads)
        hp.analyze_picture(rec)
        self.handles[poc] = h_rec
        self.pending.append((pc, poc, h_org, h_rec, h_pre))

    def collect(self):
        hp = self.hp
        for pc, poc, h_org, h_rec, h_pre in self.pending:
            log = hp.picture_log(h_rec) if self.compare_oracle else None
            r = hp.picture_fetch(h_rec)
            r.update(hp.picture_maps(h_rec))
            r["poc"], r["rec"] = poc, [np.ascontiguousarray(a) for a in hp.pic_download(h_pre, False)]
            r["post"] = [np.ascontiguousarray(a) for a in hp.pic_download(h_rec, False)]
            r["cus"] = leaf_cus_of(r["scu"], self.w, self.h)
            hp.pic_destroy(h_pre)
            hp.pic_destroy(h_org)
            if self.compare_oracle:
                o = self.oracle.encode(pc)
                cu, it = log
                assert len(cu) == len(o["cu_log"]) and len(it) == len(o["intra_log"]), (poc, len(cu), len(o["cu_log"]), len(it), len(o["intra_log"]))
                for name, a, b, skip in (("inter", cu, o["cu_log"], ("cur_pic", "ref_pic", "coef_hash", "rec_hash", "me_first", "me_cnt", "pad0_", "pad1_",
                                                                     "state_in", "state_out", "rate_idx", "out_off")),
                                         ("intra", it, o["intra_log"], ("cur_pic", "coef_hash", "rec_hash", "pad0_", "pad1_", "state_in",
                                                                        "state_out", "rate_idx", "out_off", "nb_off"))):
                    for f in a.dtype.names:
                        if f in skip:
                            continue
                        bad = [i for i in range(len(a)) if not np.array_equal(a[f][i], b[f][i])]
                        assert not bad, f"POC {poc}: {name} CU analysis #{bad[0]} of {len(a)} differs in {f}: device {a[f][bad[0]]} oracle {b[f][bad[0]]} " \
                                        f"(x {a['x'][bad[0]]} y {a['y'][bad[0]]} log2 {a['log2_cuw'][bad[0]]})"
            if self.check:
                e = pc["expect"]
                st = r["states"].view(rh.STATE) if r["states"].dtype.itemsize == rh.STATE.itemsize else r["states"]
                assert len(e["state_in"]) == len(st), poc
                for col, key in ((0, "state_in"), (1, "state_out")):
                    ex = np.asarray(e[key]).reshape(-1)
                    bad = [i for i in range(len(st)) if st[i, col].tobytes() != ex[i].tobytes()]
                    w_lcu = (self.w + 63) >> 6
                    assert not bad, f"POC {poc}: coder {key} differs at {len(bad)} of {len(st)} CTUs, first CTU {bad[0]} (x {bad[0] % w_lcu} y {bad[0] // w_lcu})"
                assert np.array_equal(r["map_scu"] & 0x81FF8000, e["map_scu"] & 0x81FF8000), poc   # coded, luma cbf, skip, QP, intra
                assert np.array_equal(r["map_refi"].reshape(-1), np.asarray(e["map_refi"]).reshape(-1)), poc
                assert np.array_equal(r["map_mv"].reshape(-1), np.asarray(e["map_mv"]).reshape(-1)), poc
                ec = e["cus"]
                assert len(r["cus"]) == len(ec) and np.array_equal(r["cus"][:, 0], ec["x"]) and np.array_equal(r["cus"][:, 1], ec["y"]) and \
                    np.array_equal(r["cus"][:, 2], ec["log2_cuw"]), poc
                assert e["pre"] is None or all(np.array_equal(a, b) for a, b in zip(r["rec"], e["pre"])), f"POC {poc}: unfiltered picture differs"
                assert all(np.array_equal(a, b) for a, b in zip(r["post"], e["post"])), f"POC {poc}: deblocked picture differs"
            self.results.append(r)
        self.pending = []
        return self.results

    def encode(self, pc):
        self.enqueue(pc)
        return self.collect()[-1]
