"""Frame work lists for the hot path: every call the inter analysis of ONE B picture makes,
laid out so that each stage is one frame-wide grid.

The reference walks a picture CTU by CTU and CU by CU (src_base/xeve_enc.c:103-175,
src_base/xeve_mode.c:2007-2374) and calls, for every CU of the 64/32/16/8 quad-tree,
pi->fn_me for list 0 and list 1, the bi-prediction search of analyze_bi (fn_mc -> get_org_bi ->
fn_me with bi = 1) and pinter_residue_rdo for the L0, L1 and BI candidates
(src_base/xeve_pinter.c:1839-2056).  This module enumerates exactly those calls for all CUs of a
picture at once.  What it cannot take from the reference's serial chain -- the motion-vector
predictor of each CU (it depends on the neighbours' final decisions) and the CABAC-derived RDOQ
rate tables -- is synthesised: MVPs are the clip's true motion plus a seeded jitter, rate tables
are those of a fresh CABAC state (every bin costs one bit).  DESIGN.md discusses the dependency.

Stages (each one C-ABI call = one or a few kernel launches over the whole frame):
    1. me_uni   : 2 items per CU (L0, L1)                       -> xb200_me
    2. bi_org   : 1 item per CU (prediction of the better list)  -> xb200_bi_org
    3. me_bi    : 1 item per CU (other list, bi = 1)             -> xb200_me
    4. residue  : 3 items per CU (L0, L1, BI)                    -> xb200_residue
"""
from __future__ import annotations

import math

import numpy as np

from . import api


def cu_grid(w, h):
    """All CUs of the 64/32/16/8 quad-tree that lie inside the picture: arrays (x, y, log2)."""
    xs, ys, ls = [], [], []
    for l2 in (6, 5, 4, 3):
        s = 1 << l2
        gx, gy = np.meshgrid(np.arange(0, w - s + 1, s), np.arange(0, h - s + 1, s))
        xs.append(gx.ravel()); ys.append(gy.ravel()); ls.append(np.full(gx.size, l2))
    return np.concatenate(xs).astype(np.int16), np.concatenate(ys).astype(np.int16), np.concatenate(ls).astype(np.uint8)


def fresh_rates():
    """RDOQ rate tables of an untrained CABAC state: p = 0.5 -> 32768 (1 bit) per bin
    (the reference derives them in xeve_rdoq_bit_est, src_base/xeve_mode.c:326-373)."""
    r = np.zeros(1, api.RATES)
    for f in ("cbf_all", "cbf_luma", "cbf_cb", "cbf_cr", "run", "level", "last"):
        r[f] = 32768
    return r


def lambdas(qp):
    """lambda[0] = 0.57 * 2^((qp-12)/3) (src_base/xeve_enc.c:1515); chroma lambdas scaled by the usual
    Baseline chroma weight at this QP (about 0.63); lambda_mv = floor(65536 * sqrt(lambda))."""
    lam = 0.57 * 2.0 ** ((qp - 12) / 3.0)
    return np.array([lam, lam * 0.63, lam * 0.63]), int(math.floor(65536.0 * math.sqrt(lam)))


class FrameWork:
    """Work lists of one B picture (poc) predicted from ref_pocs = (L0 poc, L1 poc)."""

    def __init__(self, w, h, poc, ref_pocs, pan, cur_pic, ref_pics, qp=38, me_range=32, gop_size=16, seed=0, rows=None):
        rng = np.random.default_rng(seed)
        x, y, l2 = cu_grid(w, h)
        if rows is not None:  # bounded sample: only CUs inside the first `rows` CTU rows
            keep = y + (1 << l2.astype(np.int32)) <= rows * 64
            x, y, l2 = x[keep], y[keep], l2[keep]
        n = len(x)
        self.n_cu, self.x, self.y, self.l2 = n, x, y, l2
        self.w, self.h, self.poc, self.ref_pocs = w, h, poc, ref_pocs
        self.cur_pic, self.ref_pics = cur_pic, ref_pics
        self.lam, self.lambda_mv = lambdas(qp)
        self.qp = np.array([qp + 12, qp + 10, qp + 10], np.uint8)  # core->qp_y/u/v incl. 6*(bd-8)
        self.rates = fresh_rates()
        # true motion of the panning background towards each reference, quarter-pel
        true_mv = [np.array([pan[0] * (poc - rp) * 4, pan[1] * (poc - rp) * 4]) for rp in ref_pocs]
        # ---- stage 1: uni-directional search, 2 items per CU (item 2*i + lidx) ----------------------
        me = np.zeros(2 * n, api.ME_ITEM)
        for l in (0, 1):
            s = me[l::2]
            s["poc"], s["cur_pic"], s["ref_pic"], s["ref_poc"] = poc, cur_pic, ref_pics[l], ref_pocs[l]
            s["x"], s["y"], s["log2_cuw"], s["log2_cuh"], s["lidx"] = x, y, l2, l2, l
            s["mvp"] = (true_mv[l] + rng.integers(-6, 7, (n, 2))).astype(np.int16)
        me["bi"], me["refi"], me["num_refp"] = 0, 0, 1
        me["lambda_mv"], me["max_search_range"], me["gop_size"], me["org_bi_off"] = self.lambda_mv, me_range, gop_size, -1
        self.me_uni = me
        sz = (1 << (2 * l2.astype(np.int64)))
        self.side_off = np.zeros(n, np.int64)
        np.cumsum(sz[:-1], out=self.side_off[1:])
        self.side_elems = int(sz.sum())
        self.res_off = np.zeros(3 * n, np.int64)
        rsz = np.repeat(sz * 3 // 2, 3)
        np.cumsum(rsz[:-1], out=self.res_off[1:])
        self.res_elems = int(rsz.sum())

    # ---- stages 2-4 are built from stage-1 results (host logic of analyze_bi / analyze_cu) ---------
    def build_bi(self, me_uni_out):
        """analyze_bi, first iteration (src_base/xeve_pinter.c:1588-1656): predict from the list with the
        smaller ME cost, search the other list against 2*org - pred starting at its uni-search MV."""
        n = self.n_cu
        c0, c1 = me_uni_out["cost"][0::2], me_uni_out["cost"][1::2]
        ref_l = np.where(c0 <= c1, 0, 1).astype(np.int8)  # lidx_ref
        cnd_l = 1 - ref_l
        idx = np.arange(n)
        mv = np.stack([me_uni_out["mv_out"][0::2], me_uni_out["mv_out"][1::2]], 1)  # [n, list, 2]
        mc = np.zeros(n, api.MC_ITEM)
        mc["poc"], mc["x"], mc["y"] = self.poc, self.x, self.y
        mc["w"] = mc["h"] = (1 << self.l2.astype(np.int16))
        for l in (0, 1):
            on = ref_l == l
            mc["refi"][:, l] = np.where(on, 0, -1)
            mc["ref_pic"][:, l] = np.where(on, self.ref_pics[l], -1)
            mc["ref_poc"][:, l] = np.where(on, self.ref_pocs[l], -1)
        mc["mv"] = mv
        self.bi_mc = mc
        self.bi_cur = np.full(n, self.cur_pic, np.int32)
        me = me_uni_out[2 * idx + cnd_l].copy()
        me["bi"] = 1
        me["mv_in"] = mv[idx, cnd_l]
        me["org_bi_off"] = self.side_off
        mb = np.stack([me_uni_out["mot_bits_out"][0::2, 0], me_uni_out["mot_bits_out"][1::2, 1]], 1)  # bits of each list
        me["mot_bits_in"] = mb
        self.me_bi = me
        self._ref_l, self._cnd_l, self._mv_uni = ref_l, cnd_l, mv
        return mc, me

    def build_residue(self, me_bi_out):
        """pinter_residue_rdo candidates L0, L1, BI of every CU (item 3*i + pidx)."""
        n = self.n_cu
        idx = np.arange(n)
        res = np.zeros(3 * n, api.RESIDUE_ITEM)
        mv_bi = self._mv_uni.copy()
        mv_bi[idx, self._cnd_l] = me_bi_out["mv_out"]
        for pidx in range(3):
            s = res[pidx::3]
            m = s["mc"]
            m["poc"], m["x"], m["y"] = self.poc, self.x, self.y
            m["w"] = m["h"] = (1 << self.l2.astype(np.int16))
            for l in (0, 1):
                on = pidx == 2 or pidx == l
                m["refi"][:, l] = 0 if on else -1
                m["ref_pic"][:, l] = self.ref_pics[l] if on else -1
                m["ref_poc"][:, l] = self.ref_pocs[l] if on else -1
            m["mv"] = mv_bi if pidx == 2 else self._mv_uni
            s["mc"] = m
        res["cur_pic"], res["slice_type"], res["run_stats"], res["qp"] = self.cur_pic, 0, 7, self.qp
        res["rate_idx"], res["lambda"], res["out_off"] = 0, self.lam, self.res_off
        self.residue = res
        return res

    # ---- the fused per-CU decision (xb200_analyze_cu): one item per CU of the quad-tree ---------------------------
    def build_cu(self, rates_of_state, max_search_range=32, seed=0):
        """xeve_pinter_analyze_cu inputs of every CU: MVP candidates per list = {stage-1 MVP, a jittered copy, (1,1) for
        an unavailable neighbour, the colocated (true-motion) vector} as xeve_get_motion orders them; fresh coder state
        (xeve_sbac_reset: range 16384, every model PROB_INIT) and the RDOQ tables derived from it by xb200_rdoq_rates."""
        rng = np.random.default_rng(seed + 77)
        n = self.n_cu
        cu = np.zeros(n, api.CU_ITEM)
        cu["poc"], cu["cur_pic"], cu["x"], cu["y"], cu["log2_cuw"], cu["log2_cuh"] = self.poc, self.cur_pic, self.x, self.y, self.l2, self.l2
        cu["slice_type"], cu["all_preds"], cu["num_refp"], cu["qp"], cu["max_search_range"] = 0, 1, 1, self.qp, max_search_range
        cu["ref_pic"], cu["ref_poc"] = -1, -1
        for l in (0, 1):
            cu["ref_pic"][:, l, 0], cu["ref_poc"][:, l, 0] = self.ref_pics[l], self.ref_pocs[l]
            m0 = self.me_uni["mvp"][l::2]
            cu["mvp"][:, l, 0] = m0
            cu["mvp"][:, l, 1] = m0 + rng.integers(-3, 4, (n, 2))
            cu["mvp"][:, l, 2] = 1
            cu["mvp"][:, l, 3] = m0 + rng.integers(-2, 3, (n, 2))
            cu["mv_dir"][:, l] = m0
        cu["ctx_skip"], cu["ctx_pred_mode"] = rng.integers(0, 2, n), rng.integers(0, 3, n)
        cu["lambda_mv"], cu["lambda"] = self.lambda_mv, self.lam
        cu["dist_chroma_weight"] = 1.0 / 0.63
        cu["rate_idx"], cu["state_in"], cu["state_out"] = 0, 0, 1 + np.arange(n)
        sz = (3 << (2 * self.l2.astype(np.int64))) >> 1
        cu["out_off"] = np.concatenate([[0], np.cumsum(sz)[:-1]])
        states = np.zeros(n + 1, api.SBAC)
        states["range"], states["m"] = 16384, 512
        self.cu, self.cu_states, self.cu_elems = cu, states, int(sz.sum())
        self.cu_rates = rates_of_state(states[:1])
        return cu

    # ---- bookkeeping for bench.py ------------------------------------------------------------------------
    def counts(self):
        return dict(cus=self.n_cu, me_uni=2 * self.n_cu, bi_org=self.n_cu, me_bi=self.n_cu, residue=3 * self.n_cu)


def synth_deblock(w, h, seed, intra_frac=0.1):
    """Random but well-formed deblocking input at any size: a random quad-tree per 64x64 CTU down to 4x4 (z-scan order, CUs
    clipped to the picture like the reference's implicit boundary splits), random per-CU intra / cbf / QP / motion, random
    samples.  Returns dict(pre=(Y, U, V), cus, pp, map_scu, map_refi, map_mv)."""
    rng = np.random.default_rng(seed)
    ws, hs = w // 4, h // 4
    cus = []

    def tree(x, y, l2):
        if x >= w or y >= h:
            return
        size = 1 << l2
        if x + size > w or y + size > h or (l2 > 2 and rng.random() < (0.9 if l2 > 4 else 0.45)):
            for dy in (0, size // 2):
                for dx in (0, size // 2):
                    tree(x + dx, y + dy, l2 - 1)
        else:
            cus.append((x, y, l2, l2, (0, 0)))
    for y in range(0, h, 64):
        for x in range(0, w, 64):
            tree(x, y, 6)
    cus = np.array(cus, api.DF_CU)
    map_scu = np.zeros((hs, ws), np.uint32)
    map_refi = np.zeros((hs, ws, 2), np.int8)
    map_mv = np.zeros((hs, ws, 2, 2), np.int16)
    n = len(cus)
    intra = rng.random(n) < intra_frac
    cbf = rng.random(n) < 0.4
    qp = rng.integers(22, 52, n)
    refi = rng.integers(-1, 2, (n, 2)).astype(np.int8)
    refi[(refi < 0).all(1), 0] = 0
    mv = rng.integers(-6, 7, (n, 2, 2)).astype(np.int16)
    for i, c in enumerate(cus):
        xs, ys, cw = int(c["x"]) >> 2, int(c["y"]) >> 2, (1 << int(c["log2_cuw"])) >> 2
        v = (int(qp[i]) << 16) | (int(intra[i]) << 15) | (int(cbf[i] and not intra[i]) << 24) | (1 << 31)
        map_scu[ys:ys + cw, xs:xs + cw] = v
        map_refi[ys:ys + cw, xs:xs + cw] = -1 if intra[i] else refi[i]
        map_mv[ys:ys + cw, xs:xs + cw] = 0 if intra[i] else mv[i]
    pp = np.zeros(1, api.DF_PIC)
    pp["w_scu"], pp["h_scu"], pp["qp_u_offset"], pp["qp_v_offset"] = ws, hs, int(rng.integers(-2, 3)), int(rng.integers(-2, 3))
    tab = np.arange(-12, 58)
    pp["chroma_qp"][0, 0] = np.where(tab < 30, tab, 30 + (tab - 30) * 3 // 4)   # a plausible monotone mapping
    pp["chroma_qp"][0, 1] = np.where(tab < 33, tab, 33 + (tab - 33) * 2 // 3)
    base = rng.integers(0, 1024, (h // 8 + 1, w // 8 + 1))
    yy = np.kron(base, np.ones((8, 8), np.int64))[:h, :w] + rng.integers(-20, 21, (h, w))
    pre = [np.clip(yy, 0, 1023).astype(np.int16), np.clip(yy[::2, ::2] + rng.integers(-9, 10, (h // 2, w // 2)), 0, 1023).astype(np.int16),
           rng.integers(0, 1024, (h // 2, w // 2)).astype(np.int16)]
    return dict(pre=pre, cus=cus, pp=pp[0], map_scu=map_scu.reshape(-1), map_refi=map_refi.reshape(-1, 2), map_mv=map_mv.reshape(-1, 2, 2))


def synth_intra(w, h, planes10, cur_pic, rates_of_state, qp=29, seed=0, sizes=(5, 4, 3, 2)):
    """pintra_analyze_cu inputs for every CU of the 32/16/8/4 quad-tree of an intra picture, all at once (frame-parallel
    mode): the reference samples are taken from the ORIGINAL picture -- what the reference's own
    look-ahead variant does (pintra_get_nbr_simple, src_base/xeve_pintra.c:463-490) -- with the left column, the upper row
    and the upper-right extension available inside the picture and the lower-left extension unavailable; fresh coder state.
    planes10: (Y, U, V) original picture at the internal bit depth.  Returns (items, states, rates, side, elems)."""
    rng = np.random.default_rng(seed + 5)
    lam, _ = lambdas(qp)
    items_all, side_all, pos = [], [], 0
    half = 512
    for l2 in sizes:
        s = 1 << l2
        gx, gy = np.meshgrid(np.arange(0, w - s + 1, s), np.arange(0, h - s + 1, s))
        xs, ys = gx.ravel().astype(np.int64), gy.ravel().astype(np.int64)
        n = len(xs)
        blocks = []
        for c, pl in enumerate(planes10):
            nn = s if c == 0 else s >> 1
            px, py = (xs, ys) if c == 0 else (xs >> 1, ys >> 1)
            P = 2 * nn + 2
            pad = np.pad(pl.astype(np.int16), P, constant_values=half)
            k = np.arange(-1, 2 * nn)
            left = pad[py[:, None] + k[None, :] + P, (px - 1)[:, None] + P].copy()
            left[:, nn + 1:] = half                                    # lower-left extension: not yet coded
            up = pad[(py - 1)[:, None] + P, px[:, None] + k[None, :] + P].copy()
            left[:, 0] = up[:, 0]
            blocks += [left, up]
        side_all.append(np.concatenate(blocks, axis=1).reshape(-1))
        it = np.zeros(n, api.INTRA_ITEM)
        it["cur_pic"], it["x"], it["y"], it["log2_cuw"], it["log2_cuh"] = cur_pic, xs, ys, l2, l2
        it["nb_off"] = pos + np.arange(n) * (8 * s + 6)
        pos += n * (8 * s + 6)
        items_all.append(it)
    items = np.concatenate(items_all)
    n = len(items)
    items["slice_type"], items["all_preds"], items["qp"] = 2, 1, np.array([qp + 12, qp + 10, qp + 10], np.uint8)
    items["mpm"] = np.array([0, 2, 3, 1, 4], np.uint8)                 # xeve_tbl_mpm[DC][DC]
    items["inter_satd"] = 0xFFFFFFFF
    items["lambda"], items["sqrt_lambda0"], items["dist_chroma_weight"] = lam, math.sqrt(lam[0]), 1.0 / 0.63
    items["cm_ipm_in"] = 512
    items["rate_idx"], items["state_in"], items["state_out"] = 0, 0, 1 + np.arange(n)
    sz = (3 << (2 * items["log2_cuw"].astype(np.int64))) >> 1
    items["out_off"] = np.concatenate([[0], np.cumsum(sz)[:-1]])
    states = np.zeros(n + 1, api.SBAC)
    states["range"], states["m"] = 16384, 512
    return items, states, rates_of_state(states[:1]), np.concatenate(side_all), int(sz.sum())
