// xb200_lanecoder.cuh -- the cbf decisions of pinter_residue_rdo with one CODER PER LANE.
//
// k_cu_decide (xb200_pipeline.cuh) runs the serial CABAC bit counter on one lane of a warp per candidate: 31 lanes idle,
// and the serial, branchy bin loop issues one instruction every ~10 cycles.  Here every lane owns a candidate.  To keep the
// 32 coders of a warp convergent the work is expressed as a bin-level state machine -- every loop iteration each lane
// produces its next (context, bin) or bypass run from {header script, cbf script, non-zero coefficient list} and performs
// one engine step on models held in shared memory as [copy][context][lane] (bank-conflict free).  Only the 28 context
// models the inter syntax touches are carried (skip, pred_mode, direct, inter_dir, refi, mvp_idx, mvd, cbf x4,
// run 0-3, last 0-1, level 0-3); the other 40 pass through unchanged.
// The coefficient planes are first compacted to (zero-run, level) lists in zig-zag order by k_cu_nzlist (warp per
// candidate, ballot compaction), so a lane reads its symbols sequentially.
#pragma once
#include "xb200_pipeline.cuh"

constexpr int LC_NCTX = 28, LC_HDR = 32, LC_CBF = 4;
enum { LC_IN = 0, LC_RUN = 1, LC_MODE = 2, LC_CPREV = 3, LC_CRUN = 4, LC_COPIES = 5 };
XB_DEV int lc_ctx(int idx) // XB200_CM_* index -> compact index
{
    return idx < XB200_CM_RUN ? idx : (idx < XB200_CM_LAST ? 18 + (idx - XB200_CM_RUN) : (idx < XB200_CM_LEVEL ? 22 + (idx - XB200_CM_LAST) : 24 + (idx - XB200_CM_LEVEL)));
}
XB_DEV int lc_full(int k) // compact index -> XB200_CM_* index
{
    return k < 18 ? k : (k < 22 ? XB200_CM_RUN + (k - 18) : (k < 24 ? XB200_CM_LAST + (k - 22) : XB200_CM_LEVEL + (k - 24)));
}

// (zero-run, level) lists of every coded plane of every residue slot; meta bit c: the last non-zero of plane c sits at the
// final scan position (no `last` flag follows it).  One warp per slot.
// the list of the plane stored at coef[off] lives at nz[lc_nzoff(off)]: the coefficient planes of mode m start at 3 * m * elems
XB_DEV int64_t lc_nzoff(int64_t off, int64_t elems) { return (off / (3 * elems)) * elems + off % (3 * elems); }

__global__ void __launch_bounds__(128) k_cu_nzlist(const xb200_residue_item *__restrict__ res, int n_slots, const int16_t *__restrict__ coef,
                                                   uint16_t *__restrict__ nz_run, int16_t *__restrict__ nz_lev, uint8_t *__restrict__ meta, int64_t elems, int w_lo,
                                                   int w_hi)
{
    const int lane = threadIdx.x & 31, sl = blockIdx.x * 4 + (threadIdx.x >> 5);
    if(sl >= n_slots) return;
    const xb200_residue_item &it = res[sl];
    if(it.mc.w < w_lo || it.mc.w > w_hi) return;
    int l2 = 0;
    while((1 << l2) < it.mc.w) l2++;
    const int ny = 1 << (2 * l2);
    int       m = 0;
    for(int c = 0; c < 3; c++) {
        if(it.nnz[c] == 0) continue;
        const int       l = c ? l2 - 1 : l2, n = 1 << (2 * l);
        const int64_t   base = it.out_off + (c == 0 ? 0 : (c == 1 ? ny : ny + (ny >> 2))), nb = lc_nzoff(base, elems);
        const uint16_t *scan = g_scan + scan_base(l);
        int             cnt = 0, carry = 0;
        for(int b0 = 0; b0 < n; b0 += 32) {
            const int      sp = b0 + lane;
            const int      v = sp < n ? coef[base + scan[sp]] : 0;
            const unsigned mask = __ballot_sync(0xffffffffu, v != 0);
            if(v != 0) {
                const unsigned lower = mask & ((1u << lane) - 1u);
                const int      run = lower ? lane - (31 - __clz(lower)) - 1 : lane + carry;
                const int      k = cnt + __popc(lower);
                nz_run[nb + k] = (uint16_t)run;
                nz_lev[nb + k] = (int16_t)v;
                if(sp == n - 1) m |= 1 << c;
            }
            cnt += __popc(mask);
            carry = mask ? __clz(mask) : carry + 32; // zeros after the last non-zero of this chunk (bit 31 = lane 31)
        }
    }
    m = __reduce_or_sync(0xffffffffu, m);
    if(lane == 0) meta[sl] = (uint8_t)m;
}

struct LcShared {
    uint16_t m[LC_COPIES][LC_NCTX][32];
    uint16_t hdr[LC_HDR][32];   // script entries: bit 15 = bypass run (low bits = count), else ctx << 1 | bin
    uint16_t cbf[LC_CBF][32];
};

// one lane = one residue candidate: all cbf combination counts + costs; per_cu / pidx mapping as k_cu_decide
__global__ void __launch_bounds__(128) k_cu_decide_lanes(const xb200_cu_item *__restrict__ items, int n_slots, int per_cu,
                                                         const xb200_sbac *__restrict__ st_in, CuState *__restrict__ states,
                                                         const xb200_residue_item *__restrict__ res, const uint16_t *__restrict__ nz_run,
                                                         const int16_t *__restrict__ nz_lev, const uint8_t *__restrict__ meta, int64_t elems, int w_lo, int w_hi)
{
    extern __shared__ __align__(16) unsigned char lc_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    LcShared &L = reinterpret_cast<LcShared *>(lc_smem)[w];
    const int sl = blockIdx.x * blockDim.x + threadIdx.x;
    bool      live = sl < n_slots;
    const xb200_residue_item *it = live ? &res[sl] : nullptr;
    if(live && (it->mc.w < w_lo || it->mc.w > w_hi)) live = false;
    if(!__any_sync(0xffffffffu, live)) return;
    // ---- per-lane inputs -----------------------------------------------------------------------------------------
    int      pidx = 0, ci = 0, l2 = 3, store[3] = {0, 0, 0}, nh = 0;
    int64_t  d0[3] = {0, 0, 0}, d1[3] = {0, 0, 0}, base = 0;
    double   lam[3] = {0, 0, 0}, w0 = 0, w1 = 0;
    uint32_t range_in = 16384;
    int      at_end = 0;
    if(live) {
        ci = sl / per_cu;
        const int k = sl - ci * per_cu;
        pidx = per_cu == 1 ? 2 : (k == 0 ? 4 : k - 1);
        const xb200_cu_item &cu = items[ci];
        const CuMode        &M = states[ci].md[pidx];
        const xb200_sbac    &s = st_in[cu.state_in];
        range_in = s.range;
        for(int q = 0; q < LC_NCTX; q++) L.m[LC_IN][q][lane] = s.m[lc_full(q)];
        l2 = cu.log2_cuw;
        for(int c = 0; c < 3; c++) { store[c] = it->nnz[c]; d0[c] = it->dist_pred[c]; d1[c] = it->dist_rec[c]; lam[c] = cu.lambda[c]; }
        w0 = cu.dist_chroma_weight[0]; w1 = cu.dist_chroma_weight[1];
        base = lc_nzoff(it->out_off, elems);
        at_end = meta[sl];
        // header script of xeve_rdo_bit_cnt_cu_inter (src_base/xeve_mode.c:185-281) up to the coefficients
        auto put = [&](int ctx, int bin) { L.hdr[nh++][lane] = (uint16_t)((lc_ctx(ctx) << 1) | (bin ? 1 : 0)); };
        auto byp = [&](int cnt) { if(cnt > 0) L.hdr[nh++][lane] = (uint16_t)(0x8000 | cnt); };
        const bool B = cu.slice_type == 0;
        if(cu.slice_type != 2) {
            put(XB200_CM_SKIP_FLAG + cu.ctx_skip, 0);
            if(cu.all_preds) put(XB200_CM_PRED_MODE + cu.ctx_pred_mode, 0);
            put(XB200_CM_DIRECT, pidx == 4);
            if(pidx != 4) {
                if(M.refi[0] >= 0 && M.refi[1] >= 0) put(XB200_CM_INTER_DIR, 0);
                else {
                    if(B) put(XB200_CM_INTER_DIR, 1);
                    put(XB200_CM_INTER_DIR + 1, M.refi[0] < 0);
                }
                for(int l = 0; l < 2; l++) {
                    if(M.refi[l] < 0 || (l == 1 && !B)) continue;
                    const int nr = cu.num_refp[l], r = M.refi[l];
                    if(nr > 1) {   // xeve_eco_refi
                        put(XB200_CM_REFI, r != 0);
                        if(r != 0) {
                            int nb = 0;
                            for(int i = 2; i < nr; i++) {
                                const int bin = i != r + 1;
                                if(i == 2) put(XB200_CM_REFI + 1, bin); else nb++;
                                if(!bin) break;
                            }
                            byp(nb);
                        }
                    }
                    for(int i = 0; i < 3; i++) {   // mvp_idx, truncated unary
                        put(XB200_CM_MVP_IDX + i, i != M.mvp_idx[l]);
                        if(i == M.mvp_idx[l]) break;
                    }
                    for(int q = 0; q < 2; q++) {   // mvd: exp-golomb, two context bins then bypass (+ sign)
                        const uint32_t a = (uint32_t)abs((int)M.mvd[l][q]);
                        const int      len_i = 32 - __clz((a + 1) >> 1), len_c = 2 * len_i + 1;
                        const uint32_t code = (1u << len_i) | ((a + 1 - (1u << len_i)) & ((1u << len_i) - 1));
                        put(XB200_CM_MVD, (code >> (len_c - 1)) & 1);
                        if(len_c > 1) put(XB200_CM_MVD, (code >> (len_c - 2)) & 1);
                        byp(max(0, len_c - 2) + (a != 0));
                    }
                }
            }
        }
    }
    __syncwarp();
    const int tnnz = store[0] + store[1] + store[2];
    // ---- phases ------------------------------------------------------------------------------------------------------
    // 0: all-zero (or the only count when nothing was coded), 1: as-is, 2..7: plane p = (ph-2)/2 with j = (ph-2)&1, 8: idx_best mix
    double   best = CU_MAX_COST, comp_best = CU_MAX_COST;
    int      cbf = 0, idx_best[3] = {0, 0, 0};
    uint32_t rg[LC_COPIES];
    rg[LC_IN] = range_in;
    if(live) {
        for(int q = 0; q < LC_NCTX; q++) L.m[LC_CPREV][q][lane] = L.m[LC_IN][q][lane];
        rg[LC_CPREV] = range_in;
    }
    for(int ph = 0; ph < 9; ph++) {
        const int p = (ph - 2) >> 1, j = (ph - 2) & 1;
        bool      on = live;
        int       n0 = 0, n1 = 0, n2 = 0, kind = 1, src = LC_IN;
        if(ph == 0) on = on && (tnnz == 0 || pidx != 4);
        else if(ph == 1) { on = on && tnnz != 0; n0 = store[0] > 0; n1 = store[1] > 0; n2 = store[2] > 0; }
        else if(ph < 8) {
            on = on && tnnz != 0 && store[p] > 0;
            kind = 3;
            n0 = p == 0 ? j : 1; n1 = p == 1 ? j : 1; n2 = p == 2 ? j : 1;   // other planes at their stored values
            src = LC_CRUN;
            if(on && j == 0) {   // SBAC_LOAD(s_temp_prev_comp_run, s_temp_prev_comp_best)
                for(int q = 0; q < LC_NCTX; q++) L.m[LC_CRUN][q][lane] = L.m[LC_CPREV][q][lane];
                rg[LC_CRUN] = rg[LC_CPREV];
                comp_best = CU_MAX_COST;
            }
        }
        else {
            const bool any = idx_best[0] || idx_best[1] || idx_best[2];
            const bool differs = (idx_best[0] ? store[0] : 0) != store[0] || (idx_best[1] ? store[1] : 0) != store[1] ||
                                 (idx_best[2] ? store[2] : 0) != store[2];
            on = on && tnnz != 0 && any && differs;
            n0 = idx_best[0]; n1 = idx_best[1]; n2 = idx_best[2];
        }
        if(!__any_sync(0xffffffffu, on)) continue;
        const int nn[3] = {n0 ? store[0] : 0, n1 ? store[1] : 0, n2 ? store[2] : 0};
        // cbf script of xeve_eco_cbf for this count (src_base/xeve_eco.c:793-893)
        int  nc = 0;
        bool planes = true;
        if(on) {
            auto putc = [&](int ctx, int bin) { L.cbf[nc++][lane] = (uint16_t)((lc_ctx(ctx) << 1) | (bin ? 1 : 0)); };
            const int f0 = nn[0] != 0, f1 = nn[1] != 0, f2 = nn[2] != 0;
            if(kind == 1) {
                const int all = f0 + f1 + f2;
                putc(XB200_CM_CBF_ALL, all != 0);
                if(!all) planes = false;
                else {
                    putc(XB200_CM_CBF_CB, f1); putc(XB200_CM_CBF_CR, f2);
                    if(f1 + f2 != 0) putc(XB200_CM_CBF_LUMA, f0);
                }
            }
            else {
                if(p == 1) putc(XB200_CM_CBF_CB, f1);
                if(p == 2) putc(XB200_CM_CBF_CR, f2);
                if(p == 0 && (f1 + f2 != 0)) putc(XB200_CM_CBF_LUMA, f0);
            }
            for(int q = 0; q < LC_NCTX; q++) L.m[LC_RUN][q][lane] = L.m[src][q][lane];
        }
        uint32_t range = rg[src], bits = 0;
        // generator state
        int  stage = kind == 1 ? 0 : 1, pos = 0, rem = 0, tail_ctx = 0;
        int  plane = -1, sym = 0, sub = 0, cnt = 0, run_v = 0, lev_v = 0;
        int64_t pbase = 0;
        bool done = !on;
        const int run_mask = kind == 1 ? 7 : (1 << p);
        while(!__all_sync(0xffffffffu, done)) {
            int ctx = -1, bin = 0, nbyp = 0;
            if(!done) {
                if(rem > 0) { ctx = tail_ctx; bin = rem > 1; rem--; }
                else if(stage == 0) {
                    if(pos < nh) {
                        const uint16_t e = L.hdr[pos++][lane];
                        if(e & 0x8000) nbyp = e & 0x7fff; else { ctx = e >> 1; bin = e & 1; }
                    }
                    else { stage = 1; pos = 0; }
                }
                else if(stage == 1) {
                    if(pos < nc) { const uint16_t e = L.cbf[pos++][lane]; ctx = e >> 1; bin = e & 1; }
                    else if(!planes) done = true;
                    else { stage = 2; plane = -1; sub = 4; }
                }
                else {
                    if(sub == 4) {   // advance to the next coded plane
                        plane++;
                        while(plane < 3 && !(nn[plane] != 0 && ((run_mask >> plane) & 1))) plane++;
                        if(plane >= 3) done = true;
                        else {
                            const int ny = 1 << (2 * l2);
                            pbase = base + (plane == 0 ? 0 : (plane == 1 ? ny : ny + (ny >> 2)));
                            cnt = nn[plane]; sym = 0; sub = 0;
                        }
                    }
                    else if(sub == 0) {   // unary(run) on RUN + t0, RUN + t0 + 1
                        run_v = nz_run[pbase + sym]; lev_v = nz_lev[pbase + sym];
                        const int t0 = plane == 0 ? 0 : 2;
                        ctx = 18 + t0; bin = run_v != 0; rem = run_v; tail_ctx = 18 + t0 + 1;
                        sub = 1;
                    }
                    else if(sub == 1) {   // unary(|level| - 1) on LEVEL + t0, LEVEL + t0 + 1
                        const int t0 = plane == 0 ? 0 : 2, lv = abs(lev_v) - 1;
                        ctx = 24 + t0; bin = lv != 0; rem = lv; tail_ctx = 24 + t0 + 1;
                        sub = 2;
                    }
                    else if(sub == 2) { nbyp = 1; sub = 3; }   // sign
                    else {   // last flag, unless the symbol sits at the final scan position
                        const bool final_sym = sym == cnt - 1;
                        if(final_sym && ((at_end >> plane) & 1)) sub = 4;
                        else {
                            ctx = 22 + (plane != 0); bin = final_sym;
                            sub = final_sym ? 4 : 0;
                        }
                        sym++;
                    }
                }
            }
            if(ctx >= 0) {
                const uint32_t model = L.m[LC_RUN][ctx][lane];
                uint32_t       mps = model & 1, state = model >> 1;
                cb_step(range, bits, state, mps, (uint32_t)bin);
                L.m[LC_RUN][ctx][lane] = (uint16_t)((state << 1) | mps);
            }
            else if(nbyp) { range &= ~1u; bits += nbyp; }
        }
        // ---- cost of this count and bookkeeping of pinter_residue_rdo -------------------------------------------------
        if(on) {
            double cost;
            if(kind == 1) {
                if(tnnz == 0) cost = __dadd_rn(__dadd_rn(__ll2double_rn(d0[0]), __dmul_rn(w0, __ll2double_rn(d0[1]))), __dmul_rn(w1, __ll2double_rn(d0[2])));
                else cost = __dadd_rn(__ll2double_rn(n0 ? d1[0] : d0[0]),
                                      __dadd_rn(__dmul_rn(__ll2double_rn(n1 ? d1[1] : d0[1]), w0), __dmul_rn(__ll2double_rn(n2 ? d1[2] : d0[2]), w1)));
                cost = __dadd_rn(cost, __dmul_rn((double)bits, lam[0]));
                if(cost < best) {
                    best = cost; cbf = (n0 ? 1 : 0) | (n1 ? 2 : 0) | (n2 ? 4 : 0);
                    for(int q = 0; q < LC_NCTX; q++) L.m[LC_MODE][q][lane] = L.m[LC_RUN][q][lane];
                    rg[LC_MODE] = range;
                }
            }
            else {
                const int64_t dd = j ? d1[p] : d0[p];
                cost = p == 0 ? __ll2double_rn(dd) : __dmul_rn(__ll2double_rn(dd), p == 1 ? w0 : w1);
                cost = __dadd_rn(cost, __dmul_rn((double)bits, lam[p]));
                if(cost < comp_best) {
                    comp_best = cost; idx_best[p] = j;
                    for(int q = 0; q < LC_NCTX; q++) L.m[LC_CPREV][q][lane] = L.m[LC_RUN][q][lane];
                    rg[LC_CPREV] = range;
                }
            }
        }
    }
    if(live) {
        CuState &S = states[ci];
        const xb200_sbac &s = st_in[items[ci].state_in];
        xb200_sbac       &o = S.st[pidx];
        for(int q = 0; q < XB200_CM_COUNT; q++) o.m[q] = s.m[q];
        for(int q = 0; q < LC_NCTX; q++) o.m[lc_full(q)] = L.m[LC_MODE][q][lane];
        o.range = rg[LC_MODE];
        S.cost[pidx] = best;
        S.md[pidx].cbf = cbf;
        S.md[pidx].nnz[0] = (cbf & 1) ? store[0] : 0; S.md[pidx].nnz[1] = (cbf & 2) ? store[1] : 0; S.md[pidx].nnz[2] = (cbf & 4) ? store[2] : 0;
    }
}
