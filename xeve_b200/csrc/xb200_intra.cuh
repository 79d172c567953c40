// xb200_intra.cuh -- intra analysis of a list of CUs.
//
// Reference: pintra_analyze_cu / make_ipred_list / pintra_residue_rdo (src_base/xeve_pintra.c:69-374, 544-698), the five
// Baseline predictors (src_base/xeve_ipred.c:99-228) and the intra bit counters (src_base/xeve_mode.c:81-171) over the
// CABAC syntax of src_base/xeve_eco.c:793-905 (cbf), 1104-1121 (intra_dir), run-length coefficients.
//
// One kernel instance per CU size (all block dimensions compile-time).  A TEAM owns a CU from the mode ranking to the final
// bit count: one warp for CUs up to 32x32 (four teams per CTA, warp-synchronous, no block barriers: the serial coder sections of
// one team overlap the transforms of the others), 128 threads for 64x64 -- the team transforms / parallel RDOQ of xb200_residue2.cuh.  Everything stays in shared memory: original
// block, reference samples, the quantised levels of the current / best candidate in zig-zag order for the coder.
// Predictions are never stored: a predictor is two or three shared-memory reads, so each use recomputes it.  The SATD of
// the five modes is evaluated in parallel, one (mode, 8x8 tile) pair per thread.  CTAs are persistent (the DCT matrix is
// staged once per CTA).
// Costs are IEEE doubles evaluated in the reference's operation order with explicit _rn operations (no FMA contraction).
#pragma once
#define XB200_DEVICE_FUNCS_ONLY
#include "xb200_cabac.cuh"
#include "xb200_had.cuh"
#include "xb200_residue2.cuh"

#define IN_CM_IPM XB200_CM_COUNT // the two ctx.intra_dir models follow the inter models in the shared-memory copy
#define IN_CM_N   (XB200_CM_COUNT + 2)
#define IN_MAX_COST 1.7e+308

// zig-zag position of raster element e of an N x N block (same closed form as quant_block, xb200_tq.cuh)
XB_DEV int zz_of(int e, int l2)
{
    const int N = 1 << l2, n = N * N, x = e & (N - 1), y = e >> l2, d = x + y;
    const int before = d < N ? (d * (d + 1)) >> 1 : n - (((2 * N - 1 - d) * (2 * N - d)) >> 1);
    const int mx = min(d, N - 1);
    return before + ((d & 1) ? mx - x : mx - y);
}

// src_base/xeve_ipred.c:99-228; le / up point at sample 0 of the left column / upper row (index -1 = corner)
XB_DEV int ipred_at(const int16_t *le, const int16_t *up, int ipm, int i, int j, int dc)
{
    switch(ipm) {
    case 0: return dc;
    case 1: return le[i];
    case 2: return up[j];
    case 3: return i > j ? le[i - j - 1] : (i == j ? up[-1] : up[j - i - 1]);
    default: return (up[i + j + 1] + le[i + j + 1]) >> 1;
    }
}
XB_DEV int ipred_dc(const int16_t *le, const int16_t *up, int l2)
{
    const int w = 1 << l2;
    int dc = 0;
    for(int k = 0; k < w; k++) dc += le[k] + up[k];
    return (dc + w) >> (l2 + 1);
}

// xeve_eco_run_length_cc over levels held in zig-zag order in shared memory; warp-cooperative (all 32 lanes of warp 0
// call), engine state valid on lane 0
XB_DEV void cb_run_length_sm(Cabac &c, const int16_t *lev, int n, int num_sig, int ch, int lane)
{
    const int t0 = ch == 0 ? 0 : 2;
    uint32_t  run = 0;
    bool      done = false;
    for(int base = 0; base < n && !done; base += 32) {
        const int sp = base + lane;
        const int v = sp < n ? lev[sp] : 0;
        uint32_t  nzm = __ballot_sync(0xffffffffu, v != 0);
        int       prev = -1;
        while(nzm && !done) {
            const int b = __ffs(nzm) - 1;
            nzm &= nzm - 1;
            const int cv = __shfl_sync(0xffffffffu, v, b);
            if(lane == 0) {
                run += b - prev - 1;
                cb_unary(c, run, XB200_CM_RUN + t0);
                cb_unary(c, (uint32_t)abs(cv) - 1, XB200_CM_LEVEL + t0);
                cb_ep(c); // sign
            }
            prev = b;
            run = 0;
            if(base + b == n - 1) { done = true; break; }
            num_sig--;
            if(lane == 0) cb_bin(c, XB200_CM_LAST + (ch != 0), num_sig == 0);
            if(num_sig == 0) done = true;
        }
        run += 31 - prev;
    }
}


constexpr int I4_THREADS = 128;
struct TCabac { // engine of xb200_cabac.cuh with strided models: model idx of this thread / lane at m[idx * CS]
    uint32_t  range, bits;
    uint16_t *m;
};
template <int CS = I4_THREADS> XB_DEV void tc_bin(TCabac &c, int idx, int bin)
{
    const uint32_t model = c.m[idx * CS];
    uint32_t       mps = model & 1, state = model >> 1;
    cb_step(c.range, c.bits, state, mps, (uint32_t)(bin != 0));
    c.m[idx * CS] = (uint16_t)((state << 1) | mps);
}
template <int CS = I4_THREADS> XB_DEV void tc_unary(TCabac &c, uint32_t sym, int idx) // sbac_write_unary_sym with two contexts
{
    tc_bin<CS>(c, idx, sym != 0);
    if(sym == 0) return;
    const uint32_t model = c.m[(idx + 1) * CS];
    uint32_t       mps = model & 1, state = model >> 1;
    for(; sym > 1; sym--) cb_step(c.range, c.bits, state, mps, 1);
    cb_step(c.range, c.bits, state, mps, 0);
    c.m[(idx + 1) * CS] = (uint16_t)((state << 1) | mps);
}
XB_DEV void tc_ep(TCabac &c) { c.range &= ~1u; c.bits++; }


// Hadamard SATD tile of (org - predictor) with the predictor evaluated on the fly (src_base/xeve_sad.c:417-607)
template <int TN, class F> XB_DEV int had_tile_fn(const int16_t *a, int sa, F pred)
{
    int m[TN][TN];
#pragma unroll
    for(int y = 0; y < TN; y++)
#pragma unroll
        for(int x = 0; x < TN; x++) m[y][x] = (int)a[y * sa + x] - pred(y, x);
#pragma unroll
    for(int pass = 0; pass < 2; pass++)
#pragma unroll
        for(int r = 0; r < TN; r++)
#pragma unroll
            for(int len = 1; len < TN; len <<= 1)
#pragma unroll
                for(int i = 0; i < TN; i += len << 1)
#pragma unroll
                    for(int jj = i; jj < i + len; jj++) {
                        int &p = pass ? m[jj][r] : m[r][jj], &q = pass ? m[jj + len][r] : m[r][jj + len];
                        const int u = p + q, v = p - q;
                        p = u; q = v;
                    }
    int s = abs(m[0][0]) >> 2;
#pragma unroll
    for(int y = 0; y < TN; y++)
#pragma unroll
        for(int x = 0; x < TN; x++)
            if(x | y) s += abs(m[y][x]);
    return TN == 8 ? (s + 2) >> 2 : (s + 1) >> 1;
}

template <int L2> struct IntraCfg {
    static constexpr int T     = L2 <= 5 ? 32 : 128;   // (a 256-thread team for 64x64 showed a shared-memory hazard under racecheck)
    static constexpr int TEAMS = L2 <= 5 ? 4 : 1;
    static constexpr int CTA   = T * TEAMS;
    static constexpr int N     = 1 << L2, NY = N * N, NCH = NY / 4;
    static constexpr int TILES = L2 == 2 ? 1 : (N / 8) * (N / 8);   // SATD tiles per mode (4x4 CU: one 4x4 tile)
};

// XB200_INTRA64_CAND_GLOBAL (decision-chain kernel): the five candidate level arrays of a 64x64 CU (41 KB, the largest item of any
// working set) live in global memory handed in by the caller instead of shared memory, so that more chains fit on an SM.
#ifdef XB200_INTRA64_CAND_GLOBAL
template <int L2> constexpr bool IN_CAND_SHARED = L2 < 6;
#else
template <int L2> constexpr bool IN_CAND_SHARED = true;
#endif
template <int L2> struct IntraTeam {
    static constexpr int N = 1 << L2, NY = N * N;
    int32_t  TB[NY < 32 ? 32 : NY];        // DCT stage buffer / RDOQ scratch
    int16_t  org[NY * 3 / 2];              // Y | U | V original block
    int16_t  blk[NY];                      // transform working block
    int16_t  candS[IN_CAND_SHARED<L2> ? 5 : 1][IN_CAND_SHARED<L2> ? NY + 2 : 2]; // luma levels of every candidate mode, zig-zag order (+2: lanes on different banks)
    uint16_t cm_lane[IN_CM_N * 8];         // one model set per candidate lane, [model][lane]
    int64_t  cand_ssd[5];
    int32_t  cand_nnz[5];
    uint32_t cand_bits[5];
    int16_t  chS[NY / 2];                  // chroma levels U | V, zig-zag order
    int16_t  nb[8 * N + 8];                // per plane: left[-1 .. 2n-1], up[-1 .. 2n-1]
    uint16_t cm_base[IN_CM_N + 2], cm_run[IN_CM_N + 2];
    int32_t  satd_part[5 * IntraCfg<L2>::TILES];
    uint32_t bits, range_run;
    int32_t  list[5], pred_cnt;
    TeamScratch X;
};

// pintra_analyze_cu of ONE CU by one team (one warp up to 32x32, the first 128 threads of the CTA for 64x64)
template <int L2>
__device__ __noinline__ void intra_cu_one(IntraTeam<L2> &M, const int8_t *tm, const int8_t *tmT, const PicDev *__restrict__ pics,
                                          xb200_intra_item &it, const xb200_rates *rates, const xb200_sbac *st_in, xb200_sbac *st_out,
                                          const int16_t *side, int16_t *coef, int16_t *rec, const SeqDev &sq, int tt,
                                          int16_t *cand_global = nullptr)
{
    using Cf = IntraCfg<L2>;
    constexpr int T = Cf::T, N = Cf::N, NY = Cf::NY, NC = N / 2, NCH = Cf::NCH, TILES = Cf::TILES, LC = L2 - 1;
    // candidate level arrays: [5][NY + 2] in the team's shared memory, or in the caller's global buffer (see IN_CAND_SHARED)
    int16_t (*const candS)[NY + 2] = IN_CAND_SHARED<L2> ? reinterpret_cast<int16_t (*)[NY + 2]>(&M.candS[0][0])
                                                         : reinterpret_cast<int16_t (*)[NY + 2]>(cand_global);
    const int lane = tt & 31;
    const int bd = sq.bd, maxv = (1 << bd) - 1, sh = (bd - 8) << 1;
    const int slice_type = it.slice_type, all_preds = it.all_preds, ctx_skip = it.ctx_skip, ctx_pm = it.ctx_pred_mode;
    const xb200_rates *rt = &rates[it.rate_idx];
    const double lambda0 = it.lambda[0];
    uint32_t range_base;
    uint8_t  mpm[5];
    // ---- stage inputs: original block, reference samples, coder state ------------------------------------------
    {
        const PicDev   p = pics[it.cur_pic];
        const int      x0 = it.x, y0 = it.y;
        const int16_t *gy = p.p[0] + (ptrdiff_t)y0 * p.s[0] + x0;
        for(int e = tt; e < NY; e += T) M.org[e] = gy[(ptrdiff_t)(e >> L2) * p.s[0] + (e & (N - 1))];
#pragma unroll
        for(int c = 1; c < 3; c++) {
            const int16_t *gc = p.p[c] + (ptrdiff_t)(y0 >> 1) * p.s[c] + (x0 >> 1);
            for(int e = tt; e < NCH; e += T) M.org[NY + (c - 1) * NCH + e] = gc[(ptrdiff_t)(e >> LC) * p.s[c] + (e & (NC - 1))];
        }
        const int16_t *gn = side + it.nb_off;
        for(int e = tt; e < 8 * N + 6; e += T) M.nb[e] = gn[e];
        const xb200_sbac &s0 = st_in[it.state_in];
        for(int k = tt; k < XB200_CM_COUNT; k += T) M.cm_base[k] = s0.m[k];
        if(tt == 0) { M.cm_base[IN_CM_IPM] = it.cm_ipm_in[0]; M.cm_base[IN_CM_IPM + 1] = it.cm_ipm_in[1]; }
        range_base = s0.range;
#pragma unroll
        for(int k = 0; k < 5; k++) mpm[k] = it.mpm[k];
    }
    team_sync<T>();
    const int16_t *leY = M.nb + 1, *upY = M.nb + (2 * N + 1) + 1;
    const int16_t *leC[2] = {M.nb + 2 * (2 * N + 1) + 1, M.nb + 2 * (2 * N + 1) + 2 * (N + 1) + 1};
    const int16_t *upC[2] = {leC[0] + (N + 1), leC[1] + (N + 1)};
    const int dcY = ipred_dc(leY, upY, L2);

    // ---- make_ipred_list (src_base/xeve_pintra.c:308-374): SATD of the five modes, one (mode, tile) pair per thread ----
    for(int w = tt; w < 5 * TILES; w += T) {
        const int ipm = w / TILES, t = w % TILES;
        if(L2 == 2) M.satd_part[w] = had_tile_fn<4>(M.org, N, [&](int y, int x) { return ipred_at(leY, upY, ipm, y, x, dcY); });
        else {
            constexpr int TW = N >= 8 ? N / 8 : 1;
            const int ty = (t / TW) * 8, tx = (t % TW) * 8;
            M.satd_part[w] = had_tile_fn<8>(M.org + ty * N + tx, N, [&](int y, int x) { return ipred_at(leY, upY, ipm, ty + y, tx + x, dcY); });
        }
    }
    team_sync<T>();
    if(tt == 0) {
        double   cand_cost[5];
        uint32_t cand_satd[5];
        int      list[5];
#pragma unroll
        for(int k = 0; k < 5; k++) { list[k] = 0; cand_cost[k] = IN_MAX_COST; cand_satd[k] = 0xffffffffu; }
        for(int ipm = 0; ipm < 5; ipm++) {
            int sum = 0;
            for(int t = 0; t < TILES; t++) sum += M.satd_part[ipm * TILES + t];
            const uint32_t satd = (uint32_t)(sum >> (bd - 8));
            Cabac c;
            c.range = range_base; c.bits = 0; c.m = M.cm_run;
            M.cm_run[IN_CM_IPM] = M.cm_base[IN_CM_IPM]; M.cm_run[IN_CM_IPM + 1] = M.cm_base[IN_CM_IPM + 1];
            cb_unary(c, mpm[ipm], IN_CM_IPM);
            const double cost = __dadd_rn((double)satd, __dmul_rn((double)c.bits, it.sqrt_lambda0));
            int shift = 0;
            while(shift < 5 && cost < cand_cost[4 - shift]) shift++;
            if(shift) {
                for(int j = 1; j < shift; j++) { list[5 - j] = list[4 - j]; cand_cost[5 - j] = cand_cost[4 - j]; cand_satd[5 - j] = cand_satd[4 - j]; }
                list[5 - shift] = ipm; cand_cost[5 - shift] = cost; cand_satd[5 - shift] = satd;
            }
        }
        int          pc = 5;
        const double thr = __dmul_rn((double)it.inter_satd, 1.2);
        for(int i = 4; i >= 1; i--) {
            if((double)cand_satd[i] > thr) pc--;
            else break;
        }
#pragma unroll
        for(int k = 0; k < 5; k++) M.list[k] = list[k];
        M.pred_cnt = pc;
    }
    team_sync<T>();
    const int pred_cnt = M.pred_cnt;

    // ---- luma RDO per surviving mode (pintra_residue_rdo mode 0, src_base/xeve_pintra.c:97-152) ---------------------
    // pass 1 (team): transform, RDOQ, reconstruction and SSD of every surviving mode; levels kept in zig-zag order
    int16_t *g_coef = coef + it.out_off, *g_rec = rec ? rec + it.out_off : nullptr;
    for(int j = 0; j < pred_cnt; j++) {
        const int ipm = M.list[j];
        for(int e = tt; e < NY; e += T) M.blk[e] = (int16_t)(M.org[e] - ipred_at(leY, upY, ipm, e >> L2, e & (N - 1), dcY));
        team_sync<T>();
        fwd_dct_t<L2, T>(M.blk, M.TB, tm, tmT, bd, tt);
        const int nnz = quant_team<L2, T, true>(M.blk, M.TB, it.qp[0], lambda0, 0, slice_type, rt, bd, sq.rdoq, tt, M.X);
        for(int e = tt; e < NY; e += T) candS[j][zz_of(e, L2)] = M.blk[e];
        if(nnz) {
            team_sync<T>();
            dequant_team<L2, T>(M.blk, it.qp[0], bd, tt);
            inv_dct_t<L2, T>(M.blk, M.TB, tm, bd, tt);
        }
        int64_t ssd = 0;
        for(int e = tt; e < NY; e += T) {
            const int     pr = ipred_at(leY, upY, ipm, e >> L2, e & (N - 1), dcY);
            const int16_t t = nnz ? (int16_t)(M.blk[e] + pr) : (int16_t)pr;
            const int     r = clip3i(0, maxv, t), d = r - M.org[e];
            ssd += (int64_t)((d * d) >> sh);
        }
        team_sync<T>();
        ssd = team_sum_s64<T>(ssd, tt, M.X);
        if(tt == 0) { M.cand_ssd[j] = ssd; M.cand_nnz[j] = nnz; }
    }
    team_sync<T>();
    // pass 2 (one lane per mode): xeve_rdo_bit_cnt_cu_intra_luma (src_base/xeve_mode.c:81-119) of all modes at once --
    // the bins of one mode are serial, the modes are independent, so lane j codes mode j on its own copy of the models
    if(tt < pred_cnt) {
        uint16_t *ml = M.cm_lane + tt;
        for(int k = 0; k < IN_CM_N; k++) ml[k * 8] = M.cm_base[k];
        TCabac c;
        c.range = range_base; c.bits = 0; c.m = ml;
        if(slice_type != 2 && all_preds) {
            tc_bin<8>(c, XB200_CM_SKIP_FLAG + ctx_skip, 0);
            tc_bin<8>(c, XB200_CM_PRED_MODE + ctx_pm, 1);
        }
        tc_unary<8>(c, mpm[M.list[tt]], IN_CM_IPM);
        int num_sig = M.cand_nnz[tt];
        tc_bin<8>(c, XB200_CM_CBF_LUMA, num_sig != 0);
        if(num_sig) { // xeve_eco_run_length_cc over the zig-zag ordered levels; ends with the last significant one
            const int16_t *lv = candS[tt];
            uint32_t       run = 0;
            for(int sp = 0; sp < NY; sp++) {
                const int v = lv[sp];
                if(v == 0) { run++; continue; }
                tc_unary<8>(c, run, XB200_CM_RUN);
                tc_unary<8>(c, (uint32_t)abs(v) - 1, XB200_CM_LEVEL);
                tc_ep(c);
                if(sp == NY - 1) break;
                run = 0;
                num_sig--;
                tc_bin<8>(c, XB200_CM_LAST, num_sig == 0);
                if(num_sig == 0) break;
            }
        }
        M.cand_bits[tt] = c.bits;
    }
    __syncwarp();
    team_sync<T>();
    double  cost = IN_MAX_COST;
    int     best_j = 0;
    int32_t best_dist_y = 0;
    for(int j = 0; j < pred_cnt; j++) { // first minimum in list order (strict <), every thread alike
        double        cost_t = (double)M.cand_ssd[j];
        const int32_t dist_t = (int32_t)cost_t;
        cost_t = __dadd_rn(cost_t, __dmul_rn((double)M.cand_bits[j], lambda0));
        if(cost_t < cost) { cost = cost_t; best_dist_y = dist_t; best_j = j; }
    }
    const int best_ipd = M.list[best_j], nnz_best0 = M.cand_nnz[best_j];
    // the winner's levels go out in raster order; its reconstruction is rebuilt from them (cheaper than keeping five)
    for(int e = tt; e < NY; e += T) {
        const int16_t v = candS[best_j][zz_of(e, L2)];
        g_coef[e] = v;
        M.blk[e] = v;
    }
    if(g_rec) {
        team_sync<T>();
        if(nnz_best0) {
            dequant_team<L2, T>(M.blk, it.qp[0], bd, tt);
            inv_dct_t<L2, T>(M.blk, M.TB, tm, bd, tt);
        }
        for(int e = tt; e < NY; e += T) {
            const int     pr = ipred_at(leY, upY, best_ipd, e >> L2, e & (N - 1), dcY);
            const int16_t t = nnz_best0 ? (int16_t)(M.blk[e] + pr) : (int16_t)pr;
            g_rec[e] = (int16_t)clip3i(0, maxv, t);
        }
    }
    team_sync<T>();

    // ---- chroma with the winning luma mode (pintra_residue_rdo mode 1, :153-270); its own bit count is never used ---
    int     nnzc[2] = {0, 0};
    int64_t ssdc[2] = {0, 0};
#pragma unroll
    for(int c = 1; c < 3; c++) {
        const int16_t *le = leC[c - 1], *up = upC[c - 1], *og = M.org + NY + (c - 1) * NCH;
        const int      dc = ipred_dc(le, up, LC);
        for(int e = tt; e < NCH; e += T) M.blk[e] = (int16_t)(og[e] - ipred_at(le, up, best_ipd, e >> LC, e & (NC - 1), dc));
        team_sync<T>();
        fwd_dct_t<LC, T>(M.blk, M.TB, tm, tmT, bd, tt);
        const int nz = quant_team<LC, T, true>(M.blk, M.TB, it.qp[c], it.lambda[c], c, slice_type, rt, bd, sq.rdoq, tt, M.X);
        nnzc[c - 1] = nz;
        for(int e = tt; e < NCH; e += T) {
            const int16_t v = M.blk[e];
            g_coef[NY + (c - 1) * NCH + e] = v;
            M.chS[(c - 1) * NCH + zz_of(e, LC)] = v;
        }
        team_sync<T>();
        if(nz) {
            dequant_team<LC, T>(M.blk, it.qp[c], bd, tt);
            inv_dct_t<LC, T>(M.blk, M.TB, tm, bd, tt);
        }
        int64_t ssd = 0;
        for(int e = tt; e < NCH; e += T) {
            const int     pr = ipred_at(le, up, best_ipd, e >> LC, e & (NC - 1), dc);
            const int16_t t = nz ? (int16_t)(M.blk[e] + pr) : (int16_t)pr;
            const int     r = clip3i(0, maxv, t), d = r - og[e];
            if(g_rec) g_rec[NY + (c - 1) * NCH + e] = (int16_t)r;
            ssd += (int64_t)((d * d) >> sh);
        }
        team_sync<T>();
        ssdc[c - 1] = team_sum_s64<T>(ssd, tt, M.X);
    }
    const int32_t best_dist_c = (int32_t)__dadd_rn(__dmul_rn(it.dist_chroma_weight[0], (double)ssdc[0]),
                                                   __dmul_rn(it.dist_chroma_weight[1], (double)ssdc[1]));

    // ---- final bit count of the CU from the input state (xeve_rdo_bit_cnt_cu_intra, src_base/xeve_mode.c:141-171) ----
    for(int k = tt; k < IN_CM_N; k += T) M.cm_run[k] = M.cm_base[k];
    team_sync<T>();
    if(tt < 32) {
        Cabac c;
        c.range = range_base; c.bits = 0; c.m = M.cm_run;
        if(lane == 0) {
            if(slice_type != 2) {
                cb_bin(c, XB200_CM_SKIP_FLAG + ctx_skip, 0);
                cb_bin(c, XB200_CM_PRED_MODE + ctx_pm, 1);
            }
            cb_unary(c, mpm[best_ipd], IN_CM_IPM);
            cb_bin(c, XB200_CM_CBF_CB, nnzc[0] != 0);
            cb_bin(c, XB200_CM_CBF_CR, nnzc[1] != 0);
            cb_bin(c, XB200_CM_CBF_LUMA, nnz_best0 != 0);
        }
        if(nnz_best0) cb_run_length_sm(c, candS[best_j], NY, nnz_best0, 0, lane);
        if(nnzc[0]) cb_run_length_sm(c, M.chS, NCH, nnzc[0], 1, lane);
        if(nnzc[1]) cb_run_length_sm(c, M.chS + NCH, NCH, nnzc[1], 2, lane);
        if(lane == 0) { M.bits = c.bits; M.range_run = c.range; }
        __syncwarp();
    }
    team_sync<T>();
    if(tt == 0) {
        double ct = __dmul_rn((double)M.bits, lambda0);
        ct = __dadd_rn(ct, (double)best_dist_y);
        ct = __dadd_rn(ct, (double)best_dist_c);
        it.cost = ct;
        it.dist_cu = best_dist_y + best_dist_c;
        it.ipm[0] = it.ipm[1] = (int8_t)best_ipd;
        it.nnz[0] = nnz_best0; it.nnz[1] = nnzc[0]; it.nnz[2] = nnzc[1];
        it.cm_ipm_out[0] = M.cm_run[IN_CM_IPM]; it.cm_ipm_out[1] = M.cm_run[IN_CM_IPM + 1];
        st_out[it.state_out].range = M.range_run;
    }
    {
        xb200_sbac &so = st_out[it.state_out];
        for(int k = tt; k < XB200_CM_COUNT; k += T) so.m[k] = M.cm_run[k];
    }
    team_sync<T>();
}

template <int L2>
__global__ void __launch_bounds__(IntraCfg<L2>::CTA) k_intra(const PicDev *__restrict__ pics, xb200_intra_item *items,
                                                             const int32_t *__restrict__ order, int cnt,
                                                             const xb200_rates *__restrict__ rates, const xb200_sbac *__restrict__ st_in,
                                                             xb200_sbac *__restrict__ st_out, const int16_t *__restrict__ side,
                                                             int16_t *__restrict__ coef, int16_t *__restrict__ rec,
                                                             const int8_t *__restrict__ g_tm64, SeqDev sq)
{
    using Cf = IntraCfg<L2>;
    constexpr int T = Cf::T;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int8_t *tm = reinterpret_cast<int8_t *>(smem_raw), *tmT = tm + 4096;
    const int team = threadIdx.x / T, tt = threadIdx.x % T;
    IntraTeam<L2> &M = reinterpret_cast<IntraTeam<L2> *>(smem_raw + 8192)[team];
    for(int e = threadIdx.x; e < 4096; e += Cf::CTA) {
        const int8_t v = g_tm64[e];
        tm[e] = v;
        tmT[(e & 63) * 64 + (e >> 6)] = v;
    }
    __syncthreads();

    for(int ii = blockIdx.x * Cf::TEAMS + team; ii < cnt; ii += gridDim.x * Cf::TEAMS)
        intra_cu_one<L2>(M, tm, tmT, pics, items[order[ii]], rates, st_in, st_out, side, coef, rec, sq, tt);
}

// =====================================================================================================================
// 4x4 and 8x8 CUs: one THREAD per CU.  Nine tenths of the CUs of an intra picture are 4x4 or 8x8: a warp per CU leaves most
// lanes idle through five transform + RDOQ + coder passes (16 k warp-instructions per 4x4 CU).  Here every step is the
// plain sequential algorithm on arrays private to the thread (registers for 4x4, local memory for 8x8); only the context
// models live in shared memory, laid out [model][thread] so that the 32 lanes of a warp hit 32 different banks.
// =====================================================================================================================

// zig-zag scan position -> raster index: closed forms for N = 2 and N = 4, a shared-memory table (built by the kernel from
// zz_of) for N = 8
template <int LN> XB_DEV int zz_raster(int sp, const uint8_t *zinv8)
{
    if(LN == 1) return sp;
    if(LN == 2) return (int)((0xFEB7ADC963258410ull >> (4 * sp)) & 15); // 0 1 4 8 5 2 3 6 9 12 13 10 7 11 14 15
    return zinv8[sp];
}

// xeve_eco_run_length_cc of one block whose levels are in raster order
template <int LN, int CS = I4_THREADS> XB_DEV void tc_run_length(TCabac &c, const int16_t *lev, int num_sig, int ch, const uint8_t *zinv8)
{
    constexpr int n = 1 << (2 * LN);
    const int     t0 = ch == 0 ? 0 : 2;
    uint32_t      run = 0;
#pragma unroll(LN <= 2 ? 16 : 1)
    for(int sp = 0; sp < n; sp++) {
        const int v = lev[zz_raster<LN>(sp, zinv8)];
        if(v == 0) { run++; continue; }
        tc_unary<CS>(c, run, XB200_CM_RUN + t0);
        tc_unary<CS>(c, (uint32_t)abs(v) - 1, XB200_CM_LEVEL + t0);
        tc_ep(c);
        if(sp == n - 1) break;
        run = 0;
        num_sig--;
        tc_bin<CS>(c, XB200_CM_LAST + (ch != 0), num_sig == 0);
        if(num_sig == 0) break;
    }
}

// forward / inverse transform of a private N x N block (src_base/xeve_tq.c:396-404, src_base/xeve_itdq.c:435-440)
template <int LN> XB_DEV void fwd_dct_thr(int16_t *b, int bd)
{
    constexpr int N = 1 << LN;
    const int     shift = (LN - 1 + bd - 8) + (LN + 6);
    int           t[N * N];
#pragma unroll(LN <= 2 ? 4 : 1)
    for(int y = 0; y < N; y++)
#pragma unroll
        for(int u = 0; u < N; u++) {
            int acc = 0;
#pragma unroll
            for(int x = 0; x < N; x++) acc += tmN(LN, u, x) * (int)b[y * N + x];
            t[y * N + u] = acc;
        }
#pragma unroll(LN <= 2 ? 4 : 1)
    for(int u = 0; u < N; u++)
#pragma unroll
        for(int v = 0; v < N; v++) {
            int64_t acc = 0;
#pragma unroll
            for(int y = 0; y < N; y++) acc += (int64_t)tmN(LN, v, y) * (int64_t)t[y * N + u];
            b[v * N + u] = (int16_t)((acc + ((int64_t)1 << (shift - 1))) >> shift);
        }
}
template <int LN> XB_DEV void inv_dct_thr(int16_t *b, int bd)
{
    constexpr int N = 1 << LN;
    const int     shift = 7 + 12 - (bd - 8);
    int           t[N * N];
#pragma unroll(LN <= 2 ? 4 : 1)
    for(int u = 0; u < N; u++)
#pragma unroll
        for(int y = 0; y < N; y++) {
            int acc = 0; // |sum| <= 8 * 32768 * 90 fits 32 bits
#pragma unroll
            for(int v = 0; v < N; v++) acc += tmN(LN, v, y) * (int)b[v * N + u];
            t[y * N + u] = acc;
        }
#pragma unroll(LN <= 2 ? 4 : 1)
    for(int y = 0; y < N; y++)
#pragma unroll
        for(int x = 0; x < N; x++) {
            int64_t acc = 0;
#pragma unroll
            for(int u = 0; u < N; u++) acc += (int64_t)tmN(LN, u, x) * (int64_t)t[y * N + u];
            acc = (acc + ((int64_t)1 << (shift - 1))) >> shift;
            b[y * N + x] = (int16_t)max((int64_t)-32768, min((int64_t)32767, acc));
        }
}

// xeve_quant_nnz + xeve_rdoq_run_length_cc (src_base/xeve_tq.c:497-730), sequential, intra block; b: coefficients -> levels
template <int LN>
XB_DEV int rdoq_thr(int16_t *b, int qp, double d_lambda, int ch, int slice_type, const xb200_rates *__restrict__ rt, int bd, int use_rdoq,
                    const uint8_t *zinv8)
{
    constexpr int n = 1 << (2 * LN);
    const int     q = c_quant_scale[qp % 6], qbits = 14 + (15 - bd - LN) + qp / 6;
    if(!use_rdoq) {
        const int32_t off = (int32_t)(slice_type == 2 ? 171 : 85) << (qbits - 9);
        int           cnt = 0;
#pragma unroll(LN <= 2 ? 16 : 1)
        for(int e = 0; e < n; e++) {
            const int     c = b[e];
            const int32_t lev = (int16_t)(((int32_t)abs(c) * q + off) >> qbits);
            b[e] = (int16_t)(c < 0 ? -lev : lev);
            cnt += lev != 0;
        }
        return cnt;
    }
    const int64_t thr = ((int64_t)1 << qbits) - ((int64_t)(slice_type == 2 ? 201 : 153) << (qbits - 9));
    RdoqEnv E;
    E.lambda = (int64_t)(d_lambda * 32768.0 + 0.5);
    E.es     = c_err_scale[bd - 8][qp % 6][LN];
    E.qbits  = qbits;
    int     coded = 0, any = 0;
    int64_t unc_blk = 0;
#pragma unroll(LN <= 2 ? 16 : 1)
    for(int e = 0; e < n; e++) {
        const int c = b[e];
        int64_t   ld; uint32_t maxl;
        rq_quant(c, q, qbits, ld, maxl);
        const int64_t e0 = (ld * E.es) >> 20;
        unc_blk += e0 * e0;
        coded |= ((int64_t)abs(c) * q >= thr);
        any |= maxl != 0;
    }
    if(!coded || !any) {
#pragma unroll(LN <= 2 ? 16 : 1)
        for(int e = 0; e < n; e++) b[e] = 0;
        return 0;
    }
    {
        const int ctx = ch == 0 ? 0 : 2;
        E.run[0][0] = rt->run[ctx][0]; E.run[0][1] = rt->run[ctx][1];
        E.run[1][0] = rt->run[ctx + 1][0]; E.run[1][1] = rt->run[ctx + 1][1];
        E.lev[0][0] = rt->level[ctx][0]; E.lev[0][1] = rt->level[ctx][1];
        E.lev[1][0] = rt->level[ctx + 1][0]; E.lev[1][1] = rt->level[ctx + 1][1];
    }
    const int32_t *cbf = ch == 0 ? rt->cbf_luma : (ch == 1 ? rt->cbf_cb : rt->cbf_cr);
    int64_t        best = unc_blk + (int64_t)cbf[0] * E.lambda, base = unc_blk + (int64_t)cbf[1] * E.lambda;
    const int64_t  last0 = (int64_t)rt->last[ch == 0 ? 0 : 1][0] * E.lambda, last1 = (int64_t)rt->last[ch == 0 ? 0 : 1][1] * E.lambda;
    int            state = 0, best_last = 0;
#pragma unroll(LN <= 2 ? 16 : 1)
    for(int sp = 0; sp < n; sp++) {
        const int p = zz_raster<LN>(sp, zinv8), c = b[p];
        int64_t   ld, dl; uint32_t maxl;
        rq_quant(c, q, qbits, ld, maxl);
        const uint32_t lev = rq_level(E, ld, maxl, state, dl);
        b[p] = (int16_t)(c > 0 ? (int)lev : -(int)lev);
        base += dl;
        if(lev) {
            const int64_t as_last = base + last1;
            base += last0;
            if(as_last < best) { best = as_last; best_last = sp + 1; }
            state = 0;
        }
        else state = 1;
    }
    int nnz = 0;
#pragma unroll(LN <= 2 ? 16 : 1)
    for(int sp = 0; sp < n; sp++) {
        const int p = zz_raster<LN>(sp, zinv8);
        if(sp < best_last) nnz += b[p] != 0;
        else b[p] = 0;
    }
    return nnz;
}
template <int LN> XB_DEV void dequant_thr(int16_t *b, int qp, int bd)
{
    constexpr int n = 1 << (2 * LN);
    const int     shift = 20 - 14 - (15 - bd - LN);
    const int64_t scale = (int64_t)c_dequant_scale[qp % 6] << (qp / 6), off = shift ? (int64_t)1 << (shift - 1) : 0;
#pragma unroll(LN <= 2 ? 16 : 1)
    for(int e = 0; e < n; e++) {
        const int64_t v = ((int64_t)b[e] * scale + off) >> shift;
        b[e] = (int16_t)max((int64_t)-32768, min((int64_t)32767, v));
    }
}

// the context models a Baseline intra CU can touch: skip_flag[ctx], pred_mode[ctx], intra_dir[0..1], cbf_luma / cb / cr,
// run[0..3], level[0..3], last[0..1]
XB_DEV int i4_model(int k, int ctx_skip, int ctx_pm)
{
    return k == 0 ? XB200_CM_SKIP_FLAG + ctx_skip : k == 1 ? XB200_CM_PRED_MODE + ctx_pm : k < 4 ? IN_CM_IPM + (k - 2)
         : k == 4 ? XB200_CM_CBF_LUMA : k == 5 ? XB200_CM_CBF_CB : k == 6 ? XB200_CM_CBF_CR : k < 11 ? XB200_CM_RUN + (k - 7)
         : k < 15 ? XB200_CM_LEVEL + (k - 11) : XB200_CM_LAST + (k - 15);
}
constexpr int I4_MODELS = 17;

// pintra_analyze_cu of ONE 4x4 / 8x8 CU by ONE thread; mb / mr: the thread's two model sets, model k at [k * CS]
template <int L2, int CS>
__device__ __forceinline__ void intra_thr_one(const PicDev *__restrict__ pics, xb200_intra_item &it, const xb200_rates *rates,
                                              const xb200_sbac *st_in, xb200_sbac *st_out, const int16_t *side, int16_t *coef, int16_t *rec,
                                              const SeqDev &sq, uint16_t *mb, uint16_t *mr, const uint8_t *zinv8)
{
    constexpr int N = 1 << L2, NY = N * N, NC = N / 2, NCH = NC * NC, LC = L2 - 1;
    constexpr int NBY = 2 * N + 1, NBC = N + 1;   // samples per reference array (index -1 .. 2n-1)
    constexpr int UR = L2 == 2 ? 64 : 1;          // arrays of 4x4 CUs stay in registers (every loop over them fully unrolled)
    const int bd = sq.bd, maxv = (1 << bd) - 1, sh = (bd - 8) << 1;
    const int slice_type = it.slice_type, all_preds = it.all_preds, ctx_skip = it.ctx_skip, ctx_pm = it.ctx_pred_mode;
    const xb200_rates *rt = &rates[it.rate_idx];
    const double       lambda0 = it.lambda[0];
    const xb200_sbac &s0 = st_in[it.state_in];
    const uint32_t    range_base = s0.range;
#pragma unroll
    for(int k = 0; k < I4_MODELS; k++) {
        const int idx = i4_model(k, ctx_skip, ctx_pm);
        mb[idx * CS] = idx >= IN_CM_IPM ? it.cm_ipm_in[idx - IN_CM_IPM] : s0.m[idx];
    }
    uint8_t mpm[5];
#pragma unroll
    for(int k = 0; k < 5; k++) mpm[k] = it.mpm[k];
    // original block and reference samples
    int16_t org[NY], orgc[2][NCH];
    {
        const PicDev   p = pics[it.cur_pic];
        const int      x0 = it.x, y0 = it.y;
        const int16_t *gy = p.p[0] + (ptrdiff_t)y0 * p.s[0] + x0;
#pragma unroll(UR)
        for(int q = 0; q < NY / 4; q++) { // x0 is a multiple of 4: 8-byte aligned quads
            const int   r = q / (N / 4), cq = (q % (N / 4)) * 4;
            const uint2 v = *reinterpret_cast<const uint2 *>(gy + (ptrdiff_t)r * p.s[0] + cq);
            org[r * N + cq + 0] = (int16_t)(v.x & 0xffff); org[r * N + cq + 1] = (int16_t)(v.x >> 16);
            org[r * N + cq + 2] = (int16_t)(v.y & 0xffff); org[r * N + cq + 3] = (int16_t)(v.y >> 16);
        }
#pragma unroll
        for(int c = 1; c < 3; c++) {
            const int16_t *gc = p.p[c] + (ptrdiff_t)(y0 >> 1) * p.s[c] + (x0 >> 1);
#pragma unroll(UR)
            for(int q = 0; q < NCH / 2; q++) {
                const int      r = q / (NC / 2), cq = (q % (NC / 2)) * 2;
                const uint32_t v = *reinterpret_cast<const uint32_t *>(gc + (ptrdiff_t)r * p.s[c] + cq);
                orgc[c - 1][r * NC + cq] = (int16_t)(v & 0xffff); orgc[c - 1][r * NC + cq + 1] = (int16_t)(v >> 16);
            }
        }
    }
    int16_t nbY[2 * NBY], nbC[2][2 * NBC]; // left[-1 .. 2n-1] | up[-1 .. 2n-1]
    {
        const int16_t *gn = side + it.nb_off;
#pragma unroll(UR)
        for(int k = 0; k < 2 * NBY; k++) nbY[k] = gn[k];
#pragma unroll
        for(int c = 0; c < 2; c++)
#pragma unroll(UR)
            for(int k = 0; k < 2 * NBC; k++) nbC[c][k] = gn[2 * NBY + c * 2 * NBC + k];
    }
    const int16_t *leY = nbY + 1, *upY = nbY + NBY + 1;
    const int      dcY = ipred_dc(leY, upY, L2);

    // ---- make_ipred_list -------------------------------------------------------------------------------------------------
    double   cand_cost[5];
    uint32_t cand_satd[5];
    int      list[5];
#pragma unroll
    for(int k = 0; k < 5; k++) { list[k] = 0; cand_cost[k] = IN_MAX_COST; cand_satd[k] = 0xffffffffu; }
    TCabac c;
    c.m = mr;
#pragma unroll 1
    for(int ipm = 0; ipm < 5; ipm++) {
        uint32_t satd;
        if(L2 == 2) satd = (uint32_t)had_tile_fn<4>(org, N, [&](int y, int x) { return ipred_at(leY, upY, ipm, y, x, dcY); });
        else satd = (uint32_t)had_tile_fn<8>(org, N, [&](int y, int x) { return ipred_at(leY, upY, ipm, y, x, dcY); });
        satd >>= (bd - 8);
        c.range = range_base; c.bits = 0;
        mr[IN_CM_IPM * CS] = mb[IN_CM_IPM * CS]; mr[(IN_CM_IPM + 1) * CS] = mb[(IN_CM_IPM + 1) * CS];
        tc_unary<CS>(c, mpm[ipm], IN_CM_IPM);
        const double cost = __dadd_rn((double)satd, __dmul_rn((double)c.bits, it.sqrt_lambda0));
        int shift = 0;
        while(shift < 5 && cost < cand_cost[4 - shift]) shift++;
        if(shift) {
            for(int j = 1; j < shift; j++) { list[5 - j] = list[4 - j]; cand_cost[5 - j] = cand_cost[4 - j]; cand_satd[5 - j] = cand_satd[4 - j]; }
            list[5 - shift] = ipm; cand_cost[5 - shift] = cost; cand_satd[5 - shift] = satd;
        }
    }
    int pred_cnt = 5;
    {
        const double thr = __dmul_rn((double)it.inter_satd, 1.2);
        for(int i = 4; i >= 1; i--) {
            if((double)cand_satd[i] > thr) pred_cnt--;
            else break;
        }
    }

    // ---- luma RDO per surviving mode --------------------------------------------------------------------------------------
    double  cost = IN_MAX_COST;
    int     best_ipd = 0, nnz_best0 = 0;
    int32_t best_dist_y = 0;
    int16_t best_lev[NY], best_rec[NY];
#pragma unroll 1
    for(int j = 0; j < pred_cnt; j++) {
        const int ipm = list[j];
        int16_t   b[NY], lev[NY];
#pragma unroll(UR)
        for(int e = 0; e < NY; e++) b[e] = (int16_t)(org[e] - ipred_at(leY, upY, ipm, e >> L2, e & (N - 1), dcY));
        fwd_dct_thr<L2>(b, bd);
        const int nnz = rdoq_thr<L2>(b, it.qp[0], lambda0, 0, slice_type, rt, bd, sq.rdoq, zinv8);
#pragma unroll(UR)
        for(int e = 0; e < NY; e++) lev[e] = b[e];
#pragma unroll
        for(int k = 0; k < I4_MODELS; k++) { const int idx = i4_model(k, ctx_skip, ctx_pm); mr[idx * CS] = mb[idx * CS]; }
        c.range = range_base; c.bits = 0;
        if(slice_type != 2 && all_preds) {
            tc_bin<CS>(c, XB200_CM_SKIP_FLAG + ctx_skip, 0);
            tc_bin<CS>(c, XB200_CM_PRED_MODE + ctx_pm, 1);
        }
        tc_unary<CS>(c, mpm[ipm], IN_CM_IPM);
        tc_bin<CS>(c, XB200_CM_CBF_LUMA, nnz != 0);
        if(nnz) {
            tc_run_length<L2, CS>(c, lev, nnz, 0, zinv8);
            dequant_thr<L2>(b, it.qp[0], bd);
            inv_dct_thr<L2>(b, bd);
        }
        int64_t ssd = 0;
#pragma unroll(UR)
        for(int e = 0; e < NY; e++) {
            const int     pr = ipred_at(leY, upY, ipm, e >> L2, e & (N - 1), dcY);
            const int16_t t = nnz ? (int16_t)(b[e] + pr) : (int16_t)pr;
            const int     r = clip3i(0, maxv, t), d = r - org[e];
            b[e] = (int16_t)r;
            ssd += (int64_t)((d * d) >> sh);
        }
        double        cost_t = (double)ssd;
        const int32_t dist_t = (int32_t)cost_t;
        cost_t = __dadd_rn(cost_t, __dmul_rn((double)c.bits, lambda0));
        if(cost_t < cost) {
            cost = cost_t; best_dist_y = dist_t; best_ipd = ipm; nnz_best0 = nnz;
#pragma unroll(UR)
            for(int e = 0; e < NY; e++) { best_lev[e] = lev[e]; best_rec[e] = b[e]; }
        }
    }

    // ---- chroma with the winning luma mode ---------------------------------------------------------------------------------
    int16_t *g_coef = coef + it.out_off, *g_rec = rec ? rec + it.out_off : nullptr;
    int      nnzc[2];
    int64_t  ssdc[2];
    int16_t  levc[2][NCH];
#pragma unroll
    for(int cc = 0; cc < 2; cc++) {
        const int16_t *le = nbC[cc] + 1, *up = nbC[cc] + NBC + 1;
        const int      dc = ipred_dc(le, up, LC);
        int16_t        b[NCH];
#pragma unroll
        for(int e = 0; e < NCH; e++) b[e] = (int16_t)(orgc[cc][e] - ipred_at(le, up, best_ipd, e >> LC, e & (NC - 1), dc));
        fwd_dct_thr<LC>(b, bd);
        const int nz = rdoq_thr<LC>(b, it.qp[cc + 1], it.lambda[cc + 1], cc + 1, slice_type, rt, bd, sq.rdoq, zinv8);
        nnzc[cc] = nz;
#pragma unroll
        for(int e = 0; e < NCH; e++) { levc[cc][e] = b[e]; g_coef[NY + cc * NCH + e] = b[e]; }
        if(nz) {
            dequant_thr<LC>(b, it.qp[cc + 1], bd);
            inv_dct_thr<LC>(b, bd);
        }
        int64_t ssd = 0;
#pragma unroll
        for(int e = 0; e < NCH; e++) {
            const int     pr = ipred_at(le, up, best_ipd, e >> LC, e & (NC - 1), dc);
            const int16_t t = nz ? (int16_t)(b[e] + pr) : (int16_t)pr;
            const int     r = clip3i(0, maxv, t), d = r - orgc[cc][e];
            if(g_rec) g_rec[NY + cc * NCH + e] = (int16_t)r;
            ssd += (int64_t)((d * d) >> sh);
        }
        ssdc[cc] = ssd;
    }
    const int32_t best_dist_c = (int32_t)__dadd_rn(__dmul_rn(it.dist_chroma_weight[0], (double)ssdc[0]),
                                                   __dmul_rn(it.dist_chroma_weight[1], (double)ssdc[1]));
#pragma unroll(UR)
    for(int e = 0; e < NY; e++) {
        g_coef[e] = best_lev[e];
        if(g_rec) g_rec[e] = best_rec[e];
    }

    // ---- final bit count from the input state --------------------------------------------------------------------------------
#pragma unroll
    for(int k = 0; k < I4_MODELS; k++) { const int idx = i4_model(k, ctx_skip, ctx_pm); mr[idx * CS] = mb[idx * CS]; }
    c.range = range_base; c.bits = 0;
    if(slice_type != 2) {
        tc_bin<CS>(c, XB200_CM_SKIP_FLAG + ctx_skip, 0);
        tc_bin<CS>(c, XB200_CM_PRED_MODE + ctx_pm, 1);
    }
    tc_unary<CS>(c, mpm[best_ipd], IN_CM_IPM);
    tc_bin<CS>(c, XB200_CM_CBF_CB, nnzc[0] != 0);
    tc_bin<CS>(c, XB200_CM_CBF_CR, nnzc[1] != 0);
    tc_bin<CS>(c, XB200_CM_CBF_LUMA, nnz_best0 != 0);
    if(nnz_best0) tc_run_length<L2, CS>(c, best_lev, nnz_best0, 0, zinv8);
    if(nnzc[0]) tc_run_length<LC, CS>(c, levc[0], nnzc[0], 1, zinv8);
    if(nnzc[1]) tc_run_length<LC, CS>(c, levc[1], nnzc[1], 2, zinv8);

    double ct = __dmul_rn((double)c.bits, lambda0);
    ct = __dadd_rn(ct, (double)best_dist_y);
    ct = __dadd_rn(ct, (double)best_dist_c);
    it.cost = ct;
    it.dist_cu = best_dist_y + best_dist_c;
    it.ipm[0] = it.ipm[1] = (int8_t)best_ipd;
    it.nnz[0] = nnz_best0; it.nnz[1] = nnzc[0]; it.nnz[2] = nnzc[1];
    it.cm_ipm_out[0] = mr[IN_CM_IPM * CS]; it.cm_ipm_out[1] = mr[(IN_CM_IPM + 1) * CS];
    // output state = input state with the models this CU touched replaced
    xb200_sbac &so = st_out[it.state_out];
    so.range = c.range;
    for(int k = 0; k < XB200_CM_COUNT; k++) so.m[k] = s0.m[k];
#pragma unroll
    for(int k = 0; k < I4_MODELS; k++) {
        const int idx = i4_model(k, ctx_skip, ctx_pm);
        if(idx < XB200_CM_COUNT) so.m[idx] = mr[idx * CS];
    }
}

template <int L2>
__global__ void __launch_bounds__(I4_THREADS) k_intra_thr(const PicDev *__restrict__ pics, xb200_intra_item *items,
                                                          const int32_t *__restrict__ order, int cnt, const xb200_rates *__restrict__ rates,
                                                          const xb200_sbac *__restrict__ st_in, xb200_sbac *__restrict__ st_out,
                                                          const int16_t *__restrict__ side, int16_t *__restrict__ coef,
                                                          int16_t *__restrict__ rec, SeqDev sq)
{
    __shared__ uint16_t cm_base[IN_CM_N * I4_THREADS], cm_run[IN_CM_N * I4_THREADS];
    __shared__ uint8_t  zinv8[64];
    if(threadIdx.x < 64) zinv8[zz_of(threadIdx.x, 3)] = (uint8_t)threadIdx.x;
    __syncthreads();
    const int ii = blockIdx.x * I4_THREADS + threadIdx.x;
    if(ii >= cnt) return;
    intra_thr_one<L2, I4_THREADS>(pics, items[order[ii]], rates, st_in, st_out, side, coef, rec, sq, cm_base + threadIdx.x, cm_run + threadIdx.x, zinv8);
}

// =====================================================================================================================
// Reference samples + MPM list of a CU list from the picture reconstructed so far: xeve_get_avail_intra
// (src_base/xeve_util.c:717-772), xeve_get_nbr for Y, U, V (src_base/xeve_ipred.c:33-97), xeve_get_mpm (:230-252).
// One warp per CU; lanes stride over the 8N+6 output samples, each deciding the availability of its own 4x4 unit.
// =====================================================================================================================
XB200_CONST_LINKAGE __constant__ uint8_t c_mpm_tbl[6][6][5];

#ifndef XB200_CHAIN_TU
__global__ void k_intra_nbr(const PicDev *__restrict__ pics, int pic, xb200_nbr_item *__restrict__ items, int64_t n,
                            const uint32_t *__restrict__ map_scu, const int8_t *__restrict__ map_ipm, int w_scu, int h_scu, int cip, int bd,
                            int16_t *__restrict__ side)
{
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if(i >= n) return;
    const int       lane = threadIdx.x & 31;
    xb200_nbr_item &it = items[i];
    const PicDev    p = pics[pic];
    const int x = it.x, y = it.y, N = 1 << it.log2_cuw, scuw = N >> 2, scuh = (1 << it.log2_cuh) >> 2, xs = x >> 2, ys = y >> 2;
    const int scup = xs + ys * w_scu, half = 1 << (bd - 1);
    auto COD = [&](int q) { return (int)((map_scu[q] >> 31) & 1); };
    auto IFL = [&](int q) { return (int)((map_scu[q] >> 15) & 1); };
    unsigned av = 0;
    if(xs > 0 && COD(scup - 1)) {
        av |= 1u << 1;
        if(ys + scuh + scuw - 1 < h_scu && COD(scup + w_scu * (scuw + scuh) - w_scu - 1)) av |= 1u << 7;
    }
    if(ys > 0) {
        av |= (1u << 0) | (1u << 9);
        if(xs > 0 && COD(scup - w_scu - 1)) av |= 1u << 5;
        if(xs + scuw < w_scu && COD(scup - w_scu + scuw)) av |= 1u << 6;
    }
    if(xs + scuw < w_scu && COD(scup + scuw)) {
        av |= 1u << 3;
        if(ys + scuh + scuw - 1 < h_scu && COD(scup + w_scu * (scuw + scuh - 1) + scuw)) av |= 1u << 8;
    }
    const bool ul_ok = ((av >> 5) & 1) && (!cip || IFL(scup - w_scu - 1));
    if(lane == 0) {
        it.avail = (uint16_t)av;
        int ipm_l = 0, ipm_u = 0;
        if(xs > 0 && IFL(scup - 1) && COD(scup - 1)) ipm_l = map_ipm[scup - 1] + 1;
        if(ys > 0 && IFL(scup - w_scu) && COD(scup - w_scu)) ipm_u = map_ipm[scup - w_scu] + 1;
        ipm_l = min(max(ipm_l, 0), 5); ipm_u = min(max(ipm_u, 0), 5);
#pragma unroll
        for(int k = 0; k < 5; k++) it.mpm[k] = c_mpm_tbl[ipm_l][ipm_u][k];
    }
    int16_t *out = side + it.nb_off;
    for(int e = lane; e < 8 * N + 6; e += 32) {
        // plane / array / index of output sample e: Y left (2N+1), Y up (2N+1), then U, V with N+1 each
        int c, r = e;
        if(r < 2 * (2 * N + 1)) c = 0;
        else { r -= 2 * (2 * N + 1); c = 1 + r / (2 * (N + 1)); r %= 2 * (N + 1); }
        const int  nn = c ? N >> 1 : N, per = 2 * nn + 1, unit = c ? 2 : 4;
        const bool is_up = r >= per;
        const int  k = (is_up ? r - per : r) - 1; // -1 .. 2nn-1
        const int16_t *src = p.p[c] + (ptrdiff_t)(c ? y >> 1 : y) * p.s[c] + (c ? x >> 1 : x);
        const ptrdiff_t s = p.s[c];
        int v = half;
        if(k < 0) { if(ul_ok) v = src[-s - 1]; }
        else {
            const int u = k / unit;
            if(is_up) {
                if(ys > 0 && xs + u < w_scu && COD(scup - w_scu + u) && (!cip || IFL(scup - w_scu + u))) v = src[-s + k];
            }
            else if(xs > 0 && ys + u < h_scu && COD(scup - 1 + u * w_scu) && (!cip || IFL(scup - 1 + u * w_scu))) v = src[(ptrdiff_t)k * s - 1];
        }
        out[e] = (int16_t)v;
    }
}
#endif
