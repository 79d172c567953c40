// xb200_intra.cuh -- intra analysis of a list of CUs.
//
// Reference: pintra_analyze_cu / make_ipred_list / pintra_residue_rdo (src_base/xeve_pintra.c:69-374, 544-698), the five
// Baseline predictors (src_base/xeve_ipred.c:99-228) and the intra bit counters (src_base/xeve_mode.c:81-171) over the
// CABAC syntax of src_base/xeve_eco.c:793-905 (cbf), 1104-1121 (intra_dir), run-length coefficients.
//
// One kernel instance per CU size (all block dimensions compile-time).  A TEAM owns a CU from the mode ranking to the final
// bit count: one warp for 4x4 .. 16x16 CUs (four teams per CTA, warp-synchronous, no block barriers), 128 threads for 32x32
// and 64x64 -- the team transforms / parallel RDOQ of xb200_residue2.cuh.  Everything stays in shared memory: original
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
    static constexpr int T     = L2 <= 4 ? 32 : 128;   // (a 256-thread team for 64x64 showed a shared-memory hazard under racecheck)
    static constexpr int TEAMS = L2 <= 4 ? 4 : 1;
    static constexpr int CTA   = T * TEAMS;
    static constexpr int N     = 1 << L2, NY = N * N, NCH = NY / 4;
    static constexpr int TILES = L2 == 2 ? 1 : (N / 8) * (N / 8);   // SATD tiles per mode (4x4 CU: one 4x4 tile)
};

template <int L2> struct IntraTeam {
    static constexpr int N = 1 << L2, NY = N * N;
    int32_t  TB[NY < 32 ? 32 : NY];        // DCT stage buffer / RDOQ scratch
    int16_t  org[NY * 3 / 2];              // Y | U | V original block
    int16_t  blk[NY];                      // transform working block
    int16_t  levS[NY];                     // current luma levels, zig-zag order
    int16_t  bestS[NY];                    // best luma candidate, zig-zag order
    int16_t  chS[NY / 2];                  // chroma levels U | V, zig-zag order
    int16_t  nb[8 * N + 8];                // per plane: left[-1 .. 2n-1], up[-1 .. 2n-1]
    uint16_t cm_base[IN_CM_N + 2], cm_run[IN_CM_N + 2];
    int32_t  satd_part[5 * IntraCfg<L2>::TILES];
    uint32_t bits, range_run;
    int32_t  list[5], pred_cnt;
    TeamScratch X;
};

template <int L2>
__global__ void __launch_bounds__(IntraCfg<L2>::CTA) k_intra(const PicDev *__restrict__ pics, xb200_intra_item *items,
                                                             const int32_t *__restrict__ order, int cnt,
                                                             const xb200_rates *__restrict__ rates, const xb200_sbac *__restrict__ st_in,
                                                             xb200_sbac *__restrict__ st_out, const int16_t *__restrict__ side,
                                                             int16_t *__restrict__ coef, int16_t *__restrict__ rec,
                                                             const int8_t *__restrict__ g_tm64, SeqDev sq)
{
    using Cf = IntraCfg<L2>;
    constexpr int T = Cf::T, N = Cf::N, NY = Cf::NY, NC = N / 2, NCH = Cf::NCH, TILES = Cf::TILES, LC = L2 - 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int8_t *tm = reinterpret_cast<int8_t *>(smem_raw), *tmT = tm + 4096;
    const int team = threadIdx.x / T, tt = threadIdx.x % T, lane = tt & 31;
    IntraTeam<L2> &M = reinterpret_cast<IntraTeam<L2> *>(smem_raw + 8192)[team];
    for(int e = threadIdx.x; e < 4096; e += Cf::CTA) {
        const int8_t v = g_tm64[e];
        tm[e] = v;
        tmT[(e & 63) * 64 + (e >> 6)] = v;
    }
    __syncthreads();
    const int bd = sq.bd, maxv = (1 << bd) - 1, sh = (bd - 8) << 1;

    for(int ii = blockIdx.x * Cf::TEAMS + team; ii < cnt; ii += gridDim.x * Cf::TEAMS) {
        xb200_intra_item &it = items[order[ii]];
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
        double   cost = IN_MAX_COST;
        int      best_ipd = 0, nnz_best0 = 0;
        int32_t  best_dist_y = 0;
        int16_t *g_coef = coef + it.out_off, *g_rec = rec ? rec + it.out_off : nullptr;
        for(int j = 0; j < pred_cnt; j++) {
            const int ipm = M.list[j];
            for(int e = tt; e < NY; e += T) M.blk[e] = (int16_t)(M.org[e] - ipred_at(leY, upY, ipm, e >> L2, e & (N - 1), dcY));
            team_sync<T>();
            fwd_dct_t<L2, T>(M.blk, M.TB, tm, tmT, bd, tt);
            const int nnz = quant_team<L2, T, true>(M.blk, M.TB, it.qp[0], lambda0, 0, slice_type, rt, bd, sq.rdoq, tt, M.X);
            for(int e = tt; e < NY; e += T) M.levS[zz_of(e, L2)] = M.blk[e];
            for(int k = tt; k < IN_CM_N; k += T) M.cm_run[k] = M.cm_base[k];
            team_sync<T>();
            if(tt < 32) { // xeve_rdo_bit_cnt_cu_intra_luma, src_base/xeve_mode.c:81-119
                Cabac c;
                c.range = range_base; c.bits = 0; c.m = M.cm_run;
                if(lane == 0) {
                    if(slice_type != 2 && all_preds) {
                        cb_bin(c, XB200_CM_SKIP_FLAG + ctx_skip, 0);
                        cb_bin(c, XB200_CM_PRED_MODE + ctx_pm, 1);
                    }
                    cb_unary(c, mpm[ipm], IN_CM_IPM);
                    cb_bin(c, XB200_CM_CBF_LUMA, nnz != 0);
                }
                if(nnz) cb_run_length_sm(c, M.levS, NY, nnz, 0, lane);
                if(lane == 0) M.bits = c.bits;
                __syncwarp(); // lane 0 trails the others through its last bins: reconverge before any block-wide barrier
            }
            // candidate levels are still in blk (raster): keep them in registers for the "new best" copy below
            int16_t keep[(NY + T - 1) / T];
#pragma unroll
            for(int k = 0; k < (NY + T - 1) / T; k++) { const int e = tt + k * T; keep[k] = e < NY ? M.blk[e] : (int16_t)0; }
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
                M.blk[e] = (int16_t)r;
                ssd += (int64_t)((d * d) >> sh);
            }
            team_sync<T>(); // publishes M.bits and the reconstruction in blk
            ssd = team_sum_s64<T>(ssd, tt, M.X);
            double        cost_t = (double)ssd;
            const int32_t dist_t = (int32_t)cost_t;
            cost_t = __dadd_rn(cost_t, __dmul_rn((double)M.bits, lambda0));
            if(cost_t < cost) {
                cost = cost_t; best_dist_y = dist_t; best_ipd = ipm; nnz_best0 = nnz;
#pragma unroll
                for(int k = 0; k < (NY + T - 1) / T; k++) {
                    const int e = tt + k * T;
                    if(e < NY) {
                        g_coef[e] = keep[k];
                        M.bestS[e] = M.levS[e];
                        if(g_rec) g_rec[e] = M.blk[e];
                    }
                }
            }
            team_sync<T>();
        }

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
            if(nnz_best0) cb_run_length_sm(c, M.bestS, NY, nnz_best0, 0, lane);
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
}
