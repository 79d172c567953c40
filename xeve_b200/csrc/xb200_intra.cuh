// xb200_intra.cuh -- intra analysis of a list of CUs.
//
// Reference: pintra_analyze_cu / make_ipred_list / pintra_residue_rdo (src_base/xeve_pintra.c:69-374, 544-698), the five
// Baseline predictors (src_base/xeve_ipred.c:99-228) and the intra bit counters (src_base/xeve_mode.c:81-171) over the
// CABAC syntax of src_base/xeve_eco.c:793-905 (cbf), 1104-1121 (intra_dir), run-length coefficients.
//
// One team (a CTA: one warp for CUs up to 8x8, four warps above) owns a CU from the mode ranking to the final bit count,
// everything in shared memory: original block, reference samples, the transform working set of xb200_tq.cuh, and the
// quantised levels of the current / best candidate in zig-zag order for the coder.  Predictions are never stored: a
// predictor is two or three shared-memory reads, so each use recomputes it.  CTAs are persistent (the DCT matrix is
// staged once per CTA) and size classes run as separate template instances with right-sized shared memory.
// Costs are IEEE doubles evaluated in the reference's operation order with explicit _rn operations (no FMA contraction).
#pragma once
#define XB200_DEVICE_FUNCS_ONLY
#include "xb200_cabac.cuh"
#include "xb200_had.cuh"
#include "xb200_tq.cuh"

#define IN_CM_IPM XB200_CM_COUNT // the two ctx.intra_dir models follow the inter models in the shared-memory copy
#define IN_CM_N   (XB200_CM_COUNT + 2)
#define IN_MAX_COST 1.7e+308

template <int MAXN> struct IntraSmem {
    TqSmemT<MAXN> S;
    int16_t  org[MAXN * MAXN * 3 / 2];   // Y | U | V original block
    int16_t  levR[MAXN * MAXN];          // current luma levels, raster
    int16_t  levS[MAXN * MAXN];          // current luma levels, zig-zag order
    int16_t  bestS[MAXN * MAXN];         // best luma candidate, zig-zag order
    int16_t  chS[MAXN * MAXN / 2];       // chroma levels U | V, zig-zag order
    int16_t  nb[8 * MAXN + 8];           // per plane: left[-1 .. 2n-1], up[-1 .. 2n-1]
    uint16_t cm_base[IN_CM_N + 2], cm_run[IN_CM_N + 2];
    uint32_t range_base, range_run, bits;
    int32_t  list[5], pred_cnt;
    uint32_t cand_satd[5];
    double   cand_cost[5];
};

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

template <int MAXL2, int NT>
__global__ void __launch_bounds__(NT) k_intra(const PicDev *__restrict__ pics, xb200_intra_item *items, const int32_t *__restrict__ order,
                                               int cnt, const xb200_rates *__restrict__ rates, const xb200_sbac *__restrict__ st_in,
                                               xb200_sbac *__restrict__ st_out, const int16_t *__restrict__ side,
                                               int16_t *__restrict__ coef, int16_t *__restrict__ rec, const int8_t *__restrict__ g_tm64,
                                               SeqDev sq)
{
    constexpr int MAXN = 1 << MAXL2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    IntraSmem<MAXN> &M = *reinterpret_cast<IntraSmem<MAXN> *>(smem_raw);
    auto            &S = M.S;
    const int tid = threadIdx.x, lane = tid & 31, bd = sq.bd, maxv = (1 << bd) - 1;
    tq_load_tm(S, g_tm64, tid, NT);

    for(int ii = blockIdx.x; ii < cnt; ii += gridDim.x) {
        xb200_intra_item &it = items[order[ii]];
        const int l2 = it.log2_cuw, N = 1 << l2, ny = N * N, nc = ny >> 2, NC = N >> 1;
        const int slice_type = it.slice_type;
        const xb200_rates *rt = &rates[it.rate_idx];
        // ---- stage inputs: original block, reference samples, coder state ------------------------------------------
        {
            const PicDev p = pics[it.cur_pic];
            const int16_t *gy = p.p[0] + (ptrdiff_t)it.y * p.s[0] + it.x;
            for(int e = tid; e < ny; e += NT) M.org[e] = gy[(ptrdiff_t)(e >> l2) * p.s[0] + (e & (N - 1))];
            for(int c = 1; c < 3; c++) {
                const int16_t *gc = p.p[c] + (ptrdiff_t)(it.y >> 1) * p.s[c] + (it.x >> 1);
                for(int e = tid; e < nc; e += NT) M.org[ny + (c - 1) * nc + e] = gc[(ptrdiff_t)(e >> (l2 - 1)) * p.s[c] + (e & (NC - 1))];
            }
            const int16_t *gn = side + it.nb_off;
            for(int e = tid; e < 8 * N + 6; e += NT) M.nb[e] = gn[e];
            const xb200_sbac &s0 = st_in[it.state_in];
            for(int k = tid; k < XB200_CM_COUNT; k += NT) M.cm_base[k] = s0.m[k];
            if(tid == 0) {
                M.cm_base[IN_CM_IPM] = it.cm_ipm_in[0]; M.cm_base[IN_CM_IPM + 1] = it.cm_ipm_in[1];
                M.range_base = s0.range;
                for(int k = 0; k < 5; k++) { M.list[k] = 0; M.cand_cost[k] = IN_MAX_COST; M.cand_satd[k] = 0xffffffffu; }
            }
        }
        __syncthreads();
        const int16_t *leY = M.nb + 1, *upY = M.nb + (2 * N + 1) + 1;
        const int16_t *leC[2] = {M.nb + 2 * (2 * N + 1) + 1, M.nb + 2 * (2 * N + 1) + 2 * (N + 1) + 1};
        const int16_t *upC[2] = {leC[0] + (N + 1), leC[1] + (N + 1)};
        const int dcY = ipred_dc(leY, upY, l2);

        // ---- make_ipred_list: SATD + sqrt(lambda) * mode bits, insertion-sorted (src_base/xeve_pintra.c:308-374) ------
        for(int ipm = 0; ipm < 5; ipm++) {
            for(int e = tid; e < ny; e += NT) S.blk[e] = (int16_t)ipred_at(leY, upY, ipm, e >> l2, e & (N - 1), dcY);
            __syncthreads();
            int sum = 0;
            if(N >= 8) {
                const int tw = N >> 3;
                for(int t = tid; t < tw * tw; t += NT) {
                    const int o = (t / tw) * 8 * N + (t % tw) * 8;
                    sum += had_tile_dev<8>(M.org + o, N, S.blk + o, N);
                }
            }
            else if(tid == 0) sum = had_tile_dev<4>(M.org, N, S.blk, N);
            const uint32_t satd = (uint32_t)(block_sum_s32(S, sum, tid, NT) >> (bd - 8));
            if(tid == 0) {
                Cabac c;
                c.range = M.range_base; c.bits = 0; c.m = M.cm_run;
                M.cm_run[IN_CM_IPM] = M.cm_base[IN_CM_IPM]; M.cm_run[IN_CM_IPM + 1] = M.cm_base[IN_CM_IPM + 1];
                cb_unary(c, it.mpm[ipm], IN_CM_IPM);
                const double cost = __dadd_rn((double)satd, __dmul_rn((double)c.bits, it.sqrt_lambda0));
                int shift = 0;
                while(shift < 5 && cost < M.cand_cost[4 - shift]) shift++;
                if(shift) {
                    for(int j = 1; j < shift; j++) {
                        M.list[5 - j] = M.list[4 - j]; M.cand_cost[5 - j] = M.cand_cost[4 - j]; M.cand_satd[5 - j] = M.cand_satd[4 - j];
                    }
                    M.list[5 - shift] = ipm; M.cand_cost[5 - shift] = cost; M.cand_satd[5 - shift] = satd;
                }
            }
            __syncthreads();
        }
        if(tid == 0) {
            int pc = 5;
            const double thr = __dmul_rn((double)it.inter_satd, 1.2);
            for(int i = 4; i >= 1; i--) {
                if((double)M.cand_satd[i] > thr) pc--;
                else break;
            }
            M.pred_cnt = pc;
        }
        __syncthreads();
        const int pred_cnt = M.pred_cnt;

        // ---- luma RDO per surviving mode (pintra_residue_rdo mode 0, src_base/xeve_pintra.c:97-152) ---------------------
        double  cost = IN_MAX_COST;
        int     best_ipd = 0, nnz_best0 = 0;
        int32_t best_dist_y = 0;
        int16_t *g_coef = coef + it.out_off, *g_rec = rec ? rec + it.out_off : nullptr;
        for(int j = 0; j < pred_cnt; j++) {
            const int ipm = M.list[j];
            for(int e = tid; e < ny; e += NT)
                S.blk[e] = (int16_t)(M.org[e] - ipred_at(leY, upY, ipm, e >> l2, e & (N - 1), dcY));
            __syncthreads();
            fwd_dct(S, l2, bd, tid, NT);
            const int nnz = quant_block(S, l2, it.qp[0], it.lambda[0], 1, 0, slice_type, rt, bd, sq.rdoq, tid, NT);
            for(int e = tid; e < ny; e += NT) {
                const int16_t v = S.blk[e];
                M.levR[e] = v;
                M.levS[zz_of(e, l2)] = v;
            }
            for(int k = tid; k < IN_CM_N; k += NT) M.cm_run[k] = M.cm_base[k];
            __syncthreads();
            if(tid < 32) { // xeve_rdo_bit_cnt_cu_intra_luma, src_base/xeve_mode.c:81-119
                Cabac c;
                c.range = M.range_base; c.bits = 0; c.m = M.cm_run;
                if(lane == 0) {
                    if(slice_type != 2 && it.all_preds) {
                        cb_bin(c, XB200_CM_SKIP_FLAG + it.ctx_skip, 0);
                        cb_bin(c, XB200_CM_PRED_MODE + it.ctx_pred_mode, 1);
                    }
                    cb_unary(c, it.mpm[ipm], IN_CM_IPM);
                    cb_bin(c, XB200_CM_CBF_LUMA, nnz != 0);
                }
                if(nnz) cb_run_length_sm(c, M.levS, ny, nnz, 0, lane);
                if(lane == 0) M.bits = c.bits;
            }
            if(nnz) {
                dequant_block(S, l2, it.qp[0], bd, tid, NT);
                inv_dct(S, l2, bd, tid, NT);
            }
            int64_t ssd = 0;
            for(int e = tid; e < ny; e += NT) {
                const int     pr = ipred_at(leY, upY, ipm, e >> l2, e & (N - 1), dcY);
                const int16_t t = nnz ? (int16_t)(S.blk[e] + pr) : (int16_t)pr;
                const int     r = clip3i(0, maxv, t), d = r - M.org[e];
                S.blk[e] = (int16_t)r;
                ssd += (int64_t)((d * d) >> ((bd - 8) << 1));
            }
            ssd = block_sum_s64(S, ssd, tid, NT); // (contains the barriers that publish M.bits and S.blk)
            double        cost_t = (double)ssd;
            const int32_t dist_t = (int32_t)cost_t;
            cost_t = __dadd_rn(cost_t, __dmul_rn((double)M.bits, it.lambda[0]));
            if(cost_t < cost) {
                cost = cost_t; best_dist_y = dist_t; best_ipd = ipm; nnz_best0 = nnz;
                for(int e = tid; e < ny; e += NT) {
                    g_coef[e] = M.levR[e];
                    M.bestS[e] = M.levS[e];
                    if(g_rec) g_rec[e] = S.blk[e];
                }
            }
            __syncthreads();
        }

        // ---- chroma with the winning luma mode (pintra_residue_rdo mode 1, :153-270); its own bit count is never used ---
        int     nnzc[2] = {0, 0};
        int64_t ssdc[2] = {0, 0};
        for(int c = 1; c < 3; c++) {
            const int16_t *le = leC[c - 1], *up = upC[c - 1], *og = M.org + ny + (c - 1) * nc;
            const int      dc = ipred_dc(le, up, l2 - 1);
            for(int e = tid; e < nc; e += NT) S.blk[e] = (int16_t)(og[e] - ipred_at(le, up, best_ipd, e >> (l2 - 1), e & (NC - 1), dc));
            __syncthreads();
            fwd_dct(S, l2 - 1, bd, tid, NT);
            const int nz = quant_block(S, l2 - 1, it.qp[c], it.lambda[c], 1, c, slice_type, rt, bd, sq.rdoq, tid, NT);
            nnzc[c - 1] = nz;
            for(int e = tid; e < nc; e += NT) {
                const int16_t v = S.blk[e];
                g_coef[ny + (c - 1) * nc + e] = v;
                M.chS[(c - 1) * nc + zz_of(e, l2 - 1)] = v;
            }
            __syncthreads();
            if(nz) {
                dequant_block(S, l2 - 1, it.qp[c], bd, tid, NT);
                inv_dct(S, l2 - 1, bd, tid, NT);
            }
            int64_t ssd = 0;
            for(int e = tid; e < nc; e += NT) {
                const int     pr = ipred_at(le, up, best_ipd, e >> (l2 - 1), e & (NC - 1), dc);
                const int16_t t = nz ? (int16_t)(S.blk[e] + pr) : (int16_t)pr;
                const int     r = clip3i(0, maxv, t), d = r - og[e];
                if(g_rec) g_rec[ny + (c - 1) * nc + e] = (int16_t)r;
                ssd += (int64_t)((d * d) >> ((bd - 8) << 1));
            }
            ssdc[c - 1] = block_sum_s64(S, ssd, tid, NT);
        }
        const int32_t best_dist_c = (int32_t)__dadd_rn(__dmul_rn(it.dist_chroma_weight[0], (double)ssdc[0]),
                                                       __dmul_rn(it.dist_chroma_weight[1], (double)ssdc[1]));

        // ---- final bit count of the CU from the input state (xeve_rdo_bit_cnt_cu_intra, src_base/xeve_mode.c:141-171) ----
        for(int k = tid; k < IN_CM_N; k += NT) M.cm_run[k] = M.cm_base[k];
        __syncthreads();
        if(tid < 32) {
            Cabac c;
            c.range = M.range_base; c.bits = 0; c.m = M.cm_run;
            if(lane == 0) {
                if(slice_type != 2) {
                    cb_bin(c, XB200_CM_SKIP_FLAG + it.ctx_skip, 0);
                    cb_bin(c, XB200_CM_PRED_MODE + it.ctx_pred_mode, 1);
                }
                cb_unary(c, it.mpm[best_ipd], IN_CM_IPM);
                cb_bin(c, XB200_CM_CBF_CB, nnzc[0] != 0);
                cb_bin(c, XB200_CM_CBF_CR, nnzc[1] != 0);
                cb_bin(c, XB200_CM_CBF_LUMA, nnz_best0 != 0);
            }
            if(nnz_best0) cb_run_length_sm(c, M.bestS, ny, nnz_best0, 0, lane);
            if(nnzc[0]) cb_run_length_sm(c, M.chS, nc, nnzc[0], 1, lane);
            if(nnzc[1]) cb_run_length_sm(c, M.chS + nc, nc, nnzc[1], 2, lane);
            if(lane == 0) { M.bits = c.bits; M.range_run = c.range; }
        }
        __syncthreads();
        if(tid == 0) {
            double ct = __dmul_rn((double)M.bits, it.lambda[0]);
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
            for(int k = tid; k < XB200_CM_COUNT; k += NT) so.m[k] = M.cm_run[k];
        }
        __syncthreads();
    }
}
