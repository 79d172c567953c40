// xb200_pipeline.cuh -- xeve_pinter_analyze_cu as a pipeline of frame-wide grids.
//
// The fused kernel of xb200_analyze.cuh walks one CU through the whole decision sequence with one team; at ~50 k dynamic
// instructions per CU over ~26 k static ones it is bound by instruction fetch (ncu: stall_no_instruction 46 per issue,
// profiles/).  Here the same sequence is cut at the points where the reference calls its kernels, and every segment runs
// for ALL CUs of the work list at once, so that each grid executes one compact piece of code:
//
//   k_cu_skip  ->  [k_cu_make_me_uni -> xb200_me]  ->  k_cu_after_uni  ->  [xb200_residue DIR|L0|L1]  ->  k_cu_decide
//              ->  4 x ( k_cu_bi_prep -> [xb200_bi_org -> xb200_me] -> k_cu_bi_update )            (analyze_bi, BI_ITER = 4)
//              ->  k_cu_bi_emit -> [xb200_residue BI] -> k_cu_decide  ->  k_cu_final
//
// The bracketed stages are the work-list operators themselves (device-resident item arrays in fixed slots, empty slots are
// skipped by the size-binning kernels).  The per-CU state that the reference keeps in XEVE_PINTER between those calls lives
// in CuState records in HBM; candidate modes keep {coef, rec, pred} in mode-major scratch planes.
#pragma once
#include "xb200_analyze.cuh"

struct CuState {
    CuMode     md[5];                             // PRED_L0, L1, BI, SKIP, DIR
    double     cost[5];
    xb200_sbac st[5];                             // s_temp_best of each mode
    int16_t    mv_scale[2][XB200_MAX_REFP][2];    // pi->mv_scale
    int32_t    mot_bits[2];                       // pi->mot_bits
    uint32_t   bi_best_me;
    int8_t     bi_refi[2];                        // refi[] of analyze_bi
    int8_t     lidx_ref, refi_best, active, bi_active, num_refp_cur, pad0;
    uint8_t    mvp_idx[2], pad1[2];
};
// scratch plane k (0 coef, 1 rec, 2 pred) of mode m starts at ((m * 3 + k) * elems) s16 elements
XB_DEV int64_t cu_plane(int m, int k, int64_t elems) { return (int64_t)(m * 3 + k) * elems; }

template <int L2> struct SkipCfg {
    using R = Res2Cfg<L2>;
    static constexpr int T = R::T, TEAMS = R::TEAMS, CTA = R::CTA, N = R::N, NY = N * N, NCH = NY >> 2, NP = R::PRED;
    static constexpr int HDR = ((int)sizeof(CuHdr) + 15) & ~15;
    static constexpr int TEAM_BYTES = HDR + (2 * NP + (N + 8) * N) * 2;
    static constexpr int SMEM = TEAMS * TEAM_BYTES;
};

// xeve_analyze_skip (src_base/xeve_pinter.c:1337-1530): one team per CU
template <int L2>
__global__ void __launch_bounds__(SkipCfg<L2>::CTA) k_cu_skip(const PicDev *__restrict__ pics, const xb200_cu_item *__restrict__ items,
                                                              const int32_t *__restrict__ order, int n, const xb200_sbac *__restrict__ st_in,
                                                              CuState *__restrict__ states, int16_t *__restrict__ scratch, int64_t elems, SeqDev sq)
{
    using Cf = SkipCfg<L2>;
    constexpr int T = Cf::T, N = Cf::N, NY = Cf::NY, NCH = Cf::NCH, NP = Cf::NP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int      team = threadIdx.x / T, tt = threadIdx.x % T;
    unsigned char *tb = smem_raw + (size_t)team * Cf::TEAM_BYTES;
    CuHdr         &H = *reinterpret_cast<CuHdr *>(tb);
    int16_t       *pred = reinterpret_cast<int16_t *>(tb + Cf::HDR), *aux = pred + NP, *tmp = aux + NP;
    const int      sh = (sq.bd - 8) << 1;
    for(int i = blockIdx.x * Cf::TEAMS + team; i < n; i += gridDim.x * Cf::TEAMS) {
        const int            ci = order[i];
        const xb200_cu_item *git = &items[ci];
        {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(git);
            uint32_t       *dst = reinterpret_cast<uint32_t *>(&H.cu);
            for(int e = tt; e < (int)(sizeof(xb200_cu_item) / 4); e += T) dst[e] = src[e];
            const xb200_sbac &s = st_in[git->state_in];
            for(int k = tt; k < XB200_CM_COUNT; k += T) H.st[ST_IN][k] = s.m[k];
            if(tt == 0) H.rg[ST_IN] = s.range;
        }
        team_sync<T>();
        const xb200_cu_item &cu = H.cu;
        CuState             &S = states[ci];
        const PicDev        &o = pics[cu.cur_pic];
        const int16_t       *org0 = o.p[0] + (ptrdiff_t)cu.y * o.s[0] + cu.x;
        const int16_t       *org1 = o.p[1] + (ptrdiff_t)(cu.y >> 1) * o.s[1] + (cu.x >> 1);
        const int16_t       *org2 = o.p[2] + (ptrdiff_t)(cu.y >> 1) * o.s[2] + (cu.x >> 1);
        int16_t             *gpred = scratch + cu_plane(3, 2, elems) + cu.out_off;
        const bool           B = cu.slice_type == 0;
        const double         w0 = cu.dist_chroma_weight[0], w1 = cu.dist_chroma_weight[1];
        int64_t              best_ssd = (int64_t)1 << (2 * L2 + 16);
        double               sb = CU_MAX_COST;
        for(int idx0 = 0; idx0 < sq.merge_num; idx0++) {
            bool dup = false;
            for(int t = idx0 - 1; t >= 0; t--) dup |= cu.mvp[0][t][0] == cu.mvp[0][idx0][0] && cu.mvp[0][t][1] == cu.mvp[0][idx0][1];
            if(dup) continue;
            const int cnt = B ? sq.merge_num : 1;
            for(int idx1 = 0; idx1 < cnt; idx1++) {
                dup = false;
                for(int t = idx1 - 1; t >= 0; t--) dup |= cu.mvp[1][t][0] == cu.mvp[1][idx1][0] && cu.mvp[1][t][1] == cu.mvp[1][idx1][1];
                if(dup) continue;
                const int r0 = cu.refi_pred[0][idx0], r1 = B ? cu.refi_pred[1][idx1] : -1;
                if(r0 < 0 && r1 < 0) continue;
                cu_predict<L2>(pics, cu, sq, r0, r1, cu.mvp[0][idx0][0], cu.mvp[0][idx0][1], cu.mvp[1][idx1][0], cu.mvp[1][idx1][1], pred, aux, tmp, tt);
                const int64_t cy = ssd_plane_t<L2, T>(org0, o.s[0], pred, sh, tt, H.X);
                const int64_t cb = ssd_plane_t<L2 - 1, T>(org1, o.s[1], pred + NY, sh, tt, H.X);
                const int64_t cr = ssd_plane_t<L2 - 1, T>(org2, o.s[2], pred + NY + NCH, sh, tt, H.X);
                xb200_bits_item bi = cu_bits_item(cu, 0, 3, 0);
                bi.mvp_idx[0] = (uint8_t)idx0; bi.mvp_idx[1] = (uint8_t)idx1;
                const uint32_t bits = cu_count<T>(H, bi, nullptr, ST_IN, tt);
                double cost = __dadd_rn(__dadd_rn(__ll2double_rn(cy), __dmul_rn(w0, __ll2double_rn(cb))), __dmul_rn(w1, __ll2double_rn(cr)));
                cost = __dadd_rn(cost, __dmul_rn((double)bits, cu.lambda[0]));
                if(cost < sb) {
                    sb = cost;
                    best_ssd = cy + cb + cr;
                    for(int e = tt; e < NP; e += T) gpred[e] = pred[e];
                    if(tt == 0) {
                        CuMode &M = S.md[3];
                        M.mvp_idx[0] = (uint8_t)idx0; M.mvp_idx[1] = (uint8_t)idx1;
                        M.refi[0] = (int8_t)r0; M.refi[1] = (int8_t)r1;
                        M.mv[0][0] = cu.mvp[0][idx0][0]; M.mv[0][1] = cu.mvp[0][idx0][1]; M.mv[1][0] = cu.mvp[1][idx1][0]; M.mv[1][1] = cu.mvp[1][idx1][1];
                        M.mvd[0][0] = M.mvd[0][1] = M.mvd[1][0] = M.mvd[1][1] = 0;
                        M.nnz[0] = M.nnz[1] = M.nnz[2] = 0; M.cbf = 0;
                    }
                    cu_st_save<T>(H, ST_MODE, ST_RUN, tt);
                }
            }
        }
        if(tt < 32) {
            for(int k = tt; k < XB200_CM_COUNT; k += 32) S.st[3].m[k] = H.st[ST_MODE][k];
            if(tt == 0) {
                S.st[3].range = H.rg[ST_MODE];
                S.cost[0] = S.cost[1] = S.cost[2] = S.cost[4] = CU_MAX_COST;
                S.cost[3] = sb;
                S.active = (sb < CU_MAX_COST && best_ssd > 0) ? 1 : 0;   // skip_th = 0: anything but a perfect skip goes on (quirk q2)
                S.bi_active = 0; S.mot_bits[0] = S.mot_bits[1] = 0; S.mvp_idx[0] = S.mvp_idx[1] = 0; S.num_refp_cur = 0;
            }
        }
        team_sync<T>();
    }
}

// ---- uni-directional search: one fn_me slot per (CU, list, reference) ------------------------------------------------
__global__ void k_cu_make_me_uni(const xb200_cu_item *__restrict__ items, int n, const CuState *__restrict__ states, xb200_me_item *__restrict__ me,
                                 SeqDev sq)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if(j >= n * 8) return;
    const int            ci = j >> 3, lidx = (j >> 2) & 1, r = j & 3;
    const xb200_cu_item &cu = items[ci];
    const CuState       &S = states[ci];
    xb200_me_item        m;
    memset(&m, 0, sizeof(m));
    if(S.active && (lidx == 0 || cu.slice_type == 0) && r < cu.num_refp[lidx]) {
        const int mi = S.md[3].mvp_idx[lidx];
        m.poc = cu.poc; m.cur_pic = cu.cur_pic; m.ref_pic = cu.ref_pic[lidx][r]; m.ref_poc = cu.ref_poc[lidx][r];
        m.x = cu.x; m.y = cu.y; m.log2_cuw = cu.log2_cuw; m.log2_cuh = cu.log2_cuh; m.lidx = (uint8_t)lidx; m.bi = 0;
        m.refi = (int8_t)r; m.num_refp = cu.num_refp[lidx];
        m.mvp[0] = cu.mvp[lidx][mi][0]; m.mvp[1] = cu.mvp[lidx][mi][1];
        m.lambda_mv = cu.lambda_mv; m.max_search_range = cu.max_search_range; m.gop_size = sq.gop_size; m.org_bi_off = -1;
    }
    me[j] = m;
}

XB_DEV void cu_emit_residue(xb200_residue_item &it, const xb200_cu_item &cu, const CuMode &M, int pidx, int64_t elems)
{
    const int8_t  refi[2] = {M.refi[0], M.refi[1]};
    const int16_t mv[2][2] = {{M.mv[0][0], M.mv[0][1]}, {M.mv[1][0], M.mv[1][1]}};
    cu_mc_item(cu, 1 << cu.log2_cuw, refi, mv, it.mc);
    it.cur_pic = cu.cur_pic; it.slice_type = cu.slice_type; it.run_stats = 7;
    it.qp[0] = cu.qp[0]; it.qp[1] = cu.qp[1]; it.qp[2] = cu.qp[2]; it.pad_[0] = it.pad_[1] = it.pad_[2] = 0;
    it.rate_idx = cu.rate_idx;
    it.lambda[0] = cu.lambda[0]; it.lambda[1] = cu.lambda[1]; it.lambda[2] = cu.lambda[2];
    it.out_off = cu_plane(pidx, 0, elems) + cu.out_off;
}

constexpr int PIPE_WARPS = 4;
XB_DEV void cu_load_hdr(CuHdr &H, const xb200_cu_item *git, const xb200_sbac *st_in, int lane)
{
    const uint32_t *src = reinterpret_cast<const uint32_t *>(git);
    uint32_t       *dst = reinterpret_cast<uint32_t *>(&H.cu);
    for(int e = lane; e < (int)(sizeof(xb200_cu_item) / 4); e += 32) dst[e] = src[e];
    const xb200_sbac &s = st_in[git->state_in];
    for(int k = lane; k < XB200_CM_COUNT; k += 32) H.st[ST_IN][k] = s.m[k];
    if(lane == 0) H.rg[ST_IN] = s.range;
    __syncwarp();
}

// after the uni-directional searches: best reference per list, check_best_mvp (src_base/xeve_pinter.c:1772-1837), the DIRECT
// candidate, and the residue work items of DIR | L0 | L1 (slot 3 * cu + {0, 1, 2}); one warp per CU
__global__ void __launch_bounds__(PIPE_WARPS * 32) k_cu_after_uni(const xb200_cu_item *__restrict__ items, int n, const xb200_sbac *__restrict__ st_in,
                                                                  CuState *__restrict__ states, const xb200_me_item *__restrict__ me,
                                                                  xb200_residue_item *__restrict__ res, int64_t elems)
{
    __shared__ CuHdr hdr[PIPE_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ci = blockIdx.x * PIPE_WARPS + w;
    if(ci >= n) return;
    CuHdr    &H = hdr[w];
    CuState  &S = states[ci];
    xb200_residue_item *slot = res + (size_t)3 * ci;
    if(!S.active) {
        if(lane < 3) slot[lane].mc.w = slot[lane].mc.h = 0;
        return;
    }
    cu_load_hdr(H, &items[ci], st_in, lane);
    const xb200_cu_item &cu = H.cu;
    const bool           B = cu.slice_type == 0;
    uint8_t              mvp_idx[2] = {0, 0};
    for(int lidx = 0; lidx <= (B ? 1 : 0); lidx++) {
        const xb200_me_item *mi = me + (size_t)ci * 8 + lidx * 4;
        const int            nr = min((int)cu.num_refp[lidx], XB200_MAX_REFP);
        uint32_t             best_me = 0xffffffffu;
        int                  refi_t = 0;
        for(int r = 0; r < nr; r++)
            if(mi[r].cost < best_me) { best_me = mi[r].cost; refi_t = r; }
        const int mvx = mi[refi_t].mv_out[0], mvy = mi[refi_t].mv_out[1];
        mvp_idx[lidx] = S.md[3].mvp_idx[lidx];
        const int16_t(*cand)[2] = cu.mvp[lidx];
        xb200_bits_item bi = cu_bits_item(cu, 2, lidx, 0);
        bi.refi[lidx] = (int8_t)refi_t;
        bi.mvp_idx[0] = mvp_idx[lidx];
        bi.mvd[lidx][0] = (int16_t)(mvx - cand[mvp_idx[lidx]][0]); bi.mvd[lidx][1] = (int16_t)(mvy - cand[mvp_idx[lidx]][1]);
        const double ref_cost = __dmul_rn((double)cu_count<32>(H, bi, nullptr, ST_IN, lane), cu.lambda[0]);
        int          best = mvp_idx[lidx];
        for(int idx = 0; idx < 4; idx++) {   // the reference cost is never updated inside the loop (quirk q1)
            bool dup = false;
            for(int t = idx - 1; t >= 0; t--) dup |= cand[idx][0] == cand[t][0] && cand[idx][1] == cand[t][1];
            if(dup) continue;
            bi.mvp_idx[0] = (uint8_t)idx;
            bi.mvd[lidx][0] = (int16_t)(mvx - cand[idx][0]); bi.mvd[lidx][1] = (int16_t)(mvy - cand[idx][1]);
            const double c = __dmul_rn((double)cu_count<32>(H, bi, nullptr, ST_IN, lane), cu.lambda[0]);
            if(c < ref_cost) best = idx;
        }
        mvp_idx[lidx] = (uint8_t)best;
        if(lane == 0) {
            CuMode &M = S.md[lidx];
            M.refi[lidx] = (int8_t)refi_t; M.refi[1 - lidx] = -1;
            M.mv[lidx][0] = (int16_t)mvx; M.mv[lidx][1] = (int16_t)mvy; M.mv[1 - lidx][0] = M.mv[1 - lidx][1] = 0;
            M.mvd[lidx][0] = (int16_t)(mvx - cand[best][0]); M.mvd[lidx][1] = (int16_t)(mvy - cand[best][1]);
            M.mvd[1 - lidx][0] = M.mvd[1 - lidx][1] = 0;
            M.mvp_idx[lidx] = (uint8_t)best; M.mvp_idx[1 - lidx] = 0;
            for(int r = 0; r < nr; r++) { S.mv_scale[lidx][r][0] = mi[r].mv_out[0]; S.mv_scale[lidx][r][1] = mi[r].mv_out[1]; }
            S.mot_bits[lidx] = mi[nr - 1].mot_bits_out[lidx];
            S.mvp_idx[lidx] = (uint8_t)best;
            S.num_refp_cur = (int8_t)nr;
            cu_emit_residue(slot[1 + lidx], cu, M, lidx, elems);
        }
    }
    if(lane == 0) {
        if(B) {
            CuMode &M = S.md[4];
            M.refi[0] = M.refi[1] = 0; M.mvp_idx[0] = M.mvp_idx[1] = 0;
            M.mv[0][0] = cu.mv_dir[0][0]; M.mv[0][1] = cu.mv_dir[0][1]; M.mv[1][0] = cu.mv_dir[1][0]; M.mv[1][1] = cu.mv_dir[1][1];
            M.mvd[0][0] = M.mvd[0][1] = M.mvd[1][0] = M.mvd[1][1] = 0;
            cu_emit_residue(slot[0], cu, M, 4, elems);
        }
        else { slot[0].mc.w = slot[0].mc.h = 0; slot[2].mc.w = slot[2].mc.h = 0; }
    }
}

// cbf decisions + RD cost of every evaluated residue candidate: one warp per residue slot.  per_cu = 3: slots DIR | L0 | L1,
// per_cu = 1: the BI candidate
__global__ void __launch_bounds__(PIPE_WARPS * 32) k_cu_decide(const xb200_cu_item *__restrict__ items, int n_slots, int per_cu,
                                                               const xb200_sbac *__restrict__ st_in, CuState *__restrict__ states,
                                                               const xb200_residue_item *__restrict__ res, const int16_t *__restrict__ scratch,
                                                               int w_lo, int w_hi)
{
    __shared__ CuHdr hdr[PIPE_WARPS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int sl = blockIdx.x * PIPE_WARPS + w;
    if(sl >= n_slots) return;
    const xb200_residue_item &it = res[sl];
    if(it.mc.w < w_lo || it.mc.w > w_hi) return;   // empty slot (w = 0) or a size class handled by the lane coder
    const int ci = sl / per_cu, k = sl - ci * per_cu;
    const int pidx = per_cu == 1 ? 2 : (k == 0 ? 4 : k - 1);
    CuHdr    &H = hdr[w];
    CuState  &S = states[ci];
    cu_load_hdr(H, &items[ci], st_in, lane);
    const int     store[3] = {it.nnz[0], it.nnz[1], it.nnz[2]};
    const int64_t d0[3] = {it.dist_pred[0], it.dist_pred[1], it.dist_pred[2]}, d1[3] = {it.dist_rec[0], it.dist_rec[1], it.dist_rec[2]};
    const CuMode  M = S.md[pidx];
    int           cbf;
    const double  best = cu_cbf_decide<32>(H, pidx, M, M.mvp_idx[0], M.mvp_idx[1], store, d0, d1, scratch + it.out_off, cbf, lane);
    for(int q = lane; q < XB200_CM_COUNT; q += 32) S.st[pidx].m[q] = H.st[ST_MODE][q];
    if(lane == 0) {
        S.st[pidx].range = H.rg[ST_MODE];
        S.cost[pidx] = best;
        S.md[pidx].cbf = cbf;
        S.md[pidx].nnz[0] = (cbf & 1) ? store[0] : 0; S.md[pidx].nnz[1] = (cbf & 2) ? store[1] : 0; S.md[pidx].nnz[2] = (cbf & 4) ? store[2] : 0;
    }
}

// ---- analyze_bi (src_base/xeve_pinter.c:1567-1683), one pass of its BI_ITER loop per launch pair ---------------------------
__global__ void k_cu_bi_prep(const xb200_cu_item *__restrict__ items, int n, CuState *__restrict__ states, int iter, xb200_mc_item *__restrict__ mc,
                             int32_t *__restrict__ cur_pic, int64_t *__restrict__ side_off, xb200_me_item *__restrict__ me, SeqDev sq)
{
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if(ci >= n) return;
    const xb200_cu_item &cu = items[ci];
    CuState             &S = states[ci];
    if(iter == 0) {
        S.bi_active = (S.active && cu.slice_type == 0) ? 1 : 0;
        if(S.bi_active) {
            const int lr = S.cost[0] <= S.cost[1] ? 0 : 1;
            CuMode   &M = S.md[2];
            M.mvp_idx[0] = S.md[0].mvp_idx[0]; M.mvp_idx[1] = S.md[1].mvp_idx[1];
            M.refi[0] = S.md[0].refi[0]; M.refi[1] = S.md[1].refi[1];
            M.mv[0][0] = S.md[0].mv[0][0]; M.mv[0][1] = S.md[0].mv[0][1]; M.mv[1][0] = S.md[1].mv[1][0]; M.mv[1][1] = S.md[1].mv[1][1];
            S.lidx_ref = (int8_t)lr;
            S.bi_refi[lr] = M.refi[lr]; S.bi_refi[1 - lr] = -1;
            S.bi_best_me = 0xffffffffu; S.refi_best = 0;
        }
    }
    xb200_mc_item m;
    memset(&m, 0, sizeof(m));
    xb200_me_item e[XB200_MAX_REFP];
    memset(e, 0, sizeof(e));
    cur_pic[ci] = cu.cur_pic;
    side_off[ci] = cu.out_off / 3 * 2;   // luma block of the CU inside a buffer of 2/3 * elems samples
    if(S.bi_active) {
        const CuMode &M = S.md[2];
        const int8_t  refi[2] = {S.bi_refi[0], S.bi_refi[1]};
        const int16_t mv[2][2] = {{M.mv[0][0], M.mv[0][1]}, {M.mv[1][0], M.mv[1][1]}};
        cu_mc_item(cu, 1 << cu.log2_cuw, refi, mv, m);
        // SWAP(refi[lidx_ref], refi[lidx_cnd]); SWAP(lidx_ref, lidx_cnd)
        const int lr = 1 - S.lidx_ref;
        S.lidx_ref = (int8_t)lr;
        { const int8_t t = S.bi_refi[0]; S.bi_refi[0] = S.bi_refi[1]; S.bi_refi[1] = t; }
        const int mi = M.mvp_idx[lr];
        for(int r = 0; r < S.num_refp_cur; r++) {
            xb200_me_item &q = e[r];
            q.poc = cu.poc; q.cur_pic = cu.cur_pic; q.ref_pic = cu.ref_pic[lr][r]; q.ref_poc = cu.ref_poc[lr][r];
            q.x = cu.x; q.y = cu.y; q.log2_cuw = cu.log2_cuw; q.log2_cuh = cu.log2_cuh; q.lidx = (uint8_t)lr; q.bi = 1;
            q.refi = (int8_t)r; q.num_refp = (uint8_t)S.num_refp_cur;
            q.mvp[0] = cu.mvp[lr][mi][0]; q.mvp[1] = cu.mvp[lr][mi][1];
            q.mv_in[0] = S.mv_scale[lr][r][0]; q.mv_in[1] = S.mv_scale[lr][r][1];
            q.lambda_mv = cu.lambda_mv; q.mot_bits_in[0] = S.mot_bits[0]; q.mot_bits_in[1] = S.mot_bits[1];
            q.max_search_range = cu.max_search_range; q.gop_size = sq.gop_size; q.org_bi_off = (int32_t)side_off[ci];
        }
    }
    mc[ci] = m;
    for(int r = 0; r < XB200_MAX_REFP; r++) me[(size_t)ci * XB200_MAX_REFP + r] = e[r];
}

__global__ void k_cu_bi_update(int n, CuState *__restrict__ states, const xb200_me_item *__restrict__ me, int *__restrict__ n_active)
{
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if(ci >= n) return;
    CuState &S = states[ci];
    if(!S.bi_active) return;
    const int lr = S.lidx_ref;
    CuMode   &M = S.md[2];
    bool      changed = false;
    for(int r = 0; r < S.num_refp_cur; r++) {
        const xb200_me_item &q = me[(size_t)ci * XB200_MAX_REFP + r];
        S.mv_scale[lr][r][0] = q.mv_out[0]; S.mv_scale[lr][r][1] = q.mv_out[1];
        if(q.cost < S.bi_best_me) {
            S.refi_best = (int8_t)r; S.bi_best_me = q.cost; changed = true;
            M.refi[lr] = (int8_t)r;
            M.mv[lr][0] = q.mv_out[0]; M.mv[lr][1] = q.mv_out[1];
        }
    }
    S.bi_refi[lr] = S.refi_best; S.bi_refi[1 - lr] = -1;
    if(!changed) S.bi_active = 0;
    else atomicAdd(n_active, 1);
}

__global__ void k_cu_bi_emit(const xb200_cu_item *__restrict__ items, int n, CuState *__restrict__ states, xb200_residue_item *__restrict__ res,
                             int64_t elems)
{
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    if(ci >= n) return;
    const xb200_cu_item &cu = items[ci];
    CuState             &S = states[ci];
    xb200_residue_item   it;
    memset(&it, 0, sizeof(it));
    if(S.active && cu.slice_type == 0) {
        CuMode &M = S.md[2];
        for(int l = 0; l < 2; l++) {
            M.mvd[l][0] = (int16_t)(M.mv[l][0] - cu.mvp[l][M.mvp_idx[l]][0]);
            M.mvd[l][1] = (int16_t)(M.mv[l][1] - cu.mvp[l][M.mvp_idx[l]][1]);
        }
        cu_emit_residue(it, cu, M, 2, elems);
    }
    res[ci] = it;
}

// winner of the sequential strict-less comparison chain SKIP, DIR, L0, L1, BI; coefficients with dropped planes zeroed,
// reconstruction (prediction where a plane was dropped), XEVE_MODE fields, s_next_best.  One CTA per CU.
__global__ void __launch_bounds__(128) k_cu_final(xb200_cu_item *__restrict__ items, int n, const CuState *__restrict__ states,
                                                  const int16_t *__restrict__ scratch, int64_t elems, xb200_sbac *__restrict__ st_out,
                                                  int16_t *__restrict__ coef_out, int16_t *__restrict__ rec_out)
{
    const int ci = blockIdx.x;
    if(ci >= n) return;
    xb200_cu_item *git = &items[ci];
    const CuState &S = states[ci];
    int            best = 3;
    double         cb = S.cost[3];
    if(S.active) {
        const bool B = git->slice_type == 0;
        if(B && S.cost[4] < cb) { cb = S.cost[4]; best = 4; }
        if(S.cost[0] < cb) { cb = S.cost[0]; best = 0; }
        if(B && S.cost[1] < cb) { cb = S.cost[1]; best = 1; }
        if(B && S.cost[2] < cb) { cb = S.cost[2]; best = 2; }
    }
    const CuMode &M = S.md[best];
    const int     l2 = git->log2_cuw, ny = 1 << (2 * l2), nch = ny >> 2, np = ny + 2 * nch;
    const int64_t oo = git->out_off;
    const int16_t *sc = scratch + cu_plane(best, 0, elems) + oo, *sr = scratch + cu_plane(best, 1, elems) + oo,
                  *sp = scratch + cu_plane(best, 2, elems) + oo;
    const int cbf = best == 3 ? 0 : M.cbf;
    for(int e = threadIdx.x; e < np; e += blockDim.x) {
        const int  c = e < ny ? 0 : (e < ny + nch ? 1 : 2);
        const bool on = (cbf >> c) & 1;
        coef_out[oo + e] = on ? sc[e] : (int16_t)0;
        if(rec_out) rec_out[oo + e] = on ? sr[e] : sp[e];
    }
    if(threadIdx.x == 0) {
        git->cost = cb; git->best_idx = (uint8_t)best;
        for(int l = 0; l < 2; l++) {
            git->refi[l] = M.refi[l]; git->mvp_idx[l] = M.mvp_idx[l];
            git->mv[l][0] = M.mv[l][0]; git->mv[l][1] = M.mv[l][1]; git->mvd[l][0] = M.mvd[l][0]; git->mvd[l][1] = M.mvd[l][1];
        }
        git->nnz[0] = best == 3 ? 0 : M.nnz[0]; git->nnz[1] = best == 3 ? 0 : M.nnz[1]; git->nnz[2] = best == 3 ? 0 : M.nnz[2];
    }
    if(git->state_out >= 0 && threadIdx.x < 32) {
        xb200_sbac &so = st_out[git->state_out];
        for(int k = threadIdx.x; k < XB200_CM_COUNT; k += 32) so.m[k] = S.st[best].m[k];
        if(threadIdx.x == 0) so.range = S.st[best].range;
    }
}
