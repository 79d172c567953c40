// xb200_analyze_par.cuh -- xeve_pinter_analyze_cu of one 8x8 / 16x16 CU on THREE warps (decision-chain kernel only).
//
// Inside a coder-state chain the CUs come one after the other, so the only parallelism left is inside a CU.  The candidate modes of
// xeve_pinter_analyze_cu (src_base/xeve_pinter.c:1839-2056) are independent until they are compared: every skip candidate, the direct
// mode and each uni-directional list start from the SAME input coder state (SBAC_LOAD(s_temp_run, s_curr_best), :944, :1376) and write
// different result slots.  A warp-synchronous team (T = 32) of xb200_analyze.cuh is relocatable -- it only uses __syncwarp -- so the
// same device functions (cu_predict, cu_me, cu_count, cu_residue_rdo) run here on three warps with private headers / working sets:
//
//   phase 1   skip candidates, round-robin over the warps             -> ordered arg-min (first minimum, reference order)
//   phase 2   warp 0: direct mode RDO | warp 1: list 0 search + check_best_mvp + RDO | warp 2: the same for list 1
//   merge     costs compared in the reference's order (skip, direct, L0, L1), strict less
//   phase 3   warp 0: the bi-prediction loop (needs both uni results) + its RDO, then the winner's copy-out
//
// Results are bit-identical to analyze_cu_one (tests compare both against the oracle and the reference).
#pragma once
#include "xb200_analyze.cuh"

constexpr int CU_PAR_WARPS = 3;
struct CuPar {                         // exchange area of the three warps (lives in warp 0's region)
    double    cost[4];                 // [0] direct, [1] L0, [2] L1 RD costs; [3] unused
    double    skip_cost[16];           // per skip candidate pair (idx0 * 4 + idx1), CU_MAX_COST when not evaluated
    long long skip_ssd[16];
    int32_t   mot_bits[2];
    int16_t   mv_scale[2][XB200_MAX_REFP][2];
};
constexpr int CU_PAR_ORGBI = 1024;     // 16x16 bi-search target
XB_DEV void cu_par_sync() { asm volatile("bar.sync 2, %0;" ::"n"(CU_PAR_WARPS * 32) : "memory"); }
#define CU_PAR_SYNC(id) do { if(lane == 0) CH_DBG(8 + w, (id)); cu_par_sync(); } while(0)
// bytes one warp's region needs: header | exchange area | org_bi | max(search working set, residue working set)
template <int L2> __host__ __device__ inline size_t cu_par_region_bytes(int win_cap)
{
    const size_t me = me_team_bytes(L2, win_cap), rs = 16 + (size_t)Res2Cfg<L2>::TEAM_BYTES;
    return (((size_t)sizeof(CuHdr) + 15) & ~(size_t)15) + (((size_t)sizeof(CuPar) + 15) & ~(size_t)15) + CU_PAR_ORGBI + (((me > rs ? me : rs) + 15) & ~(size_t)15);
}

template <int L2>
__device__ __noinline__ void analyze_cu_par(unsigned char *base, int stride, uint64_t *bars, uint32_t *bphases, const int8_t *tm, const int8_t *tmT, const PicDev *__restrict__ pics,
                                            xb200_cu_item *git, const xb200_rates *rates, const xb200_sbac *st_in, xb200_sbac *st_out,
                                            int16_t *coef_out, int16_t *rec_out, int16_t *pred_y_out, int16_t *scratch, const SeqDev &sq, int win_cap,
                                            int *err_flag, int tt)
{
    using Cf = CuCfg<L2>;
    constexpr int T = 32, NY = Cf::NY, NCH = Cf::NCH, NP = Cf::NP, N = Cf::N;
    constexpr int HDR = ((int)sizeof(CuHdr) + 15) & ~15, PAR = ((int)sizeof(CuPar) + 15) & ~15;
    const int w = tt >> 5, lane = tt & 31;
    CuHdr *Hs[CU_PAR_WARPS];
#pragma unroll
    for(int k = 0; k < CU_PAR_WARPS; k++) Hs[k] = reinterpret_cast<CuHdr *>(base + (size_t)k * stride);
    CuPar &P = *reinterpret_cast<CuPar *>(base + HDR);
    unsigned char *reg = base + (size_t)w * stride;
    CuTeam<L2>     Tm;
    Tm.H = Hs[w];
    Tm.org_bi = reinterpret_cast<int16_t *>(reg + HDR + PAR);
    Tm.me_area = reg + HDR + PAR + CU_PAR_ORGBI;
    Tm.bar = bars + w;
    Tm.pred = reinterpret_cast<int16_t *>(Tm.me_area + 16);
    Tm.aux = Tm.pred + NP;
    Tm.blk = Tm.aux + NP;
    Tm.TB = reinterpret_cast<int32_t *>(Tm.blk + NY);
    Tm.tm = tm; Tm.tmT = tmT;
    Tm.scratch = scratch;
    CuHdr &H = *Tm.H;
    const int sh = (sq.bd - 8) << 1;
    uint32_t  phase = bphases[w];   // this warp's window mbarrier lives outside the regions (they overlay other working sets between calls)
    {   // CU record and input coder state -> this warp's header
        const uint32_t *src = reinterpret_cast<const uint32_t *>(git);
        uint32_t       *dst = reinterpret_cast<uint32_t *>(&H.cu);
        for(int e = lane; e < (int)(sizeof(xb200_cu_item) / 4); e += T) dst[e] = src[e];
        const xb200_sbac &s = st_in[git->state_in];
        for(int k = lane; k < XB200_CM_COUNT; k += T) H.st[ST_IN][k] = s.m[k];
        if(lane == 0) H.rg[ST_IN] = s.range;
        if(w == 0)
            for(int k = lane; k < 16; k += T) { P.skip_cost[k] = CU_MAX_COST; P.skip_ssd[k] = 0; }
    }
    __syncwarp();
    const xb200_cu_item &cu = H.cu;
    const xb200_rates   *rt = &rates[cu.rate_idx];
    const PicDev        &o = pics[cu.cur_pic];
    Tm.org[0] = o.p[0] + (ptrdiff_t)cu.y * o.s[0] + cu.x;
    Tm.org[1] = o.p[1] + (ptrdiff_t)(cu.y >> 1) * o.s[1] + (cu.x >> 1);
    Tm.org[2] = o.p[2] + (ptrdiff_t)(cu.y >> 1) * o.s[2] + (cu.x >> 1);
    Tm.so[0] = o.s[0]; Tm.so[1] = o.s[1]; Tm.so[2] = o.s[2];
    const bool   B = cu.slice_type == 0;
    const double w0 = cu.dist_chroma_weight[0], w1 = cu.dist_chroma_weight[1];
    double       cost_best = CU_MAX_COST, cost_l0 = CU_MAX_COST, cost_l1 = CU_MAX_COST;
    int          best_idx = 3;
    CU_PAR_SYNC(1);   // exchange area initialised
    { const int tt = lane + 32 * w; CU_PROF(0); }

    if(lane == 0) CH_DBG(12 + w, 100);
    // ---- phase 1: xeve_analyze_skip, the candidate pairs round-robin over the warps -----------------------------------------------
    // every warp walks the reference's candidate loop (duplicates pruned) and evaluates the pairs whose running number is its own
    int     my_best = -1;       // pair of this warp whose coded state sits in ST_MODE (the earliest of its minima)
    double  my_cost = CU_MAX_COST;
    {
        int n_pair = 0;
        for(int idx0 = 0; idx0 < sq.merge_num; idx0++) {
            bool dup = false;
            for(int t = idx0 - 1; t >= 0; t--) dup |= cu.mvp[0][t][0] == cu.mvp[0][idx0][0] && cu.mvp[0][t][1] == cu.mvp[0][idx0][1];
            if(dup) continue;
            const int cnt = B ? sq.merge_num : 1;
            for(int idx1 = 0; idx1 < cnt; idx1++) {
                dup = false;
                for(int t = idx1 - 1; t >= 0; t--) dup |= cu.mvp[1][t][0] == cu.mvp[1][idx1][0] && cu.mvp[1][t][1] == cu.mvp[1][idx1][1];
                if(dup) continue;
                const int8_t  refi[2] = {cu.refi_pred[0][idx0], (int8_t)(B ? cu.refi_pred[1][idx1] : -1)};
                const int16_t mv[2][2] = {{cu.mvp[0][idx0][0], cu.mvp[0][idx0][1]}, {cu.mvp[1][idx1][0], cu.mvp[1][idx1][1]}};
                if(refi[0] < 0 && refi[1] < 0) continue;
                if((n_pair++ % CU_PAR_WARPS) != w) continue;
                cu_predict<L2>(pics, cu, sq, refi[0], refi[1], mv[0][0], mv[0][1], mv[1][0], mv[1][1], Tm.pred, Tm.aux,
                               reinterpret_cast<int16_t *>(Tm.TB), lane);
                const int64_t cy = ssd_plane_t<L2, T>(Tm.org[0], Tm.so[0], Tm.pred, sh, lane, H.X);
                const int64_t cb = ssd_plane_t<L2 - 1, T>(Tm.org[1], Tm.so[1], Tm.pred + NY, sh, lane, H.X);
                const int64_t cr = ssd_plane_t<L2 - 1, T>(Tm.org[2], Tm.so[2], Tm.pred + NY + NCH, sh, lane, H.X);
                xb200_bits_item bi = cu_bits_item(cu, 0, 3, 0);
                bi.mvp_idx[0] = (uint8_t)idx0; bi.mvp_idx[1] = (uint8_t)idx1;
                const uint32_t bits = cu_count<T>(H, bi, nullptr, ST_IN, lane);
                double cost = __dadd_rn(__dadd_rn(__ll2double_rn(cy), __dmul_rn(w0, __ll2double_rn(cb))), __dmul_rn(w1, __ll2double_rn(cr)));
                cost = __dadd_rn(cost, __dmul_rn((double)bits, cu.lambda[0]));
                if(lane == 0) { P.skip_cost[idx0 * 4 + idx1] = cost; P.skip_ssd[idx0 * 4 + idx1] = cy + cb + cr; }
                if(cost < my_cost) {     // this warp's pairs come in reference order: strict less keeps its earliest minimum
                    my_cost = cost; my_best = idx0 * 4 + idx1;
                    cu_st_save<T>(H, ST_MODE, ST_RUN, lane);
                }
            }
        }
    }
    CU_PAR_SYNC(2);
    int64_t best_ssd = (int64_t)1 << (2 * L2 + 16);
    int     skip_pair = -1;
    {   // ordered arg-min over all pairs (reference order = increasing idx0, idx1; strict less)
        double sb = CU_MAX_COST;
        for(int k = 0; k < 16; k++) {
            const double c = P.skip_cost[k];
            if(c < sb) { sb = c; skip_pair = k; best_ssd = P.skip_ssd[k]; }
        }
        if(sb < cost_best) { cost_best = sb; best_idx = 3; }
    }
    const int skip_warp = skip_pair < 0 ? -1 : [&]() {     // which warp evaluated the winning pair: recount the running numbers
        int n_pair = 0, owner = -1;
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
                if(idx0 * 4 + idx1 == skip_pair) owner = n_pair % CU_PAR_WARPS;
                n_pair++;
            }
        }
        return owner;
    }();
    // md[3] and the coder state after the winning skip candidate -> every header's md[3]; warp 0's ST_BEST
    if(skip_pair >= 0) {
        const int idx0 = skip_pair >> 2, idx1 = skip_pair & 3;
        if(lane == 0) {
            CuMode &M = H.md[3];
            M.mvp_idx[0] = (uint8_t)idx0; M.mvp_idx[1] = (uint8_t)idx1;
            M.refi[0] = cu.refi_pred[0][idx0]; M.refi[1] = (int8_t)(B ? cu.refi_pred[1][idx1] : -1);
            M.mv[0][0] = cu.mvp[0][idx0][0]; M.mv[0][1] = cu.mvp[0][idx0][1]; M.mv[1][0] = cu.mvp[1][idx1][0]; M.mv[1][1] = cu.mvp[1][idx1][1];
            M.mvd[0][0] = M.mvd[0][1] = M.mvd[1][0] = M.mvd[1][1] = 0;
            M.nnz[0] = M.nnz[1] = M.nnz[2] = 0; M.cbf = 0;
        }
        // the owner's ST_MODE holds the state of ITS earliest minimum, which is the global winner when the winner is one of its pairs
        if(w == 0) {
            const CuHdr &Ho = *Hs[skip_warp];
            for(int k = lane; k < XB200_CM_COUNT; k += 32) H.st[ST_BEST][k] = Ho.st[ST_MODE][k];
            if(lane == 0) H.rg[ST_BEST] = Ho.rg[ST_MODE];
        }
        __syncwarp();
    }
    CU_PAR_SYNC(3);   // ST_MODE of the owner has been read; the warps may overwrite their slots
    { const int tt = lane + 32 * w; CU_PROF(1); }
    double cost_win = cost_best;
    const bool proceed = cost_best < CU_MAX_COST && best_ssd > 0;
    int        num_refp_cur = 0;
    if(proceed) {
        // ---- phase 2: direct | list 0 | list 1, one warp each ---------------------------------------------------------------------
        if(w == 0) {
            if(B) {
                if(lane == 0) {
                    CuMode &M = H.md[4];
                    M.refi[0] = M.refi[1] = 0; M.mvp_idx[0] = M.mvp_idx[1] = 0;
                    M.mv[0][0] = cu.mv_dir[0][0]; M.mv[0][1] = cu.mv_dir[0][1]; M.mv[1][0] = cu.mv_dir[1][0]; M.mv[1][1] = cu.mv_dir[1][1];
                    M.mvd[0][0] = M.mvd[0][1] = M.mvd[1][0] = M.mvd[1][1] = 0;
                }
                __syncwarp();
                if(lane == 0) CH_DBG(12 + w, 200);
                const double c = cu_residue_rdo<L2>(Tm, pics, rt, sq, 4, 0, 0, lane);
                if(lane == 0) P.cost[0] = c;
            }
        }
        else if(w == 1 || B) {
            const int lidx = w - 1;
            uint32_t  best_me = 0xffffffffu;
            int       refi_t = 0;
            int32_t   mot_bits[2] = {0, 0};
            const int nref = min((int)cu.num_refp[lidx], XB200_MAX_REFP);
            uint8_t   mvp_i = H.md[3].mvp_idx[lidx];
            const int16_t(*cand)[2] = cu.mvp[lidx];
            int16_t   mvs[XB200_MAX_REFP][2];
            for(int r = 0; r < nref; r++) {
                int            mx, my;
                if(lane == 0) CH_DBG(12 + w, 300 + r);
                const uint32_t mecost = cu_me<L2>(Tm, pics, sq, win_cap, err_flag, lidx, r, nref, 0, cand[mvp_i][0], cand[mvp_i][1], 0, 0, mot_bits,
                                                  phase, mx, my, lane);
                mvs[r][0] = (int16_t)mx; mvs[r][1] = (int16_t)my;
                if(mecost < best_me) { best_me = mecost; refi_t = r; }
            }
            const int mvx = mvs[refi_t][0], mvy = mvs[refi_t][1];
            if(lane == 0) CH_DBG(12 + w, 400);
            {   // check_best_mvp: the loop compares against the cost of the initial index only (quirk q1)
                xb200_bits_item bi = cu_bits_item(cu, 2, lidx, 0);
                bi.refi[lidx] = (int8_t)refi_t;
                bi.mvp_idx[0] = mvp_i;
                bi.mvd[lidx][0] = (int16_t)(mvx - cand[mvp_i][0]); bi.mvd[lidx][1] = (int16_t)(mvy - cand[mvp_i][1]);
                const double ref_cost = __dmul_rn((double)cu_count<T>(H, bi, nullptr, ST_IN, lane), cu.lambda[0]);
                int          best = mvp_i;
                for(int idx = 0; idx < 4; idx++) {
                    bool dup = false;
                    for(int t = idx - 1; t >= 0; t--) dup |= cand[idx][0] == cand[t][0] && cand[idx][1] == cand[t][1];
                    if(dup) continue;
                    bi.mvp_idx[0] = (uint8_t)idx;
                    bi.mvd[lidx][0] = (int16_t)(mvx - cand[idx][0]); bi.mvd[lidx][1] = (int16_t)(mvy - cand[idx][1]);
                    const double c = __dmul_rn((double)cu_count<T>(H, bi, nullptr, ST_IN, lane), cu.lambda[0]);
                    if(c < ref_cost) best = idx;
                }
                mvp_i = (uint8_t)best;
            }
            if(lane == 0) {
                CuMode &M = H.md[lidx];
                M.refi[lidx] = (int8_t)refi_t; M.refi[1 - lidx] = -1;
                M.mv[lidx][0] = (int16_t)mvx; M.mv[lidx][1] = (int16_t)mvy; M.mv[1 - lidx][0] = M.mv[1 - lidx][1] = 0;
                M.mvd[lidx][0] = (int16_t)(mvx - cand[mvp_i][0]); M.mvd[lidx][1] = (int16_t)(mvy - cand[mvp_i][1]);
                M.mvd[1 - lidx][0] = M.mvd[1 - lidx][1] = 0;
                M.mvp_idx[lidx] = mvp_i; M.mvp_idx[1 - lidx] = 0;
                P.mot_bits[lidx] = mot_bits[lidx];
                for(int r = 0; r < nref; r++) { P.mv_scale[lidx][r][0] = mvs[r][0]; P.mv_scale[lidx][r][1] = mvs[r][1]; }
            }
            __syncwarp();
            if(lane == 0) CH_DBG(12 + w, 500);
            // the bit counter reads mvp_idx[] of the coded list only (refi of the other list is -1)
            const double c = cu_residue_rdo<L2>(Tm, pics, rt, sq, lidx, lidx == 0 ? mvp_i : 0, lidx == 1 ? mvp_i : 0, lane);
            if(lane == 0) P.cost[1 + lidx] = c;
        }
        CU_PAR_SYNC(4);
        { const int tt = lane + 32 * w; CU_PROF(3); }
        // ---- merge in the reference's order: direct, list 0, list 1 (strict less) ----
        int state_from = -1;   // warp whose ST_MODE becomes s_next_best so far (-1: the skip state already in warp 0's ST_BEST)
        if(B) {
            const double c = P.cost[0];
            if(c < cost_best) { cost_best = c; best_idx = 4; state_from = 0; }
        }
        for(int lidx = 0; lidx <= (B ? 1 : 0); lidx++) {
            const double c = P.cost[1 + lidx];
            if(lidx == 0) cost_l0 = c; else cost_l1 = c;
            if(c < cost_best) { cost_best = c; best_idx = lidx; state_from = 1 + lidx; }
            num_refp_cur = min((int)cu.num_refp[lidx], XB200_MAX_REFP);
        }
        if(w == 0) {
            if(state_from >= 0) {
                const CuHdr &Hf = *Hs[state_from];
                for(int k = lane; k < XB200_CM_COUNT; k += 32) H.st[ST_BEST][k] = Hf.st[ST_MODE][k];
                if(lane == 0) H.rg[ST_BEST] = Hf.rg[ST_MODE];
            }
            if(lane < 2 && (lane == 0 || B)) H.md[lane] = Hs[1 + lane]->md[lane];   // the uni modes -> warp 0's header (bi loop, copy-out)
            __syncwarp();
        }
        CU_PAR_SYNC(5);   // the other warps' headers have been read
        // ---- phase 3 (warp 0): analyze_bi, start MVs = the uni results, at most BI_ITER = 4 passes ----
        if(w == 0 && B) {
            int32_t  mot_bits[2] = {P.mot_bits[0], P.mot_bits[1]};
            int16_t  mv_scale[2][XB200_MAX_REFP][2];
            for(int l = 0; l < 2; l++)
                for(int r = 0; r < XB200_MAX_REFP; r++) { mv_scale[l][r][0] = P.mv_scale[l][r][0]; mv_scale[l][r][1] = P.mv_scale[l][r][1]; }
            int      lidx_ref = cost_l0 <= cost_l1 ? 0 : 1, lidx_cnd = 1 - lidx_ref;
            int8_t   refi[2] = {-1, -1};
            int8_t   m_refi[2] = {H.md[0].refi[0], H.md[1].refi[1]};
            int16_t  m_mv[2][2] = {{H.md[0].mv[0][0], H.md[0].mv[0][1]}, {H.md[1].mv[1][0], H.md[1].mv[1][1]}};
            const uint8_t m_idx[2] = {H.md[0].mvp_idx[0], H.md[1].mvp_idx[1]};
            uint32_t best_me = 0xffffffffu;
            int      refi_best = 0;
            refi[lidx_ref] = m_refi[lidx_ref];
            for(int iter = 0; iter < 4; iter++) {
                cu_predict<L2>(pics, cu, sq, refi[0], refi[1], m_mv[0][0], m_mv[0][1], m_mv[1][0], m_mv[1][1], Tm.pred, Tm.aux,
                               reinterpret_cast<int16_t *>(Tm.TB), lane);
                for(int e = lane; e < NY; e += T)   // get_org_bi
                    Tm.org_bi[e] = (int16_t)(((int)Tm.org[0][(ptrdiff_t)(e >> L2) * Tm.so[0] + (e & (N - 1))] << 1) - (int)Tm.pred[e]);
                __syncwarp();
                { const int8_t t = refi[lidx_ref]; refi[lidx_ref] = refi[lidx_cnd]; refi[lidx_cnd] = t; }
                { const int t = lidx_ref; lidx_ref = lidx_cnd; lidx_cnd = t; }
                const int mi = m_idx[lidx_ref];
                bool      changed = false;
                for(int r = 0; r < num_refp_cur; r++) {
                    int            mx, my;
                    const uint32_t mecost = cu_me<L2>(Tm, pics, sq, win_cap, err_flag, lidx_ref, r, num_refp_cur, 1, cu.mvp[lidx_ref][mi][0],
                                                      cu.mvp[lidx_ref][mi][1], mv_scale[lidx_ref][r][0], mv_scale[lidx_ref][r][1], mot_bits,
                                                      phase, mx, my, lane);
                    mv_scale[lidx_ref][r][0] = (int16_t)mx; mv_scale[lidx_ref][r][1] = (int16_t)my;
                    if(mecost < best_me) {
                        refi_best = r; best_me = mecost; changed = true;
                        m_refi[lidx_ref] = (int8_t)r;
                        m_mv[lidx_ref][0] = (int16_t)mx; m_mv[lidx_ref][1] = (int16_t)my;
                    }
                }
                refi[lidx_ref] = (int8_t)refi_best; refi[lidx_cnd] = -1;
                if(!changed) break;
            }
            { const int tt = lane; CU_PROF(6); }
            if(lane == 0) {
                CuMode &M = H.md[2];
                for(int l = 0; l < 2; l++) {
                    M.refi[l] = m_refi[l]; M.mvp_idx[l] = m_idx[l];
                    M.mv[l][0] = m_mv[l][0]; M.mv[l][1] = m_mv[l][1];
                    M.mvd[l][0] = (int16_t)(m_mv[l][0] - cu.mvp[l][m_idx[l]][0]); M.mvd[l][1] = (int16_t)(m_mv[l][1] - cu.mvp[l][m_idx[l]][1]);
                }
            }
            __syncwarp();
            const double c = cu_residue_rdo<L2>(Tm, pics, rt, sq, 2, m_idx[0], m_idx[1], lane);
            if(c < cost_best) { cost_best = c; best_idx = 2; cu_st_save<T>(H, ST_BEST, ST_MODE, lane); }
        }
        cost_win = cost_best;   // (only warp 0's value is used below)
    }
    // ---- winner (warp 0): coefficients (dropped planes zeroed), reconstruction, XEVE_MODE fields, s_next_best ----------------------
    if(w == 0) {
        const CuMode &M = H.md[best_idx];
        int16_t      *gc = coef_out + cu.out_off, *gr = rec_out ? rec_out + cu.out_off : nullptr;
        if(best_idx == 3) {
            cu_predict<L2>(pics, cu, sq, M.refi[0], M.refi[1], M.mv[0][0], M.mv[0][1], M.mv[1][0], M.mv[1][1], Tm.pred, Tm.aux,
                           reinterpret_cast<int16_t *>(Tm.TB), lane);
            for(int e = lane; e < NP; e += T) {
                gc[e] = 0;
                if(gr) gr[e] = Tm.pred[e];
                if(pred_y_out && e < NY) pred_y_out[e] = Tm.pred[e];
            }
        }
        else {
            const int16_t *sc = Tm.scratch + (size_t)(3 * best_idx) * NP, *sr = sc + NP, *sp = sr + NP;
            const int      cbf = M.cbf;
            for(int e = lane; e < NP; e += T) {
                const int  c = e < NY ? 0 : (e < NY + NCH ? 1 : 2);
                const bool on = (cbf >> c) & 1;
                gc[e] = on ? __ldcg(sc + e) : (int16_t)0;
                if(gr) gr[e] = on ? __ldcg(sr + e) : __ldcg(sp + e);
                if(pred_y_out && e < NY) pred_y_out[e] = __ldcg(sp + e);
            }
        }
        if(lane == 0) {
            git->cost = cost_win; git->best_idx = (uint8_t)best_idx;
            for(int l = 0; l < 2; l++) {
                git->refi[l] = M.refi[l]; git->mvp_idx[l] = M.mvp_idx[l];
                git->mv[l][0] = M.mv[l][0]; git->mv[l][1] = M.mv[l][1]; git->mvd[l][0] = M.mvd[l][0]; git->mvd[l][1] = M.mvd[l][1];
            }
            git->nnz[0] = M.nnz[0]; git->nnz[1] = M.nnz[1]; git->nnz[2] = M.nnz[2];
        }
        if(cu.state_out >= 0) {
            xb200_sbac &so = st_out[cu.state_out];
            for(int k = lane; k < XB200_CM_COUNT; k += 32) so.m[k] = H.st[ST_BEST][k];
            if(lane == 0) so.range = H.rg[ST_BEST];
        }
        __syncwarp();
        { const int tt = lane; CU_PROF(8); }
    }
    if(lane == 0) bphases[w] = phase;
    CU_PAR_SYNC(6);
}
