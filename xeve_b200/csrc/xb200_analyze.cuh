// xb200_analyze.cuh -- the whole inter mode decision of one CU on the device.
//
//   reference: xeve_pinter_analyze_cu (src_base/xeve_pinter.c:1839-2056) = xeve_analyze_skip (:1337-1530) + analyze_t_direct
//   (:1532-1565) + uni-directional search with check_best_mvp (:1772-1837) + analyze_bi (:1567-1683), every residual mode
//   through pinter_residue_rdo and its cbf decisions (:906-1335), rate terms from the CABAC bit counter (xb200_rate.cuh).
//
// One TEAM per CU (a warp for 8x8 / 16x16, 128 threads for 32x32, 256 for 64x64; persistent CTAs).  The team walks the
// reference's decision sequence; inside every step the data-parallel work is spread over the team with the SAME device
// functions the work-list operators use (me_search of xb200_me.cuh, mc_item_t / residue_plane of xb200_residue2.cuh),
// and the serial coder runs on warp 0.  Decision values (distortions, bit counts, double costs) are uniform across the
// team: every thread evaluates the cost expressions redundantly, in the reference's association order, with explicit
// round-to-nearest double intrinsics so that no fused multiply-add is formed.
// Candidate modes keep {coef, rec, pred} in a per-team global scratch slot (L2 resident); the winner is copied out.
#pragma once
#include "xb200_common.cuh"
// Phase timers of the decision chain (debug builds with -DXB200_CHAIN_PROF): cycles per phase, accumulated by thread 0 of the team.
#ifdef XB200_CHAIN_PROF
__device__ unsigned long long g_chain_prof[64];
__device__ long long g_prof_last;
#define CU_PROF(k) do { if(tt == 0) { const long long now_ = clock64(); atomicAdd(&g_chain_prof[(k)], (unsigned long long)(now_ - g_prof_last)); atomicAdd(&g_chain_prof[32 + (k)], 1ull); g_prof_last = now_; } } while(0)
#else
#define CU_PROF(k) do { } while(0)
#endif

#include "xb200_me.cuh"
#include "xb200_rate.cuh"
#include "xb200_residue2.cuh"

struct CuMode { // one candidate mode of the CU (pi->refi/mv/mvd/mvp_idx[pidx], nnz_best[pidx])
    int8_t  refi[2];
    uint8_t mvp_idx[2];
    int16_t mv[2][2], mvd[2][2];
    int32_t nnz[3];   // after the cbf decision
    int32_t cbf;      // bit c: plane c keeps its coefficients
};
enum { ST_IN = 0, ST_RUN = 1, ST_MODE = 2, ST_CPREV = 3, ST_CRUN = 4, ST_BEST = 5, CU_NST = 6 };
struct CuHdr {
    xb200_cu_item cu;
    xb200_me_item me;
    uint32_t      rg[CU_NST];                    // coder ranges of the state slots
    uint32_t      bits;                          // broadcast of the last bit count
    uint16_t      st[CU_NST][XB200_CM_COUNT + 4];// coder models: input, working, s_temp_best, comp chain (2), s_next_best
    CuMode        md[5];
    TeamScratch   X;
};
#define CU_MAX_COST (1.7e+308)

template <int L2, int TEAMS_ = Res2Cfg<L2>::TEAMS> struct CuCfg { // TEAMS_: CUs in flight per CTA (fewer when the search window is large)
    using R = Res2Cfg<L2>;
    static constexpr int T = R::T, TEAMS = TEAMS_, CTA = T * TEAMS, N = R::N, NY = N * N, NCH = NY >> 2, NP = R::PRED;
    static constexpr int HDR = ((int)sizeof(CuHdr) + 15) & ~15;
    static constexpr int ORGBI = NY * 2;                       // 2*org - pred block of the bi search (shared)
    static constexpr int SCRATCH = 5 * 3 * NP;                 // s16 elements of global scratch per team
    __host__ __device__ static size_t team_bytes(int win_cap)
    {
        const size_t me = me_team_bytes(L2, win_cap), rs = 16 + (size_t)R::TEAM_BYTES;
        return HDR + ORGBI + (((me > rs ? me : rs) + 15) & ~(size_t)15);
    }
    __host__ __device__ static size_t smem_bytes(int win_cap) { return 8192 + TEAMS * team_bytes(win_cap); }
};

// ---- coder state plumbing (warp 0 of the team) ------------------------------------------------------------------------
XB_DEV void cu_st_copy(CuHdr &H, int dst, int src, int lane)
{
    for(int k = lane; k < XB200_CM_COUNT; k += 32) H.st[dst][k] = H.st[src][k];
    if(lane == 0) H.rg[dst] = H.rg[src];
    __syncwarp();
}
template <int T> __device__ __noinline__ void cu_st_save(CuHdr &H, int dst, int src, int tt)
{
    if(tt < 32) cu_st_copy(H, dst, src, tt);
    team_sync<T>();
}
// SBAC_LOAD(s_temp_run, slot src) -> xeve_sbac_bit_reset -> syntax of `it` -> xeve_get_bit_number; the coded state stays in ST_RUN
template <int T> __device__ __noinline__ uint32_t cu_count(CuHdr &H, const xb200_bits_item &it, const int16_t *coef, int src, int tt)
{
    if(tt < 32) {
        cu_st_copy(H, ST_RUN, src, tt);
        Cabac c;
        c.range = H.rg[ST_RUN]; c.bits = 0; c.m = H.st[ST_RUN];
        cb_count_item(c, it, coef, tt);
        if(tt == 0) { H.bits = c.bits; H.rg[ST_RUN] = c.range; }
        __syncwarp(); // lane 0 may trail the others through its last bins: reconverge before a block-wide barrier
    }
    team_sync<T>();
    const uint32_t b = H.bits;
    team_sync<T>();
    return b;
}
XB_DEV xb200_bits_item cu_bits_item(const xb200_cu_item &cu, int kind, int pidx, int ch)
{
    xb200_bits_item it;
    it.kind = (uint8_t)kind; it.slice_type = cu.slice_type; it.log2_cuw = cu.log2_cuw; it.log2_cuh = cu.log2_cuh;
    it.pidx = (uint8_t)pidx; it.ch = (uint8_t)ch; it.ctx_skip = cu.ctx_skip; it.ctx_pred_mode = cu.ctx_pred_mode;
    it.refi[0] = it.refi[1] = -1; it.mvp_idx[0] = it.mvp_idx[1] = 0;
    it.num_refp[0] = cu.num_refp[0]; it.num_refp[1] = cu.num_refp[1];
    it.all_preds = cu.all_preds; it.pad_ = 0;
    it.mvd[0][0] = it.mvd[0][1] = it.mvd[1][0] = it.mvd[1][1] = 0;
    it.nnz[0] = it.nnz[1] = it.nnz[2] = 0;
    it.state_in = 0; it.state_out = -1; it.coef_off = 0; it.bits = 0; it.pad2_ = 0;
    return it;
}
XB_DEV void cu_mc_item(const xb200_cu_item &cu, int n, const int8_t refi[2], const int16_t mv[2][2], xb200_mc_item &m)
{
    m.poc = cu.poc; m.x = cu.x; m.y = cu.y; m.w = m.h = (int16_t)n; m.out_hash = 0;
#pragma unroll
    for(int l = 0; l < 2; l++) {
        m.refi[l] = refi[l]; m.mv[l][0] = mv[l][0]; m.mv[l][1] = mv[l][1];
        m.ref_pic[l] = refi[l] >= 0 ? cu.ref_pic[l][refi[l] & 3] : -1;
        m.ref_poc[l] = refi[l] >= 0 ? cu.ref_poc[l][refi[l] & 3] : -1;
    }
}
template <int LN, int T> XB_DEV int64_t ssd_plane_t(const int16_t *__restrict__ org, int so, const int16_t *pr, int sh, int tt, TeamScratch &X)
{
    constexpr int N = 1 << LN;
    int64_t       part = 0;
    for(int e = tt; e < N * N; e += T) {
        const int d = (int)org[(ptrdiff_t)(e >> LN) * so + (e & (N - 1))] - (int)pr[e];
        part += (d * d) >> sh;
    }
    return team_sum_s64<T>(part, tt, X);
}

// prediction of the CU for (refi, mv) into the team's pred buffer: one out-of-line copy of the interpolation code per size
template <int L2>
__device__ __noinline__ void cu_predict(const PicDev *__restrict__ pics, const xb200_cu_item &cu, const SeqDev &sq, int refi0, int refi1, int mv00,
                                        int mv01, int mv10, int mv11, int16_t *pred, int16_t *aux, int16_t *tmp, int tt)
{
    const int8_t  refi[2] = {(int8_t)refi0, (int8_t)refi1};
    const int16_t mv[2][2] = {{(int16_t)mv00, (int16_t)mv01}, {(int16_t)mv10, (int16_t)mv11}};
    xb200_mc_item mc;
    cu_mc_item(cu, 1 << L2, refi, mv, mc);
    mc_item_t<L2, Res2Cfg<L2>::T>(pics, mc, sq, pred, aux, tmp, tt);
}

// per-team working set handed to the helpers
template <int L2> struct CuTeam {
    CuHdr         *H;
    unsigned char *me_area;        // me_team_bytes(): 16 reserved bytes, then the search working set
    uint64_t      *bar;            // mbarrier of the search-window copies (initialised once by the kernel)
    int16_t       *pred, *aux, *blk, *org_bi;
    int32_t       *TB;
    const int8_t  *tm, *tmT;
    int16_t       *scratch;        // global: mode m -> coef at (3m)*NP, rec at (3m+1)*NP, pred at (3m+2)*NP
    const int16_t *org[3];
    int            so[3];
};

// The cbf decisions of pinter_residue_rdo (src_base/xeve_pinter.c:1087-1335) for one candidate mode whose transform results are
// known: store[] = nnz of the coded planes, d0 / d1 = SSD of prediction / reconstruction, gco = the coefficient planes.
// Returns the RD cost, the chosen plane mask in cbf, and leaves s_temp_best in ST_MODE.
template <int T>
__device__ __noinline__ double cu_cbf_decide(CuHdr &H, int pidx, const CuMode &M, uint8_t mi0, uint8_t mi1, const int *store, const int64_t *d0,
                                             const int64_t *d1, const int16_t *gco, int &cbf_out, int tt)
{
    const xb200_cu_item &cu = H.cu;
    const double   w0 = cu.dist_chroma_weight[0], w1 = cu.dist_chroma_weight[1];
    xb200_bits_item bi = cu_bits_item(cu, 1, pidx, 0);
    bi.refi[0] = M.refi[0]; bi.refi[1] = M.refi[1]; bi.mvp_idx[0] = mi0; bi.mvp_idx[1] = mi1;
    bi.mvd[0][0] = M.mvd[0][0]; bi.mvd[0][1] = M.mvd[0][1]; bi.mvd[1][0] = M.mvd[1][0]; bi.mvd[1][1] = M.mvd[1][1];
    double best = CU_MAX_COST;
    int    cbf = 0;
    if(store[0] + store[1] + store[2]) {
        auto try_comb = [&](int n0, int n1, int n2) {
            bi.kind = 1;
            bi.nnz[0] = n0 ? store[0] : 0; bi.nnz[1] = n1 ? store[1] : 0; bi.nnz[2] = n2 ? store[2] : 0;
            const uint32_t bits = cu_count<T>(H, bi, gco, ST_IN, tt);
            double cost = __dadd_rn(__ll2double_rn(n0 ? d1[0] : d0[0]),
                                    __dadd_rn(__dmul_rn(__ll2double_rn(n1 ? d1[1] : d0[1]), w0), __dmul_rn(__ll2double_rn(n2 ? d1[2] : d0[2]), w1)));
            cost = __dadd_rn(cost, __dmul_rn((double)bits, cu.lambda[0]));
            if(cost < best) {
                best = cost; cbf = (n0 ? 1 : 0) | (n1 ? 2 : 0) | (n2 ? 4 : 0);
                cu_st_save<T>(H, ST_MODE, ST_RUN, tt);
            }
        };
        if(pidx != 4) try_comb(0, 0, 0);                               // forced all-zero
        try_comb(store[0] > 0, store[1] > 0, store[2] > 0);            // as it is
        int idx_best[3] = {0, 0, 0};
        cu_st_save<T>(H, ST_CPREV, ST_IN, tt);
        bi.kind = 3;
        bi.nnz[0] = store[0]; bi.nnz[1] = store[1]; bi.nnz[2] = store[2];
#pragma unroll
        for(int i = 0; i < 3; i++) {                                   // per-component cbf test, coder state chained
            if(store[i] <= 0) continue;
            double comp_best = CU_MAX_COST;
            cu_st_save<T>(H, ST_CRUN, ST_CPREV, tt);
            for(int j = 0; j < 2; j++) {
                bi.ch = (uint8_t)i;
                bi.nnz[i] = j ? store[i] : 0;
                const uint32_t bits = cu_count<T>(H, bi, gco, ST_CRUN, tt);
                double cost = i == 0 ? __ll2double_rn(j ? d1[0] : d0[0]) : __dmul_rn(__ll2double_rn(j ? d1[i] : d0[i]), i == 1 ? w0 : w1);
                cost = __dadd_rn(cost, __dmul_rn((double)bits, cu.lambda[i]));
                if(cost < comp_best) { comp_best = cost; idx_best[i] = j; cu_st_save<T>(H, ST_CPREV, ST_RUN, tt); }
            }
        }
        if(idx_best[0] || idx_best[1] || idx_best[2]) {
            const bool differs = (idx_best[0] ? store[0] : 0) != store[0] || (idx_best[1] ? store[1] : 0) != store[1] ||
                                 (idx_best[2] ? store[2] : 0) != store[2];
            if(differs) try_comb(idx_best[0], idx_best[1], idx_best[2]);
        }
    }
    else {
        best = __dadd_rn(__dadd_rn(__ll2double_rn(d0[0]), __dmul_rn(w0, __ll2double_rn(d0[1]))), __dmul_rn(w1, __ll2double_rn(d0[2])));
        bi.kind = 1;
        const uint32_t bits = cu_count<T>(H, bi, gco, ST_IN, tt);
        best = __dadd_rn(best, __dmul_rn((double)bits, cu.lambda[0]));
        cu_st_save<T>(H, ST_MODE, ST_RUN, tt);
    }
    cbf_out = cbf;
    return best;
}

// pinter_residue_rdo for mode pidx (H.md[pidx].refi/mv/mvd set and visible): fills md[pidx].nnz/cbf, leaves s_temp_best in ST_MODE
template <int L2>
__device__ __noinline__ double cu_residue_rdo(const CuTeam<L2> &Tm, const PicDev *__restrict__ pics, const xb200_rates *__restrict__ rt,
                                              const SeqDev &sq, int pidx, uint8_t mi0, uint8_t mi1, int tt)
{
    using Cf = CuCfg<L2>;
    constexpr int T = Cf::T, N = Cf::N, NY = Cf::NY, NCH = Cf::NCH, NP = Cf::NP;
    constexpr int LNMAX = L2 >= 5 ? L2 : 5;
    CuHdr               &H = *Tm.H;
    const xb200_cu_item &cu = H.cu;
    CuMode              &M = H.md[pidx];
    int16_t *gco = Tm.scratch + (size_t)(3 * pidx) * NP, *grec = gco + NP, *gpred = grec + NP;
    cu_predict<L2>(pics, cu, sq, M.refi[0], M.refi[1], M.mv[0][0], M.mv[0][1], M.mv[1][0], M.mv[1][1], Tm.pred, Tm.aux,
                   reinterpret_cast<int16_t *>(Tm.TB), tt);
    for(int e = tt; e < NP; e += T) gpred[e] = Tm.pred[e];
    CU_PROF(12);
    int     store[3];
    int64_t d0[3], d1[3];
    residue_plane<L2, T, LNMAX, false>(Tm.org[0], Tm.so[0], Tm.pred, Tm.blk, Tm.TB, Tm.tm, Tm.tmT, gco, grec, 1, cu.qp[0], cu.lambda[0], 0,
                                       cu.slice_type, rt, sq, tt, H.X, store[0], d0[0], d1[0], nullptr, nullptr);
    residue_plane<L2 - 1, T, LNMAX, false>(Tm.org[1], Tm.so[1], Tm.pred + NY, Tm.blk, Tm.TB, Tm.tm, Tm.tmT, gco + NY, grec + NY, 1, cu.qp[1],
                                           cu.lambda[1], 1, cu.slice_type, rt, sq, tt, H.X, store[1], d0[1], d1[1], nullptr, nullptr);
    residue_plane<L2 - 1, T, LNMAX, false>(Tm.org[2], Tm.so[2], Tm.pred + NY + NCH, Tm.blk, Tm.TB, Tm.tm, Tm.tmT, gco + NY + NCH, grec + NY + NCH,
                                           1, cu.qp[2], cu.lambda[2], 2, cu.slice_type, rt, sq, tt, H.X, store[2], d0[2], d1[2], nullptr, nullptr);
    team_sync<T>(); // coefficient planes visible to the coder warp
    CU_PROF(13);
    int          cbf;
    const double best = cu_cbf_decide<T>(H, pidx, M, mi0, mi1, store, d0, d1, gco, cbf, tt);
    CU_PROF(14);
    if(tt == 0) {
        M.cbf = cbf;
        M.nnz[0] = (cbf & 1) ? store[0] : 0; M.nnz[1] = (cbf & 2) ? store[1] : 0; M.nnz[2] = (cbf & 4) ? store[2] : 0;
    }
    team_sync<T>();
    return best;
}

// one pi->fn_me call built from the CU record
template <int L2>
__device__ __noinline__ uint32_t cu_me(const CuTeam<L2> &Tm, const PicDev *__restrict__ pics, const SeqDev &sq, int win_cap, int *err_flag, int lidx,
                                       int refi, int num_refp, int bi, int mvp_x, int mvp_y, int mv_in_x, int mv_in_y, int32_t *mot_bits,
                                       uint32_t &phase, int &mv_x, int &mv_y, int tt)
{
    constexpr int        T = CuCfg<L2>::T;
    CuHdr               &H = *Tm.H;
    const xb200_cu_item &cu = H.cu;
    if(tt == 0) {
        xb200_me_item &me = H.me;
        me.poc = cu.poc; me.cur_pic = cu.cur_pic; me.ref_pic = cu.ref_pic[lidx][refi]; me.ref_poc = cu.ref_poc[lidx][refi];
        me.x = cu.x; me.y = cu.y; me.log2_cuw = cu.log2_cuw; me.log2_cuh = cu.log2_cuh; me.lidx = (uint8_t)lidx; me.bi = (uint8_t)bi;
        me.refi = (int8_t)refi; me.num_refp = (uint8_t)num_refp;
        me.mvp[0] = (int16_t)mvp_x; me.mvp[1] = (int16_t)mvp_y; me.mv_in[0] = (int16_t)mv_in_x; me.mv_in[1] = (int16_t)mv_in_y;
        me.lambda_mv = cu.lambda_mv; me.mot_bits_in[0] = mot_bits[0]; me.mot_bits_in[1] = mot_bits[1];
        me.max_search_range = cu.max_search_range; me.gop_size = sq.gop_size; me.org_bi_off = bi ? 0 : -1;
    }
    team_sync<T>();
    uint32_t cost;
    int      mb;
    me_search<L2>(Tm.me_area, Tm.bar, pics, &H.me, Tm.org_bi, sq, win_cap, err_flag, tt, phase, mv_x, mv_y, cost, mb);
    mot_bits[lidx] = mb;
    return cost;
}

// xeve_pinter_analyze_cu of ONE CU by one team (the first T threads of the CTA or one warp): `git` is the CU record (inputs read,
// results written), st_in / st_out / rates / coef_out / rec_out are the arrays its indices and offsets refer to.  pred_y_out
// (optional) receives the winner's luma prediction (mi->pred_y_best, the input of core->inter_satd, src_base/xeve_mode.c:1247-1258).
// `phase` is the parity of the team's search-window mbarrier and persists across calls.
template <int L2>
__device__ __noinline__ void analyze_cu_one(CuTeam<L2> &Tm, const PicDev *__restrict__ pics, xb200_cu_item *git, const xb200_rates *rates,
                                            const xb200_sbac *st_in, xb200_sbac *st_out, int16_t *coef_out, int16_t *rec_out,
                                            int16_t *pred_y_out, const SeqDev &sq, int win_cap, int *err_flag, uint32_t &phase, int tt)
{
    using Cf = CuCfg<L2>;
    constexpr int T = Cf::T, N = Cf::N, NY = Cf::NY, NCH = Cf::NCH, NP = Cf::NP;
    CuHdr    &H = *Tm.H;
    const int sh = (sq.bd - 8) << 1;
    {   // CU record and input coder state -> shared
        const uint32_t *src = reinterpret_cast<const uint32_t *>(git);
        uint32_t       *dst = reinterpret_cast<uint32_t *>(&H.cu);
        for(int e = tt; e < (int)(sizeof(xb200_cu_item) / 4); e += T) dst[e] = src[e];
        const xb200_sbac &s = st_in[git->state_in];
        for(int k = tt; k < XB200_CM_COUNT; k += T) H.st[ST_IN][k] = s.m[k];
        if(tt == 0) H.rg[ST_IN] = s.range;
    }
    team_sync<T>();
    CU_PROF(0);
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

    // ---- xeve_analyze_skip: merge_num (x merge_num in B) candidate pairs, duplicates pruned ----------------
    int64_t best_ssd = (int64_t)1 << (2 * L2 + 16);
    {
        double sb = CU_MAX_COST;
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
                cu_predict<L2>(pics, cu, sq, refi[0], refi[1], mv[0][0], mv[0][1], mv[1][0], mv[1][1], Tm.pred, Tm.aux,
                               reinterpret_cast<int16_t *>(Tm.TB), tt);
                const int64_t cy = ssd_plane_t<L2, T>(Tm.org[0], Tm.so[0], Tm.pred, sh, tt, H.X);
                const int64_t cb = ssd_plane_t<L2 - 1, T>(Tm.org[1], Tm.so[1], Tm.pred + NY, sh, tt, H.X);
                const int64_t cr = ssd_plane_t<L2 - 1, T>(Tm.org[2], Tm.so[2], Tm.pred + NY + NCH, sh, tt, H.X);
                xb200_bits_item bi = cu_bits_item(cu, 0, 3, 0);
                bi.mvp_idx[0] = (uint8_t)idx0; bi.mvp_idx[1] = (uint8_t)idx1;
                const uint32_t bits = cu_count<T>(H, bi, nullptr, ST_IN, tt);
                double cost = __dadd_rn(__dadd_rn(__ll2double_rn(cy), __dmul_rn(w0, __ll2double_rn(cb))), __dmul_rn(w1, __ll2double_rn(cr)));
                cost = __dadd_rn(cost, __dmul_rn((double)bits, cu.lambda[0]));
                if(cost < sb) {
                    sb = cost;
                    best_ssd = cy + cb + cr;
                    if(tt == 0) {
                        CuMode &M = H.md[3];
                        M.mvp_idx[0] = (uint8_t)idx0; M.mvp_idx[1] = (uint8_t)idx1;
                        M.refi[0] = refi[0]; M.refi[1] = refi[1];
                        M.mv[0][0] = mv[0][0]; M.mv[0][1] = mv[0][1]; M.mv[1][0] = mv[1][0]; M.mv[1][1] = mv[1][1];
                        M.mvd[0][0] = M.mvd[0][1] = M.mvd[1][0] = M.mvd[1][1] = 0;
                        M.nnz[0] = M.nnz[1] = M.nnz[2] = 0; M.cbf = 0;
                    }
                    cu_st_save<T>(H, ST_MODE, ST_RUN, tt);
                }
            }
        }
        if(sb < cost_best) { cost_best = sb; best_idx = 3; cu_st_save<T>(H, ST_BEST, ST_MODE, tt); }
    }
    team_sync<T>();
    CU_PROF(1);
    double cost_win = cost_best;
    if(cost_best < CU_MAX_COST && best_ssd > 0) {
        if(B) { // ---- analyze_t_direct ----
            if(tt == 0) {
                CuMode &M = H.md[4];
                M.refi[0] = M.refi[1] = 0; M.mvp_idx[0] = M.mvp_idx[1] = 0;
                M.mv[0][0] = cu.mv_dir[0][0]; M.mv[0][1] = cu.mv_dir[0][1]; M.mv[1][0] = cu.mv_dir[1][0]; M.mv[1][1] = cu.mv_dir[1][1];
                M.mvd[0][0] = M.mvd[0][1] = M.mvd[1][0] = M.mvd[1][1] = 0;
            }
            team_sync<T>();
            const double c = cu_residue_rdo<L2>(Tm, pics, rt, sq, 4, 0, 0, tt);
            if(c < cost_best) { cost_best = c; best_idx = 4; cu_st_save<T>(H, ST_BEST, ST_MODE, tt); }
        }
        // ---- uni-directional search per list, best reference by ME cost, check_best_mvp, residue RDO ----
        int32_t mot_bits[2] = {0, 0};
        int16_t mv_scale[2][XB200_MAX_REFP][2];
        uint8_t mvp_idx[2] = {0, 0};
        int     num_refp_cur = 0;
        for(int lidx = 0; lidx <= (B ? 1 : 0); lidx++) {
            uint32_t best_me = 0xffffffffu;
            int      refi_t = 0;
            num_refp_cur = min((int)cu.num_refp[lidx], XB200_MAX_REFP);
            mvp_idx[lidx] = H.md[3].mvp_idx[lidx];
            const int16_t(*cand)[2] = cu.mvp[lidx];
            for(int r = 0; r < num_refp_cur; r++) {
                int            mx, my;
                const uint32_t mecost = cu_me<L2>(Tm, pics, sq, win_cap, err_flag, lidx, r, num_refp_cur, 0, cand[mvp_idx[lidx]][0],
                                                  cand[mvp_idx[lidx]][1], 0, 0, mot_bits, phase, mx, my, tt);
                mv_scale[lidx][r][0] = (int16_t)mx; mv_scale[lidx][r][1] = (int16_t)my;
                if(mecost < best_me) { best_me = mecost; refi_t = r; }
            }
            CU_PROF(3);
            const int mvx = mv_scale[lidx][refi_t][0], mvy = mv_scale[lidx][refi_t][1];
            // check_best_mvp: the loop compares against the cost of the initial index only (quirk q1)
            {
                xb200_bits_item bi = cu_bits_item(cu, 2, lidx, 0);
                bi.refi[lidx] = (int8_t)refi_t;
                bi.mvp_idx[0] = mvp_idx[lidx];
                bi.mvd[lidx][0] = (int16_t)(mvx - cand[mvp_idx[lidx]][0]); bi.mvd[lidx][1] = (int16_t)(mvy - cand[mvp_idx[lidx]][1]);
                const double ref_cost = __dmul_rn((double)cu_count<T>(H, bi, nullptr, ST_IN, tt), cu.lambda[0]);
                int          best = mvp_idx[lidx];
                for(int idx = 0; idx < 4; idx++) {
                    bool dup = false;
                    for(int t = idx - 1; t >= 0; t--) dup |= cand[idx][0] == cand[t][0] && cand[idx][1] == cand[t][1];
                    if(dup) continue;
                    bi.mvp_idx[0] = (uint8_t)idx;
                    bi.mvd[lidx][0] = (int16_t)(mvx - cand[idx][0]); bi.mvd[lidx][1] = (int16_t)(mvy - cand[idx][1]);
                    const double c = __dmul_rn((double)cu_count<T>(H, bi, nullptr, ST_IN, tt), cu.lambda[0]);
                    if(c < ref_cost) best = idx;
                }
                mvp_idx[lidx] = (uint8_t)best;
            }
            CU_PROF(4);
            if(tt == 0) {
                CuMode &M = H.md[lidx];
                M.refi[lidx] = (int8_t)refi_t; M.refi[1 - lidx] = -1;
                M.mv[lidx][0] = (int16_t)mvx; M.mv[lidx][1] = (int16_t)mvy; M.mv[1 - lidx][0] = M.mv[1 - lidx][1] = 0;
                M.mvd[lidx][0] = (int16_t)(mvx - cand[mvp_idx[lidx]][0]); M.mvd[lidx][1] = (int16_t)(mvy - cand[mvp_idx[lidx]][1]);
                M.mvd[1 - lidx][0] = M.mvd[1 - lidx][1] = 0;
                M.mvp_idx[lidx] = mvp_idx[lidx]; M.mvp_idx[1 - lidx] = 0;
            }
            team_sync<T>();
            const double c = cu_residue_rdo<L2>(Tm, pics, rt, sq, lidx, mvp_idx[0], mvp_idx[1], tt);
            if(lidx == 0) cost_l0 = c; else cost_l1 = c;
            if(c < cost_best) { cost_best = c; best_idx = lidx; cu_st_save<T>(H, ST_BEST, ST_MODE, tt); }
        }
        if(B) { // ---- analyze_bi: alternate the refined list, start MVs = the uni results, at most BI_ITER = 4 passes ----
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
                               reinterpret_cast<int16_t *>(Tm.TB), tt);
                for(int e = tt; e < NY; e += T)   // get_org_bi
                    Tm.org_bi[e] = (int16_t)(((int)Tm.org[0][(ptrdiff_t)(e >> L2) * Tm.so[0] + (e & (N - 1))] << 1) - (int)Tm.pred[e]);
                team_sync<T>();
                { const int8_t t = refi[lidx_ref]; refi[lidx_ref] = refi[lidx_cnd]; refi[lidx_cnd] = t; }
                { const int t = lidx_ref; lidx_ref = lidx_cnd; lidx_cnd = t; }
                const int mi = m_idx[lidx_ref];
                bool      changed = false;
                for(int r = 0; r < num_refp_cur; r++) {
                    int            mx, my;
                    const uint32_t mecost = cu_me<L2>(Tm, pics, sq, win_cap, err_flag, lidx_ref, r, num_refp_cur, 1, cu.mvp[lidx_ref][mi][0],
                                                      cu.mvp[lidx_ref][mi][1], mv_scale[lidx_ref][r][0], mv_scale[lidx_ref][r][1], mot_bits,
                                                      phase, mx, my, tt);
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
            CU_PROF(6);
            if(tt == 0) {
                CuMode &M = H.md[2];
                for(int l = 0; l < 2; l++) {
                    M.refi[l] = m_refi[l]; M.mvp_idx[l] = m_idx[l];
                    M.mv[l][0] = m_mv[l][0]; M.mv[l][1] = m_mv[l][1];
                    M.mvd[l][0] = (int16_t)(m_mv[l][0] - cu.mvp[l][m_idx[l]][0]); M.mvd[l][1] = (int16_t)(m_mv[l][1] - cu.mvp[l][m_idx[l]][1]);
                }
            }
            team_sync<T>();
            const double c = cu_residue_rdo<L2>(Tm, pics, rt, sq, 2, m_idx[0], m_idx[1], tt);
            if(c < cost_best) { cost_best = c; best_idx = 2; cu_st_save<T>(H, ST_BEST, ST_MODE, tt); }
        }
        cost_win = cost_best;
    }
    CU_PROF(7);
    // ---- winner: coefficients (dropped planes zeroed), reconstruction, XEVE_MODE fields, s_next_best ---------
    const CuMode &M = H.md[best_idx];
    int16_t      *gc = coef_out + cu.out_off, *gr = rec_out ? rec_out + cu.out_off : nullptr;
    if(best_idx == 3) {
        cu_predict<L2>(pics, cu, sq, M.refi[0], M.refi[1], M.mv[0][0], M.mv[0][1], M.mv[1][0], M.mv[1][1], Tm.pred, Tm.aux,
                       reinterpret_cast<int16_t *>(Tm.TB), tt);
        for(int e = tt; e < NP; e += T) {
            gc[e] = 0;
            if(gr) gr[e] = Tm.pred[e];
            if(pred_y_out && e < NY) pred_y_out[e] = Tm.pred[e];
        }
    }
    else {
        const int16_t *sc = Tm.scratch + (size_t)(3 * best_idx) * NP, *sr = sc + NP, *sp = sr + NP;
        const int      cbf = M.cbf;
        for(int e = tt; e < NP; e += T) {
            const int  c = e < NY ? 0 : (e < NY + NCH ? 1 : 2);
            const bool on = (cbf >> c) & 1;
            gc[e] = on ? __ldcg(sc + e) : (int16_t)0;
            if(gr) gr[e] = on ? __ldcg(sr + e) : __ldcg(sp + e);
            if(pred_y_out && e < NY) pred_y_out[e] = __ldcg(sp + e);
        }
    }
    if(tt == 0) {
        git->cost = cost_win; git->best_idx = (uint8_t)best_idx;
        for(int l = 0; l < 2; l++) {
            git->refi[l] = M.refi[l]; git->mvp_idx[l] = M.mvp_idx[l];
            git->mv[l][0] = M.mv[l][0]; git->mv[l][1] = M.mv[l][1]; git->mvd[l][0] = M.mvd[l][0]; git->mvd[l][1] = M.mvd[l][1];
        }
        git->nnz[0] = M.nnz[0]; git->nnz[1] = M.nnz[1]; git->nnz[2] = M.nnz[2];
    }
    if(cu.state_out >= 0 && tt < 32) {
        xb200_sbac &so = st_out[cu.state_out];
        for(int k = tt; k < XB200_CM_COUNT; k += 32) so.m[k] = H.st[ST_BEST][k];
        if(tt == 0) so.range = H.rg[ST_BEST];
    }
    team_sync<T>();
    CU_PROF(8);
}

template <int L2, int TEAMS>
__global__ void __launch_bounds__(CuCfg<L2, TEAMS>::CTA) k_analyze_cu(const PicDev *__restrict__ pics, xb200_cu_item *__restrict__ items,
                                                               const int32_t *__restrict__ order, int n, const xb200_rates *__restrict__ rates,
                                                               const xb200_sbac *__restrict__ st_in, xb200_sbac *__restrict__ st_out,
                                                               int16_t *__restrict__ coef_out, int16_t *__restrict__ rec_out,
                                                               int16_t *__restrict__ scratch, const int8_t *__restrict__ g_tm64, SeqDev sq,
                                                               int win_cap, int *__restrict__ err_flag)
{
    using Cf = CuCfg<L2, TEAMS>;
    constexpr int T = Cf::T, N = Cf::N, NY = Cf::NY, NCH = Cf::NCH, NP = Cf::NP;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int8_t *tm = reinterpret_cast<int8_t *>(smem_raw), *tmT = tm + 4096;
    const int      team = threadIdx.x / T, tt = threadIdx.x % T;
    unsigned char *tb = smem_raw + 8192 + (size_t)team * Cf::team_bytes(win_cap);
    CuTeam<L2>     Tm;
    Tm.H = reinterpret_cast<CuHdr *>(tb);
    Tm.org_bi = reinterpret_cast<int16_t *>(tb + Cf::HDR);
    Tm.me_area = tb + Cf::HDR + Cf::ORGBI;
    Tm.bar = reinterpret_cast<uint64_t *>(Tm.me_area);
    Tm.pred = reinterpret_cast<int16_t *>(Tm.me_area + 16);   // the residue working set overlays the search window, not the mbarrier
    Tm.aux = Tm.pred + NP;
    Tm.blk = Tm.aux + NP;
    Tm.TB = reinterpret_cast<int32_t *>(Tm.blk + NY);
    Tm.tm = tm; Tm.tmT = tmT;
    Tm.scratch = scratch + (size_t)(blockIdx.x * Cf::TEAMS + team) * Cf::SCRATCH;
    CuHdr &H = *Tm.H;
    for(int e = threadIdx.x; e < 4096; e += Cf::CTA) {
        const int8_t v = g_tm64[e];
        tm[e] = v;
        tmT[(e & 63) * 64 + (e >> 6)] = v;
    }
    if(tt == 0) mbar_init(reinterpret_cast<uint64_t *>(Tm.me_area), 1);
    __syncthreads();
    uint32_t phase = 0;

    for(int i = blockIdx.x * Cf::TEAMS + team; i < n; i += gridDim.x * Cf::TEAMS)
        analyze_cu_one<L2>(Tm, pics, &items[order[i]], rates, st_in, st_out, coef_out, rec_out, nullptr, sq, win_cap, err_flag, phase, tt);
}
