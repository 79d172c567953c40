// xb200_chain.cuh -- the decision pass of a whole picture as ONE persistent kernel (north_star: "the CTU tile loop in xeve_enc is
// lifted to the device").
//
//   reference: xeve_ctu_mt_core / the CTU loop of xeve_pic   src_base/xeve_enc.c:103-175, 327-404
//              mode_analyze_lcu, update_to_ctx_map            src_base/xeve_mode.c:2446-2608
//              mode_coding_tree                               src_base/xeve_mode.c:2007-2374
//              mode_coding_unit, mode_check_inter / _intra    src_base/xeve_mode.c:1170-1348
//              init_cu_data / copy_cu_data / copy_to_cu_data  src_base/xeve_mode.c:375-632, 868-1034
//              update_map_scu / clear_map_scu                 src_base/xeve_mode.c:1036-1155
//              mode_cpy_rec_to_ref                            src_base/xeve_mode.c:797-866
//              xeve_eco_split_mode (bit-count mode)           src_base/xeve_eco.c:1377-1429
//
// One 128-thread CTA runs one coder-state CHAIN at a time (what the reference calls a worker thread: CTU rows y, y + n, .. with
// threads = n; one chain per picture with threads = 1).  The CTA walks the chain's CTUs in raster order; inside a CTU it runs the
// quad-tree recursion as an explicit state machine, uniform over its 128 threads.  Every CU analysis is the SAME device function the
// work-list operators run (analyze_cu_one / analyze_cu_par of xb200_analyze*.cuh, intra_cu_one / intra_thr_one of xb200_intra.cuh) on
// a team made of the CTA's first 32 or all 128 threads; the bookkeeping between them (CU-data copies, frame maps, the picture under
// reconstruction) is spread over all 128.  128 threads, not 256: a chain keeps one to three warps busy most of the time, and the
// registers of idle warps are what limits the number of chains an SM can hold.
// A chain waits for the CTU above-right of its next CTU through a flag in global memory (src_base/xeve_enc.c:128-132); data written by
// other chains (frame maps, reconstructed samples) is read with L2 loads (ld.cg): the L1 of this SM may hold older copies of lines that
// straddle two CTUs.  The CTAs are the workers of the chain server at the end of this file: they take chains of whatever picture
// comes next from a device-side queue; nothing returns to the host until a picture is done.
#pragma once
#define XB200_DEVICE_FUNCS_ONLY
#define XB200_INTRA64_CAND_GLOBAL
#include "xb200_analyze.cuh"
#include "xb200_analyze_par.cuh"
#include "xb200_intra.cuh"
#include "xb200_had.cuh"

#define CH_T 128   // threads of a chain: the widest team (32x32 and, with XB200_T64 = 128, 64x64 CUs); three warps for the small CUs
#ifdef XB200_CHAIN_PROF
#define CU_PROF_T(k) do { if(t == 0) { const long long now_ = clock64(); atomicAdd(&g_chain_prof[(k)], (unsigned long long)(now_ - g_prof_last)); atomicAdd(&g_chain_prof[32 + (k)], 1ull); g_prof_last = now_; } } while(0)
#else
#define CU_PROF_T(k) do { } while(0)
#endif
#define CH_MAX_COST (1.7e+308)
enum { CH_SKIP = 0, CH_DIR = 1, CH_INTER = 2, CH_INTRA = 3 };

typedef xb200_state ChState;
constexpr int CH_STATE_WORDS = (int)sizeof(ChState) / 4;
constexpr int CH_REC_WORDS   = (int)sizeof(xb200_scu_rec) / 4;
static_assert(sizeof(ChState) % 4 == 0 && sizeof(xb200_scu_rec) == 36, "records are copied as 32-bit words");

struct ChCud {              // XEVE_CU_DATA of one quad-tree level: 4x4 units in raster order, row stride = block width in units
    xb200_scu_rec r[256];
    uint32_t      scu[256]; // map_scu words
    int16_t       rec[6144], coef[6144]; // Y | U | V of the block, stride = block width (the layout the CU analyses write)
};
struct alignas(16) ChainWs { // private working set of one chain (global memory, L2 resident)
    xb200_cu_item    cu;
    xb200_intra_item in;
    ChCud            best[5], temp[5];
    alignas(16) int16_t coef[2 * (6144 + 64)];
    alignas(16) int16_t rec_cu[2 * (6144 + 64)];
    alignas(16) int16_t pred_y[4096];
    alignas(16) int16_t cand64[5 * (4096 + 2)]; // candidate level arrays of a 64x64 intra CU (IN_CAND_SHARED)
    alignas(16) int16_t org_bi[4096];        // 2 * org - pred of the bi search (written once, copied into the search's own area)
    alignas(16) int16_t side[8 * 64 + 6 + 26];
    alignas(16) int16_t scratch[15 * 6144]; // candidate modes of the inter analysis: {coef, rec, pred} x 5
};

struct ChainPic {            // picture-level inputs of the kernel
    xb200_picture pp;
    int32_t  w, h, w_scu, h_scu, w_lcu, h_lcu, n_chain;
    int32_t  win_cap[4];     // search-window capacity (samples) per CU size 8 .. 64
    int32_t  par_stride;     // bytes between the three warps' regions of the parallel 8x8 / 16x16 inter analysis (0: serial analysis)
    int32_t  small_team;     // intra 4x4 (bit 0) / 8x8 (bit 1) CUs on a warp team instead of one thread (same results; tuning switch)
    PicDev   rec;            // the picture under reconstruction (PIC_MODE)
    uint32_t *map_scu;
    int8_t   *map_ipm, *map_refi;
    int16_t  *map_mv;
    uint8_t  *df_flags;      // per unit: bit 0 its left side is a CU edge, bit 1 its top side (what k_df_mark produces)
    const int16_t *col0, *col1; // refp[0][REFP_0 / REFP_1].map_mv
    xb200_scu_rec *scu_out;
    int16_t  *coef_out;
    ChState  *ctu_state;     // [n_lcu][2]
    double   *ctu_cost;
    int      *done;          // [n_lcu] completion flags
    unsigned long long *counts; // inter calls, intra calls
    ChainWs  *ws;
    xb200_cu_item    *cu_log;
    xb200_intra_item *intra_log;
    long long cu_cap, intra_cap;
};

// ---- the chain server: device-side work queue ---------------------------------------------------------------------------------------
// One long-lived grid of chain workers (as many CTAs as the device holds) serves the chains of EVERY picture in flight: the host
// publishes a picture as n_chain consecutive tasks, a worker takes the next ticket (FIFO), runs that chain and moves on.  One kernel
// per picture does not scale: the device runs kernels of different streams through a few hardware queues (8 by default, 32 at most),
// and a kernel that lives for a second at the head of a queue holds back whatever was launched behind it in that queue -- with a
// kernel per picture at most 8 pictures ran concurrently (61 of 444 chain slots busy, profiles/r02s12), and the short kernels around
// them (upload conversion, loop filter, border expansion) waited behind them too.
// Dead-lock freedom without admission control: chains of a picture spin on each other's flags, so all of them must run eventually --
// tickets are taken strictly in publication order, hence the earliest unfinished picture always has all its chains on workers (or
// next in line for the first free worker), completes, and frees its workers for the next one.
struct ChainTask { int32_t slot, chain; };          // picture slot, chain index
struct ChainPicTask {                               // one picture in flight
    ChainPic     P;
    const PicDev *pics;                             // the owning context's picture table
    SeqDev       sq;
    int         *err_flag;
    unsigned     seq;                               // generation written to *h_done when the last chain has finished
    unsigned     finished;                          // chains finished so far
    volatile unsigned *h_done;                      // host-mapped completion word of this slot
    unsigned long long t_first, t_last;             // %globaltimer when the first chain started / the last one finished
};
constexpr int CH_Q_TASKS = 1 << 14, CH_Q_SLOTS = 1024;
struct ChainQueue {
    unsigned head;                                  // next ticket (workers: atomicAdd)
    unsigned tail;                                  // tasks published so far (host writes, stream-ordered after the task records)
    int      stop;                                  // workers leave when they find no task and this is set
    int      pad_;
    ChainTask    tasks[CH_Q_TASKS];
    ChainPicTask slot[CH_Q_SLOTS];
};

struct ChShared {            // control block in shared memory
    uint64_t bar[4];         // search-window mbarriers: [0] serial analysis, [1..3] the three warps of the parallel small-CU analysis;
    uint32_t bphase[4];      // initialised once per kernel, never overlaid by working sets, parities kept here between CUs
    uint32_t bits, phase;
    int32_t  satd;
    // coder states of the tree walk (core->s_curr_best / s_next_best / s_temp_depth per level, src_base/xeve_type.h) and the rate tables
    // derived from them: copied at every node, so they live here and not in the chain's global working set
    ChState     curr[5], next[5], before[5], sdepth[5], chain;
    xb200_sbac  st_out;
    xb200_rates rates;
    int32_t     task_slot, task_chain;   // the task this worker is on (chain server)
    SeqDev      sq;                      // ... and copies of its picture's parameters
    ChainPic    P;
    uint8_t  zinv8[64];
    uint16_t thr_mb[IN_CM_N + 2], thr_mr[IN_CM_N + 2];
};
constexpr int CH_HDR  = ((int)sizeof(CuHdr) + 15) & ~15;
constexpr int CH_CTL  = ((int)sizeof(ChShared) + 15) & ~15;
constexpr int CH_TEAM_OFF = 8192 + CH_CTL;           // tm | tmT | control | team area
constexpr int CH_ME_OFF   = CH_HDR;                  // team area: CuHdr | mbarrier + working set
// max_cu: largest CU the picture can contain (intra teams are sized by max_cu_intra, the inter working sets by max_cu_inter): presets
// fast / medium never try a 64x64 intra CU, whose team is the largest working set of all
__host__ __device__ inline size_t chain_smem_bytes(const int32_t win_cap[4], int max_cu_intra = 64, int max_cu_inter = 64)
{
    // in B slices mode_check_intra follows the inter analysis for every CU size, whatever max_cu_intra says (it only bounds the tree of
    // an I slice): the intra team must cover the largest CU of either kind
    const int mi = max_cu_intra > max_cu_inter ? max_cu_intra : max_cu_inter;
    size_t need = 16 + (mi >= 64 ? sizeof(IntraTeam<6>) : mi >= 32 ? sizeof(IntraTeam<5>) : sizeof(IntraTeam<4>));
    for(int l2 = 3; l2 <= 6 && (1 << l2) <= max_cu_inter; l2++) {
        const size_t me = me_team_bytes(l2, win_cap[l2 - 3]);
        const size_t rs = 16 + (size_t)(l2 == 3 ? Res2Cfg<3>::TEAM_BYTES : l2 == 4 ? Res2Cfg<4>::TEAM_BYTES : l2 == 5 ? Res2Cfg<5>::TEAM_BYTES : Res2Cfg<6>::TEAM_BYTES);
        if(me > need) need = me;
        if(rs > need) need = rs;
    }
    return CH_TEAM_OFF + CH_ME_OFF + ((need + 15) & ~(size_t)15);
}
// stride of the three regions of the parallel small-CU analysis if they fit into the team area of `smem` bytes, else 0
__host__ __device__ inline int chain_par_stride(const int32_t win_cap[4], size_t smem)
{
    size_t r = cu_par_region_bytes<3>(win_cap[0]);
    const size_t r4 = cu_par_region_bytes<4>(win_cap[1]);
    if(r4 > r) r = r4;
    r = (r + 127) & ~(size_t)127;
    return CU_PAR_WARPS * r <= smem - CH_TEAM_OFF ? (int)r : 0;
}

// ---- small helpers, called by all CH_T threads -------------------------------------------------------------------------------------
XB_DEV void ch_copy32(void *dst, const void *src, int words, int t)
{
    uint32_t *d = static_cast<uint32_t *>(dst);
    const uint32_t *s = static_cast<const uint32_t *>(src);
    for(int i = t; i < words; i += CH_T) d[i] = s[i];
}
XB_DEV void ch_state_copy(ChState *dst, const ChState *src, int t) { ch_copy32(dst, src, CH_STATE_WORDS, t); }

XB_DEV void ch_cud_init(ChCud &d, int L, int t)        // init_cu_data: everything the chain reads later
{
    const int n = 1 << (2 * L);
    uint32_t *r = reinterpret_cast<uint32_t *>(d.r);
    for(int i = t; i < n * CH_REC_WORDS; i += CH_T) r[i] = 0;
    for(int i = t; i < n; i += CH_T) d.scu[i] = 0;
}
// copy_cu_data: level Ls block into level Ld block at (xoff, yoff) luma samples
XB_DEV void ch_cud_copy(ChCud &dst, int Ld, const ChCud &src, int Ls, int xoff, int yoff, int t)
{
    const int ns = 1 << Ls, nd = 1 << Ld, wsrc = 4 << Ls, wdst = 4 << Ld;
    const uint32_t *sr = reinterpret_cast<const uint32_t *>(src.r);
    uint32_t       *dr = reinterpret_cast<uint32_t *>(dst.r);
    for(int i = t; i < ns * ns * CH_REC_WORDS; i += CH_T) {
        const int u = i / CH_REC_WORDS, w = i - u * CH_REC_WORDS, j = u >> Ls, k = u & (ns - 1);
        dr[(((yoff >> 2) + j) * nd + (xoff >> 2) + k) * CH_REC_WORDS + w] = sr[i];
    }
    for(int u = t; u < ns * ns; u += CH_T) dst.scu[((yoff >> 2) + (u >> Ls)) * nd + (xoff >> 2) + (u & (ns - 1))] = src.scu[u];
    const int nys = wsrc * wsrc, nyd = wdst * wdst;
    for(int i = t; i < nys / 2; i += CH_T) {               // luma, two samples per word
        const int row = (2 * i) / wsrc, col = (2 * i) - row * wsrc, o = (yoff + row) * wdst + xoff + col;
        *reinterpret_cast<uint32_t *>(dst.rec + o)  = *reinterpret_cast<const uint32_t *>(src.rec + 2 * i);
        *reinterpret_cast<uint32_t *>(dst.coef + o) = *reinterpret_cast<const uint32_t *>(src.coef + 2 * i);
    }
    const int wcs = wsrc >> 1, wcd = wdst >> 1;
    for(int i = t; i < nys / 4; i += CH_T) {               // both chroma planes
        const int c = i >= nys / 8, e = 2 * (i - c * (nys / 8)), row = e / wcs, col = e - row * wcs;
        const int so = nys + c * (nys >> 2) + e, o = nyd + c * (nyd >> 2) + ((yoff >> 1) + row) * wcd + (xoff >> 1) + col;
        *reinterpret_cast<uint32_t *>(dst.rec + o)  = *reinterpret_cast<const uint32_t *>(src.rec + so);
        *reinterpret_cast<uint32_t *>(dst.coef + o) = *reinterpret_cast<const uint32_t *>(src.coef + so);
    }
}
XB_DEV void ch_clear_map(const ChainPic &P, int x, int y, int cuw, int t)                 // clear_map_scu
{
    const int w = (x + cuw > P.w ? P.w - x : cuw) >> 2, h = (y + cuw > P.h ? P.h - y : cuw) >> 2;
    for(int i = t; i < w * h; i += CH_T) P.map_scu[(size_t)((y >> 2) + i / w) * P.w_scu + (x >> 2) + i % w] = 0;
}
XB_DEV void ch_update_map(const ChainPic &P, const ChCud &b, int x, int y, int L, int t)  // update_map_scu
{
    const int cuw = 4 << L, n = 1 << L, w = (x + cuw > P.w ? P.w - x : cuw) >> 2, h = (y + cuw > P.h ? P.h - y : cuw) >> 2;
    for(int i = t; i < w * h; i += CH_T) {
        const int j = i / w, k = i - j * w, u = j * n + k;
        const size_t p = (size_t)((y >> 2) + j) * P.w_scu + (x >> 2) + k;
        const xb200_scu_rec &r = b.r[u];
        P.map_scu[p] = b.scu[u];
        P.map_ipm[p] = r.ipm;
        const uint32_t *mv = reinterpret_cast<const uint32_t *>(&r.mv[0][0]);   // records are 36 bytes: 4-byte aligned only
        reinterpret_cast<uint2 *>(P.map_mv)[p] = make_uint2(mv[0], mv[1]);
        P.map_refi[2 * p] = r.refi[0]; P.map_refi[2 * p + 1] = r.refi[1];
    }
}
XB_DEV void ch_rec_to_pic(const ChainPic &P, const ChCud &b, int x, int y, int L, int t)  // mode_cpy_rec_to_ref
{
    const int cuw = 4 << L, w = x + cuw > P.w ? P.w - x : cuw, h = y + cuw > P.h ? P.h - y : cuw, ny = cuw * cuw;
    int16_t *py = P.rec.p[0] + (ptrdiff_t)y * P.rec.s[0] + x;
    for(int i = t; i < (w * h) >> 1; i += CH_T) {
        const int row = (2 * i) / w, col = 2 * i - row * w;
        *reinterpret_cast<uint32_t *>(py + (ptrdiff_t)row * P.rec.s[0] + col) = *reinterpret_cast<const uint32_t *>(b.rec + row * cuw + col);
    }
    const int wc = w >> 1, hc = h >> 1, cw = cuw >> 1;
    for(int i = t; i < wc * hc; i += CH_T) {                // wc * hc / 2 words per plane, two planes
        const int c = i >= (wc * hc) >> 1, e = 2 * (i - c * ((wc * hc) >> 1)), row = e / wc, col = e - row * wc;
        int16_t *pc = P.rec.p[1 + c] + (ptrdiff_t)((y >> 1) + row) * P.rec.s[1 + c] + (x >> 1) + col;
        *reinterpret_cast<uint32_t *>(pc) = *reinterpret_cast<const uint32_t *>(b.rec + ny + c * (ny >> 2) + row * cw + col);
    }
}
// copy_to_cu_data into cu_data_temp[L]: the decision of one CU on every unit it covers, its coefficients and reconstruction
XB_DEV void ch_store_cu(ChCud &T, int L, int mode, int ipm, int tile_qp, const xb200_cu_item *cu, const int32_t *nnz, const int16_t *coef,
                        const int16_t *rec, int t)
{
    const int n = 1 << (2 * L), ny = 16 << (2 * L);
    const uint32_t word = ((uint32_t)tile_qp << 16) | ((uint32_t)(mode == CH_INTRA) << 15) | (1u << 31) | ((uint32_t)(mode == CH_SKIP) << 23);
    for(int i = t; i < n; i += CH_T) {
        xb200_scu_rec &r = T.r[i];
        r.mode = (uint8_t)mode; r.log2 = (uint8_t)(L + 2);
        r.nnz[0] = nnz[0]; r.nnz[1] = nnz[1]; r.nnz[2] = nnz[2];
        T.scu[i] = word;
        if(mode == CH_INTRA) {
            r.ipm = (int8_t)ipm; r.refi[0] = r.refi[1] = -1;
            r.mv[0][0] = r.mv[0][1] = r.mv[1][0] = r.mv[1][1] = 0;
        }
        else {
#pragma unroll
            for(int l = 0; l < 2; l++) {
                r.refi[l] = cu->refi[l]; r.mvp_idx[l] = cu->mvp_idx[l];
                r.mv[l][0] = cu->mv[l][0]; r.mv[l][1] = cu->mv[l][1]; r.mvd[l][0] = cu->mvd[l][0]; r.mvd[l][1] = cu->mvd[l][1];
            }
        }
    }
    for(int i = t; i < (3 * ny) >> 2; i += CH_T) {          // 3/2 ny samples, two per word
        reinterpret_cast<uint32_t *>(T.rec)[i]  = reinterpret_cast<const uint32_t *>(rec)[i];
        reinterpret_cast<uint32_t *>(T.coef)[i] = reinterpret_cast<const uint32_t *>(coef)[i];
    }
}
// xeve_eco_split_mode in bit-count mode on thread 0: one bin on ctx.split_cu_flag[0]; returns the bits through shared memory
XB_DEV uint32_t ch_split_flag(ChState *st, int cuw, int split, ChShared &S, int t)
{
    if(cuw < 8) return 0;
    if(t == 0) {
        Cabac c;
        c.range = st->s.range; c.bits = 0; c.m = &st->split;
        cb_bin(c, 0, split);
        st->s.range = c.range;
        S.bits = c.bits;
    }
    __syncthreads();
    const uint32_t b = S.bits;
    __syncthreads();
    return b;
}
// xeve_rdoq_bit_est (src_base/xeve_mode.c:326-373) of one state
XB_DEV void ch_rates(const ChState &st, xb200_rates &o, int t)
{
    const uint16_t *m = st.s.m;
    if(t < 24) {
        for(int b = 0; b < 2; b++) { o.run[t][b] = rate_of(m[XB200_CM_RUN + t], b); o.level[t][b] = rate_of(m[XB200_CM_LEVEL + t], b); }
    }
    else if(t < 26) {
        for(int b = 0; b < 2; b++) o.last[t - 24][b] = rate_of(m[XB200_CM_LAST + t - 24], b);
    }
    else if(t == 26) {
        for(int b = 0; b < 2; b++) {
            o.cbf_all[b] = rate_of(m[XB200_CM_CBF_ALL], b); o.cbf_luma[b] = rate_of(m[XB200_CM_CBF_LUMA], b);
            o.cbf_cb[b] = rate_of(m[XB200_CM_CBF_CB], b); o.cbf_cr[b] = rate_of(m[XB200_CM_CBF_CR], b);
        }
    }
}
// xeve_get_avail_inter + xeve_get_motion + xeve_get_mv_dir for list l of the CU (body of k_mvp, frame maps read from L2)
XB_DEV void ch_mvp(const ChainPic &P, xb200_cu_item &cu, int x, int y, int log2, int l)
{
    const int xs = x >> 2, ys = y >> 2, w = P.w_scu, h = P.h_scu, scuw = (1 << log2) >> 2, scuh = scuw, scup = xs + ys * w;
    const uint32_t *ms = P.map_scu;
    auto M   = [&](int p) { return __ldcg(ms + p); };
    auto COD = [&](int p) { return (int)((M(p) >> 31) & 1); };
    auto IF  = [&](int p) { return (int)((M(p) >> 15) & 1); };
    auto IBC = [&](int p) { return (int)((M(p) >> 26) & 1); };
    unsigned av = 0;
    if(xs > 0 && !IF(scup - 1) && COD(scup - 1) && !IBC(scup - 1)) av |= 1u << 1;
    if(ys > 0) {
        if(!IF(scup - w) && !IBC(scup - w)) av |= 1u << 0;
        if(xs + scuw < w && ((M(scup - w + scuw) >> 15) & 0x10001) == 0x10000 && COD(scup - w + scuw)) av |= 1u << 6;
    }
    (void)h; (void)scuh;
    auto mvat = [&](const int16_t *m, int p, int list, int c) { return __ldcg(m + ((size_t)p * 2 + list) * 2 + c); };
    const int      nb[3] = {scup - 1, scup - w, scup - w + scuw};
    const unsigned need[3] = {1u << 1, 1u << 0, 1u << 6};
#pragma unroll
    for(int k = 0; k < 3; k++) {
        cu.refi_pred[l][k] = 0;
        const bool ok = (av & need[k]) != 0;
        cu.mvp[l][k][0] = ok ? mvat(P.map_mv, nb[k], l, 0) : (int16_t)1;
        cu.mvp[l][k][1] = ok ? mvat(P.map_mv, nb[k], l, 1) : (int16_t)1;
    }
    const int16_t *col = l ? P.col1 : P.col0;
    cu.refi_pred[l][3] = 0;
    cu.mvp[l][3][0] = col[((size_t)scup * 2) * 2 + 0];
    cu.mvp[l][3][1] = col[((size_t)scup * 2) * 2 + 1];
    if(l == 0 && P.pp.slice_type == 0) {   // temporal direct: colocated MV of the list-1 reference at the CU's bottom-right unit
        const int br = scup + (scuw - 1) + (scuh - 1) * w;
        const int mx = P.col1[((size_t)br * 2) * 2 + 0], my = P.col1[((size_t)br * 2) * 2 + 1];
        const int dco = P.pp.ref_poc[1][0] - P.pp.col_list_poc0, d0 = P.pp.poc - P.pp.ref_poc[0][0], d1 = P.pp.ref_poc[1][0] - P.pp.poc;
        if(dco == 0) { cu.mv_dir[0][0] = cu.mv_dir[0][1] = cu.mv_dir[1][0] = cu.mv_dir[1][1] = 0; }
        else {
            cu.mv_dir[0][0] = (int16_t)(d0 * mx / dco); cu.mv_dir[0][1] = (int16_t)(d0 * my / dco);
            cu.mv_dir[1][0] = (int16_t)(-d1 * mx / dco); cu.mv_dir[1][1] = (int16_t)(-d1 * my / dco);
        }
    }
}
// xeve_get_avail_intra + xeve_get_nbr (Y, U, V) + xeve_get_mpm by warp 0 (body of k_intra_nbr; picture and maps read from L2)
XB_DEV void ch_intra_nbr(const ChainPic &P, int x, int y, int log2, int bd, int16_t *out, uint8_t *mpm, int lane)
{
    const int N = 1 << log2, scuw = N >> 2, scuh = scuw, xs = x >> 2, ys = y >> 2, w_scu = P.w_scu, h_scu = P.h_scu, cip = P.pp.cip;
    const int scup = xs + ys * w_scu, half = 1 << (bd - 1);
    const uint32_t *ms = P.map_scu;
    auto COD = [&](int q) { return (int)((__ldcg(ms + q) >> 31) & 1); };
    auto IFL = [&](int q) { return (int)((__ldcg(ms + q) >> 15) & 1); };
    unsigned av = 0;
    if(xs > 0 && COD(scup - 1)) av |= 1u << 1;
    if(ys > 0) {
        av |= (1u << 0) | (1u << 9);
        if(xs > 0 && COD(scup - w_scu - 1)) av |= 1u << 5;
    }
    (void)scuh; (void)h_scu;
    const bool ul_ok = ((av >> 5) & 1) && (!cip || IFL(scup - w_scu - 1));
    if(lane == 0) {
        int ipm_l = 0, ipm_u = 0;
        if(xs > 0 && IFL(scup - 1) && COD(scup - 1)) ipm_l = __ldcg(P.map_ipm + scup - 1) + 1;
        if(ys > 0 && IFL(scup - w_scu) && COD(scup - w_scu)) ipm_u = __ldcg(P.map_ipm + scup - w_scu) + 1;
        ipm_l = min(max(ipm_l, 0), 5); ipm_u = min(max(ipm_u, 0), 5);
#pragma unroll
        for(int k = 0; k < 5; k++) mpm[k] = c_mpm_tbl[ipm_l][ipm_u][k];
    }
    for(int e = lane; e < 8 * N + 6; e += 32) {
        int c, r = e;
        if(r < 2 * (2 * N + 1)) c = 0;
        else { r -= 2 * (2 * N + 1); c = 1 + r / (2 * (N + 1)); r %= 2 * (N + 1); }
        const int  nn = c ? N >> 1 : N, per = 2 * nn + 1, unit = c ? 2 : 4;
        const bool is_up = r >= per;
        const int  k = (is_up ? r - per : r) - 1; // -1 .. 2nn-1
        const int16_t *src = P.rec.p[c] + (ptrdiff_t)(c ? y >> 1 : y) * P.rec.s[c] + (c ? x >> 1 : x);
        const ptrdiff_t s = P.rec.s[c];
        int v = half;
        if(k < 0) { if(ul_ok) v = __ldcg(src - s - 1); }
        else {
            const int u = k / unit;
            if(is_up) {
                if(ys > 0 && xs + u < w_scu && COD(scup - w_scu + u) && (!cip || IFL(scup - w_scu + u))) v = __ldcg(src - s + k);
            }
            else if(xs > 0 && ys + u < h_scu && COD(scup - 1 + u * w_scu) && (!cip || IFL(scup - 1 + u * w_scu))) v = __ldcg(src + (ptrdiff_t)k * s - 1);
        }
        out[e] = (int16_t)v;
    }
}
// core->inter_satd = xeve_satd_16b(org, mi->pred_y_best) (src_base/xeve_mode.c:1247-1258): 8x8 Hadamard tiles, one per thread
XB_DEV uint32_t ch_satd(const int16_t *org, int so, const int16_t *pred, int cuw, int bd, ChShared &S, int t)
{
    if(t == 0) S.satd = 0;
    __syncthreads();
    const int tw = cuw >> 3, nt = tw * tw;
    int sum = 0;
    for(int k = t; k < nt; k += CH_T) {
        const int tx = (k % tw) * 8, ty = (k / tw) * 8;
        sum += had_tile_dev<8>(org + (ptrdiff_t)ty * so + tx, so, pred + ty * cuw + tx, cuw);
    }
    if(sum) atomicAdd(&S.satd, sum);
    __syncthreads();
    const uint32_t v = (uint32_t)(S.satd >> (bd - 8));
    __syncthreads();
    return v;
}

// ---- the two CU analyses on a team made of the CTA's first threads ------------------------------------------------------------
template <int L2>
__device__ __noinline__ void ch_inter(unsigned char *team, const int8_t *tm, const int8_t *tmT, const PicDev *__restrict__ pics, ChainWs *ws,
                                      int L, const SeqDev &sq, int win_cap, int *err_flag, ChShared &S, int t)
{
    using R = Res2Cfg<L2>;
    constexpr int T = R::T, NP = R::PRED, NY = R::N * R::N;
    if(t < T) {
        CuTeam<L2> Tm;
        Tm.H       = reinterpret_cast<CuHdr *>(team);
        Tm.org_bi  = ws->org_bi;
        Tm.me_area = team + CH_ME_OFF;
        Tm.bar     = &S.bar[0];
        Tm.pred    = reinterpret_cast<int16_t *>(Tm.me_area + 16);
        Tm.aux     = Tm.pred + NP;
        Tm.blk     = Tm.aux + NP;
        Tm.TB      = reinterpret_cast<int32_t *>(Tm.blk + NY);
        Tm.tm = tm; Tm.tmT = tmT;
        Tm.scratch = ws->scratch;
        uint32_t phase = S.bphase[0];
        analyze_cu_one<L2>(Tm, pics, &ws->cu, &S.rates, &S.curr[L].s, &S.st_out, ws->coef, ws->rec_cu, ws->pred_y, sq, win_cap, err_flag,
                           phase, t);
        if(t == 0) S.bphase[0] = phase;
    }
}
// 8x8 / 16x16: the candidate modes of the CU on three warps (xb200_analyze_par.cuh); the regions overlay the whole team area
template <int L2>
__device__ __noinline__ void ch_inter_par(unsigned char *team, const int8_t *tm, const int8_t *tmT, const PicDev *__restrict__ pics, ChainWs *ws,
                                          int L, const SeqDev &sq, int win_cap, int stride, int *err_flag, ChShared &S, int t)
{
    if(t < CU_PAR_WARPS * 32)
        analyze_cu_par<L2>(team, stride, S.bar + 1, S.bphase + 1, tm, tmT, pics, &ws->cu, &S.rates, &S.curr[L].s, &S.st_out, ws->coef, ws->rec_cu,
                           ws->pred_y, ws->scratch, sq, win_cap, err_flag, t);
}
template <int L2>
__device__ __noinline__ void ch_intra_team(unsigned char *team, const int8_t *tm, const int8_t *tmT, const PicDev *__restrict__ pics, ChainWs *ws,
                                           int L, int16_t *coef, int16_t *rec, const SeqDev &sq, ChShared &S, int t)
{
    if(t < IntraCfg<L2>::T) {
        IntraTeam<L2> &M = *reinterpret_cast<IntraTeam<L2> *>(team + CH_ME_OFF + 16);
        intra_cu_one<L2>(M, tm, tmT, pics, ws->in, &S.rates, &S.curr[L].s, &S.st_out, ws->side, coef, rec, sq, t, ws->cand64);
    }
}
template <int L2>
__device__ __noinline__ void ch_intra_thr(const PicDev *__restrict__ pics, ChainWs *ws, int L, int16_t *coef, int16_t *rec, const SeqDev &sq,
                                          ChShared &S, int t)
{
    if(t == 0) intra_thr_one<L2, 1>(pics, ws->in, &S.rates, &S.curr[L].s, &S.st_out, ws->side, coef, rec, sq, S.thr_mb, S.thr_mr, S.zinv8);
}

// mode_coding_unit for the CU (x, y, 4 << L): returns the best cost; cu_mode / dist_cu as the reference's core->cu_mode / dist_cu_best;
// leaves the decision in temp[L] and the coder state after the CU in next[L]
__device__ __noinline__ double ch_unit(const ChainPic &P, const PicDev *__restrict__ pics, ChainWs *ws, unsigned char *team, const int8_t *tm,
                                       const int8_t *tmT, const SeqDev &sq, int *err_flag, ChShared &S, int x, int y, int L, int &cu_mode,
                                       int &dist_cu_best, long long &n_inter, long long &n_intra, int t)
{
    const xb200_picture &pp = P.pp;
    const int log2 = L + 2, cuw = 1 << log2, ny = cuw * cuw, B = pp.slice_type == 0;
    double    cost_best = CH_MAX_COST;
    int       nnz0 = 0, nnz1 = 0, nnz2 = 0;
    ch_rates(S.curr[L], S.rates, t);                       // mode_cu_init -> xeve_rdoq_bit_est
    cu_mode = CH_INTRA;
    if(pp.slice_type != 2 && L >= 1) {                         // mode_check_inter
        xb200_cu_item &cu = ws->cu;
        if(t == 0) {
            cu.poc = pp.poc; cu.cur_pic = pp.cur_pic; cu.x = (int16_t)x; cu.y = (int16_t)y; cu.log2_cuw = cu.log2_cuh = (uint8_t)log2;
            cu.slice_type = (uint8_t)pp.slice_type; cu.ctx_skip = 0; cu.ctx_pred_mode = 0; cu.all_preds = 1;
            cu.max_search_range = pp.max_search_range; cu.lambda_mv = pp.lambda_mv;
            for(int l = 0; l < 2; l++) {
                cu.num_refp[l] = (uint8_t)pp.num_refp[l];
                for(int r = 0; r < XB200_MAX_REFP; r++) { cu.ref_pic[l][r] = pp.ref_pic[l][r]; cu.ref_poc[l][r] = pp.ref_poc[l][r]; }
            }
            for(int i = 0; i < 3; i++) { cu.qp[i] = (uint8_t)pp.qp[i]; cu.lambda[i] = pp.lambda[i]; }
            cu.pad0_ = 0;
            cu.dist_chroma_weight[0] = pp.dist_chroma_weight[0]; cu.dist_chroma_weight[1] = pp.dist_chroma_weight[1];
            cu.rate_idx = 0; cu.state_in = 0; cu.state_out = 0; cu.out_off = 0;
            cu.coef_hash = cu.rec_hash = 0; cu.me_first = cu.me_cnt = 0;
            if(!B) {
                for(int k = 0; k < 4; k++) { cu.mvp[1][k][0] = cu.mvp[1][k][1] = 0; cu.refi_pred[1][k] = 0; }
                cu.mv_dir[0][0] = cu.mv_dir[0][1] = cu.mv_dir[1][0] = cu.mv_dir[1][1] = 0;
            }
        }
        if(t == 0) { CH_DBG(0, x); CH_DBG(1, y); CH_DBG(2, log2); CH_DBG(3, (int)n_inter); }
        if(t == 32 || (t == 64 && B)) ch_mvp(P, cu, x, y, log2, t == 64);
        __syncthreads();
        CU_PROF_T(10);
#ifdef XB200_CHAIN_PROF
        const long long prof_start = clock64();   // 18 .. 21: whole inter analysis by CU size (the phases 0 .. 8, 12 .. 14 lie inside)
#endif
        const int cap = P.win_cap[log2 - 3];
        switch(log2) {
        case 3:
            if(P.par_stride) ch_inter_par<3>(team, tm, tmT, pics, ws, L, sq, cap, P.par_stride, err_flag, S, t);
            else ch_inter<3>(team, tm, tmT, pics, ws, L, sq, cap, err_flag, S, t);
            break;
        case 4:
            if(P.par_stride) ch_inter_par<4>(team, tm, tmT, pics, ws, L, sq, cap, P.par_stride, err_flag, S, t);
            else ch_inter<4>(team, tm, tmT, pics, ws, L, sq, cap, err_flag, S, t);
            break;
        case 5: ch_inter<5>(team, tm, tmT, pics, ws, L, sq, cap, err_flag, S, t); break;
        default: ch_inter<6>(team, tm, tmT, pics, ws, L, sq, cap, err_flag, S, t); break;
        }
        __syncthreads();
#ifdef XB200_CHAIN_PROF
        if(t == 0) { atomicAdd(&g_chain_prof[15 + log2], (unsigned long long)(clock64() - prof_start)); atomicAdd(&g_chain_prof[32 + 15 + log2], 1ull); }
#endif
        if(P.cu_log && P.n_chain == 1 && n_inter < P.cu_cap) ch_copy32(&P.cu_log[n_inter], &cu, (int)sizeof(xb200_cu_item) / 4, t);
        n_inter++;
        const int bi = cu.best_idx;
        cu_mode = bi == 3 ? CH_SKIP : bi == 4 ? CH_DIR : CH_INTER;
        nnz0 = cu.nnz[0]; nnz1 = cu.nnz[1]; nnz2 = cu.nnz[2];
        const double c = cu.cost;
        // SBAC_STORE(s_next_best, s_temp_best): the models of the inter syntax; ctx.intra_dir / split stay those of s_curr_best
        ch_state_copy(&S.next[L], &S.curr[L], t);
        __syncthreads();
        ch_copy32(&S.next[L].s, &S.st_out, (int)sizeof(xb200_sbac) / 4, t);
        if(c < cost_best) {
            cost_best = c;
            ch_store_cu(ws->temp[L], L, cu_mode, 0, pp.tile_qp, &cu, cu.nnz, ws->coef, ws->rec_cu, t);
        }
        __syncthreads();
    }
    if(pp.slice_type == 2 || nnz0 || nnz1 || nnz2 || cost_best == CH_MAX_COST) {   // mode_check_intra
        const PicDev &o = pics[pp.cur_pic];
        xb200_intra_item &it = ws->in;
        dist_cu_best = 0x7fffffff;
        uint32_t inter_satd = 0xffffffffu;
        if(cost_best != CH_MAX_COST) inter_satd = ch_satd(o.p[0] + (ptrdiff_t)y * o.s[0] + x, o.s[0], ws->pred_y, cuw, sq.bd, S, t);
        if(t < 32) ch_intra_nbr(P, x, y, log2, sq.bd, ws->side, it.mpm, t);
        if(t == 32) {
            it.poc = pp.poc; it.cur_pic = pp.cur_pic; it.x = (int16_t)x; it.y = (int16_t)y; it.log2_cuw = it.log2_cuh = (uint8_t)log2;
            it.slice_type = (uint8_t)pp.slice_type; it.ctx_skip = 0; it.ctx_pred_mode = 0; it.all_preds = 1;
            for(int i = 0; i < 3; i++) { it.qp[i] = (uint8_t)pp.qp[i]; it.lambda[i] = pp.lambda[i]; }
            it.pad0_[0] = it.pad0_[1] = 0;
            it.inter_satd = inter_satd;
            it.rate_idx = 0; it.state_in = 0; it.state_out = 0;
            it.cm_ipm_in[0] = S.curr[L].ipm[0]; it.cm_ipm_in[1] = S.curr[L].ipm[1];
            it.cm_ipm_out[0] = it.cm_ipm_out[1] = 0;
            it.sqrt_lambda0 = pp.sqrt_lambda0;
            it.dist_chroma_weight[0] = pp.dist_chroma_weight[0]; it.dist_chroma_weight[1] = pp.dist_chroma_weight[1];
            it.nb_off = 0; it.out_off = 0;
            it.coef_hash = it.rec_hash = 0;
        }
        __syncthreads();
        // the inter winner's reconstruction already sits in cu_data_temp; the intra trial works in its own buffers
        int16_t *coef_i = ws->coef + 3 * ny / 2 + 64, *rec_i = ws->rec_cu + 3 * ny / 2 + 64;
        switch(log2) {
        case 2:
            if(P.small_team & 1) ch_intra_team<2>(team, tm, tmT, pics, ws, L, coef_i, rec_i, sq, S, t);
            else ch_intra_thr<2>(pics, ws, L, coef_i, rec_i, sq, S, t);
            break;
        case 3:
            if(P.small_team & 2) ch_intra_team<3>(team, tm, tmT, pics, ws, L, coef_i, rec_i, sq, S, t);
            else ch_intra_thr<3>(pics, ws, L, coef_i, rec_i, sq, S, t);
            break;
        case 4: ch_intra_team<4>(team, tm, tmT, pics, ws, L, coef_i, rec_i, sq, S, t); break;
        case 5: ch_intra_team<5>(team, tm, tmT, pics, ws, L, coef_i, rec_i, sq, S, t); break;
        default: ch_intra_team<6>(team, tm, tmT, pics, ws, L, coef_i, rec_i, sq, S, t); break;
        }
        __syncthreads();
        CU_PROF_T(9);
        if(P.intra_log && P.n_chain == 1 && n_intra < P.intra_cap) ch_copy32(&P.intra_log[n_intra], &it, (int)sizeof(xb200_intra_item) / 4, t);
        n_intra++;
        const double c = it.cost;
        if(c < cost_best) {
            cost_best = c;
            cu_mode = CH_INTRA;
            dist_cu_best = it.dist_cu;
            ch_state_copy(&S.next[L], &S.curr[L], t);
            __syncthreads();
            ch_copy32(&S.next[L].s, &S.st_out, (int)sizeof(xb200_sbac) / 4, t);
            if(t == CH_T - 1) { S.next[L].ipm[0] = it.cm_ipm_out[0]; S.next[L].ipm[1] = it.cm_ipm_out[1]; }
            ch_store_cu(ws->temp[L], L, CH_INTRA, it.ipm[0], pp.tile_qp, nullptr, it.nnz, coef_i, rec_i, t);
        }
        __syncthreads();
    }
    return cost_best;
}

// the decision pass of one chain of one picture (P, sq: the worker's shared-memory copies)
__device__ __forceinline__ void chain_picture(const PicDev *__restrict__ pics, const ChainPic &P, const SeqDev &sq, int *__restrict__ err_flag,
                                              int chain, unsigned char *smem_raw)
{
    int8_t        *tm = reinterpret_cast<int8_t *>(smem_raw), *tmT = tm + 4096;
    ChShared      &S = *reinterpret_cast<ChShared *>(smem_raw + 8192);
    unsigned char *team = smem_raw + CH_TEAM_OFF;
    const int      t = threadIdx.x;
    ChainWs       *ws = P.ws + chain;
    const xb200_picture &pp = P.pp;
#ifdef XB200_CHAIN_PROF
    if(t == 0) g_prof_last = clock64();
#endif
    // xeve_sbac_reset with cm_init off: every model PROB_INIT, range 16384
    if(t < XB200_CM_COUNT) S.chain.s.m[t] = 512;
    if(t == XB200_CM_COUNT) { S.chain.s.range = 16384; S.chain.ipm[0] = S.chain.ipm[1] = S.chain.split = 512; S.chain.pad_ = 0; }
    __syncthreads();

    const int    intra_slice = pp.slice_type == 2;
    const int    check_max = intra_slice ? pp.max_cu_intra : pp.max_cu_inter, check_min = intra_slice ? pp.min_cu_intra : pp.min_cu_inter;
    const double lambda0 = pp.lambda[0];
    const int    ecu_depth = (pp.poc % 2) ? 2 : 4;                          // ENC_ECU_ADAPTIVE (src_base/xeve_mode.c:2169-2176)
    long long    n_inter = 0, n_intra = 0;                                  // CU analyses of this chain (uniform over the CTA)

    for(int row = chain; row < P.h_lcu; row += P.n_chain)
        for(int xl = 0; xl < P.w_lcu; xl++) {
            const int lcu = row * P.w_lcu + xl, xc = xl << 6, yc = row << 6;
            if(row > 0 && P.n_chain > 1) {                                   // wait for the CTU above-right (src_base/xeve_enc.c:128-132)
                if(t == 0) {
                    volatile int *f = P.done + (row - 1) * P.w_lcu + min(xl + 1, P.w_lcu - 1);
                    while(*f == 0) __nanosleep(200);
                    __threadfence();
                }
                __syncthreads();
            }
            ch_cud_init(ws->best[4], 4, t); ch_cud_init(ws->temp[4], 4, t);  // mode_init_lcu
            ch_state_copy(&S.curr[4], &S.chain, t);
            if(P.ctu_state) ch_state_copy(&P.ctu_state[2 * lcu], &S.chain, t);
            __syncthreads();

            // ---- mode_coding_tree as a state machine; node state per level ----
            int    nx[5], nyy[5], ncud[5], npart[5];
            double cbest[5], ctemp[5];
            int    L = 4;
            nx[4] = xc; nyy[4] = yc; ncud[4] = 0;
            enum { ENTER, NEXT_PART, FINISH } st = ENTER;
            for(;;) {
                if(st == ENTER) {
                    const int x0 = nx[L], y0 = nyy[L], cud = ncud[L], cuw = 4 << L;
                    const int boundary = !(x0 + cuw <= P.w && y0 + cuw <= P.h);
                    int       next_split = 1, cu_mode = 0, dist_cu = 0;
                    double    cost_best = CH_MAX_COST;
                    for(int i = t; i < CH_STATE_WORDS; i += CH_T) reinterpret_cast<uint32_t *>(&S.sdepth[L])[i] = 0;
                    ch_state_copy(&S.before[L], &S.curr[L], t);
                    __syncthreads();
                    if(!boundary && cuw <= check_max) {
                        double cost_temp = 0.0;
                        if(cuw > 4) cost_temp = __dadd_rn(cost_temp, __dmul_rn((double)ch_split_flag(&S.curr[L], cuw, 0, S, t), lambda0));
                        ch_cud_init(ws->temp[L], L, t);
                        ch_clear_map(P, x0, y0, cuw, t);
                        __syncthreads();
                        int          um = 0, ud = 0;
                        const double cu_cost = ch_unit(P, pics, ws, team, tm, tmT, sq, err_flag, S, x0, y0, L, um, ud, n_inter, n_intra, t);
                        const double cost_dqp = __dadd_rn(cost_temp, cu_cost);
                        if(cost_best > cost_dqp) {
                            cu_mode = um; dist_cu = ud;
                            ch_cud_copy(ws->best[L], L, ws->temp[L], L, 0, 0, t);
                            cost_best = cost_dqp;
                            ch_state_copy(&S.sdepth[L], &S.next[L], t);
                            __syncthreads();
                            ch_rec_to_pic(P, ws->best[L], x0, y0, L, t);
                        }
                        __syncthreads();
                    }
                    if(cost_best != CH_MAX_COST && cud >= ecu_depth && cu_mode == CH_SKIP) next_split = 0;
                    if(cost_best != CH_MAX_COST && intra_slice) {
                        const int log2 = L + 2, dist_th = 1 << (2 * log2 + 7);
                        if(dist_cu < dist_th) {
                            const int inc = (2 * log2 >= 6 ? 2 : 0) + 8;
                            if((double)dist_cu < __dmul_rn(lambda0, (double)inc)) next_split = 0;
                        }
                    }
                    cbest[L] = cost_best;
                    if(cuw > 4 && next_split && cuw > check_min) {
                        ch_cud_init(ws->temp[L], L, t);
                        ch_clear_map(P, x0, y0, cuw, t);
                        ch_state_copy(&S.curr[L], &S.before[L], t);
                        __syncthreads();
                        ctemp[L] = __dadd_rn(0.0, __dmul_rn((double)ch_split_flag(&S.curr[L], cuw, 1, S, t), lambda0));
                        npart[L] = 0;
                        st = NEXT_PART;
                    }
                    else st = FINISH;
                }
                else if(st == NEXT_PART) {
                    const int x0 = nx[L], y0 = nyy[L], half = 2 << L;
                    bool      down = false;
                    while(npart[L] < 4) {
                        const int part = npart[L], xp = x0 + (part & 1) * half, yp = y0 + (part >> 1) * half;
                        if(xp < P.w && yp < P.h) {
                            ch_state_copy(&S.curr[L - 1], part == 0 ? &S.curr[L] : &S.next[L - 1], t);
                            __syncthreads();
                            nx[L - 1] = xp; nyy[L - 1] = yp; ncud[L - 1] = ncud[L] + 2;   // a quad split is two levels of the split tree
                            L--;
                            down = true;
                            break;
                        }
                        npart[L]++;
                    }
                    if(down) { st = ENTER; continue; }
                    if(__dadd_rn(cbest[L], -0.0001) > ctemp[L]) {
                        ch_cud_copy(ws->best[L], L, ws->temp[L], L, 0, 0, t);
                        cbest[L] = ctemp[L];
                        ch_state_copy(&S.sdepth[L], &S.next[L - 1], t);
                        __syncthreads();
                    }
                    st = FINISH;
                }
                else { // FINISH
                    __syncthreads();
                    ch_rec_to_pic(P, ws->best[L], nx[L], nyy[L], L, t);
                    ch_state_copy(&S.next[L], &S.sdepth[L], t);
                    __syncthreads();
                    const double ret = cbest[L] > CH_MAX_COST ? CH_MAX_COST : cbest[L];
                    if(L == 4) { if(P.ctu_cost && t == 0) P.ctu_cost[lcu] = ret; break; }
                    const int Lp = L + 1;
                    ctemp[Lp] = __dadd_rn(ctemp[Lp], ret);
                    ch_cud_copy(ws->temp[Lp], Lp, ws->best[L], L, nx[L] - nx[Lp], nyy[L] - nyy[Lp], t);
                    ch_update_map(P, ws->best[L], nx[L], nyy[L], L, t);
                    __syncthreads();
                    npart[Lp]++;
                    L = Lp;
                    st = NEXT_PART;
                }
            }

            // ---- mode_analyze_lcu tail: update_to_ctx_map; the bitstream pass then marks the luma cbf (src_base/xeve_eco.c:1575-1603) ----
            const ChCud &b = ws->best[4];
            ch_update_map(P, b, xc, yc, 4, t);
            ch_state_copy(&S.chain, &S.next[4], t);   // B / I slices: the next CTU starts from the state this decision pass ended with
            if(P.ctu_state) ch_state_copy(&P.ctu_state[2 * lcu + 1], &S.next[4], t);
            __syncthreads();
            {
                const int wsu = (xc + 64 > P.w ? P.w - xc : 64) >> 2, hsu = (yc + 64 > P.h ? P.h - yc : 64) >> 2;
                for(int i = t; i < wsu * hsu; i += CH_T) {
                    const int j = i / wsu, k = i - j * wsu, u = j * 16 + k;
                    const size_t p = (size_t)((yc >> 2) + j) * P.w_scu + (xc >> 2) + k;
                    if(b.r[u].nnz[0] > 0) P.map_scu[p] |= 1u << 24;
                    const int cw = 1 << (b.r[u].log2 - 2);
                    P.df_flags[p] = (uint8_t)((((xc >> 2) + k) > 0 && (k & (cw - 1)) == 0 ? 1 : 0) | (((yc >> 2) + j) > 0 && (j & (cw - 1)) == 0 ? 2 : 0));
                }
                ch_copy32(P.scu_out + (size_t)lcu * 256, b.r, 256 * CH_REC_WORDS, t);
                ch_copy32(P.coef_out + (size_t)lcu * 6144, b.coef, 6144 / 2, t);
            }
            __threadfence();
            __syncthreads();
            if(t == 0) atomicExch(P.done + lcu, 1);
        }
    if(t == 0) { atomicAdd(P.counts, (unsigned long long)n_inter); atomicAdd(P.counts + 1, (unsigned long long)n_intra); }
}

XB_DEV unsigned long long ch_globaltimer()
{
    unsigned long long v;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
    return v;
}

// A chain worker: takes tickets until the host tells it to leave.
template <int MIN_BLOCKS>
__global__ void __launch_bounds__(CH_T, MIN_BLOCKS) k_chain_server(ChainQueue *__restrict__ q, const int8_t *__restrict__ g_tm64)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int8_t   *tm = reinterpret_cast<int8_t *>(smem_raw), *tmT = tm + 4096;
    ChShared &S = *reinterpret_cast<ChShared *>(smem_raw + 8192);
    const int t = threadIdx.x;
    for(int e = t; e < 4096; e += CH_T) {
        const int8_t v = g_tm64[e];
        tm[e] = v;
        tmT[(e & 63) * 64 + (e >> 6)] = v;
    }
    if(t < 64) S.zinv8[zz_of(t, 3)] = (uint8_t)t;
    if(t < 4) { mbar_init(&S.bar[t], 1); S.bphase[t] = 0; }
    if(t == 0) { S.phase = 0; S.bits = 0; S.satd = 0; }
    __syncthreads();
    for(;;) {
        if(t == 0) {
            const unsigned ticket = atomicAdd(&q->head, 1u);
            volatile unsigned *tail = &q->tail;
            volatile int      *stop = &q->stop;
            int slot = -1, chain = 0;
            for(;;) {
                if((int)(*tail - ticket) > 0) {
                    __threadfence();                      // the task records were written before tail moved
                    const volatile int32_t *tk = reinterpret_cast<const volatile int32_t *>(&q->tasks[ticket % CH_Q_TASKS]);
                    slot = tk[0]; chain = tk[1];
                    break;
                }
                if(*stop) break;
                __nanosleep(2000);
            }
            S.task_slot = slot; S.task_chain = chain;
        }
        __syncthreads();
        const int slot = S.task_slot, chain = S.task_chain;
        if(slot < 0) return;
        ChainPicTask *pt = &q->slot[slot];
        {   // the picture's parameters -> shared memory (read through L2: the host wrote them while this grid was running)
            const uint32_t *src = reinterpret_cast<const uint32_t *>(&pt->P);
            uint32_t       *dst = reinterpret_cast<uint32_t *>(&S.P);
            for(int i = t; i < (int)(sizeof(ChainPic) / 4); i += CH_T) dst[i] = __ldcg(src + i);
            const uint32_t *s2 = reinterpret_cast<const uint32_t *>(&pt->sq);
            uint32_t       *d2 = reinterpret_cast<uint32_t *>(&S.sq);
            for(int i = t; i < (int)(sizeof(SeqDev) / 4); i += CH_T) d2[i] = __ldcg(s2 + i);
        }
        if(t == 0) atomicMin(&pt->t_first, ch_globaltimer());
        __syncthreads();
        chain_picture(pt->pics, S.P, S.sq, pt->err_flag, chain, smem_raw);
        __syncthreads();
        if(t == 0) {
            __threadfence();
            const unsigned long long now = ch_globaltimer();
            atomicMax(&pt->t_last, now);
            if(atomicAdd(&pt->finished, 1u) == (unsigned)S.P.n_chain - 1) {
                __threadfence();
                *pt->h_done = pt->seq;                    // host-mapped memory: the scheduler polls it
                __threadfence_system();
            }
        }
        __syncthreads();
    }
}
