// xb200_residue2.cuh -- size-specialised fused candidate evaluation (v2 of k_residue).
//
// Same arithmetic as k_residue (MC -> residual -> SSD -> DCT + RDOQ -> dequant + IDCT -> recon -> SSD,
// the distortion/transform body of pinter_residue_rdo, reference src_base/xeve_pinter.c:961-1056), but
//   * one kernel instantiation per CU size, every block dimension a compile-time constant;
//   * a TEAM of T threads per item: one warp for 8x8 and 16x16 CUs (4 items per CTA, warp-synchronous,
//     no block barriers), 128 threads for 32x32, 256 for 64x64;
//   * persistent CTAs: the DCT matrix is brought into shared memory once per CTA, teams loop over items;
//   * all reductions / scans of the parallel RDOQ are warp-shuffle based.
#pragma once
#include "xb200_common.cuh"
#include "xb200_tq.cuh"
#include "xb200_dct_tc.cuh"

template <int T> XB_DEV void team_sync() { team_bar<T>(); }

struct TeamScratch { // cross-warp exchange, only used by teams wider than a warp
    int64_t w64[8];
    int32_t w32[8];
};

template <int T> XB_DEV int64_t team_sum_s64(int64_t v, int tt, TeamScratch &X)
{
#pragma unroll
    for(int m = 16; m > 0; m >>= 1) v += (int64_t)shfl_xor_u64((uint64_t)v, m);
    if(T > 32) {
        if((tt & 31) == 0) X.w64[tt >> 5] = v;
        team_bar<T>();
        uint64_t s = 0;
#pragma unroll
        for(int i = 0; i < T / 32; i++) s += (uint64_t)X.w64[i];
        team_bar<T>();
        v = (int64_t)s;
    }
    return v;
}
template <int T> XB_DEV int team_sum_s32(int v, int tt, TeamScratch &X)
{
#pragma unroll
    for(int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if(T > 32) {
        if((tt & 31) == 0) X.w32[tt >> 5] = v;
        team_bar<T>();
        int s = 0;
#pragma unroll
        for(int i = 0; i < T / 32; i++) s += X.w32[i];
        team_bar<T>();
        v = s;
    }
    return v;
}

// ---- prediction ---------------------------------------------------------------------------------------------
// one plane, one list; BW x BH compile-time; dst stride BW; tmp >= (BH + TAPS - 1) * BW samples
template <int TAPS, int FR, int BW, int BH, int T>
XB_DEV void mc_plane_t(const int16_t *__restrict__ ref, int sr, int gx, int gy, bool want_h, bool want_v, int16_t *__restrict__ dst,
                       int bd, int16_t *__restrict__ tmp, int tt)
{
    constexpr int HALF = TAPS / 2 - 1, FSH = FR == 4 ? 2 : 3;
    constexpr int LBW = BW == 64 ? 6 : BW == 32 ? 5 : BW == 16 ? 4 : BW == 8 ? 3 : 2;
    const int      maxv = (1 << bd) - 1;
    const int16_t *base = ref + (ptrdiff_t)(gy >> FSH) * sr + (gx >> FSH);
    const int16_t *ch = TAPS == 8 ? c_mc_l[gx & (FR - 1)] : c_mc_c[gx & (FR - 1)];
    const int16_t *cv = TAPS == 8 ? c_mc_l[gy & (FR - 1)] : c_mc_c[gy & (FR - 1)];
    if(!want_h && !want_v) {
#pragma unroll 2
        for(int e = tt; e < BW * BH; e += T) dst[e] = base[(ptrdiff_t)(e >> LBW) * sr + (e & (BW - 1))];
    }
    else if(want_h != want_v) {
        const ptrdiff_t stp = want_h ? 1 : sr;
        int             cf[TAPS];
#pragma unroll
        for(int t = 0; t < TAPS; t++) cf[t] = want_h ? ch[t] : cv[t];
        for(int e = tt; e < BW * BH; e += T) {
            const int16_t *p = base + (ptrdiff_t)(e >> LBW) * sr + (e & (BW - 1)) - HALF * stp;
            int acc = 0;
#pragma unroll
            for(int t = 0; t < TAPS; t++) acc += cf[t] * p[t * stp];
            dst[e] = (int16_t)clip3i(0, maxv, acc >> 6);
        }
    }
    else {
        const int s1 = min(4, bd - 8), s2 = max(8, 20 - bd);
        int       cf[TAPS];
#pragma unroll
        for(int t = 0; t < TAPS; t++) cf[t] = ch[t];
        for(int e = tt; e < BW * (BH + TAPS - 1); e += T) {
            const int16_t *p = base + (ptrdiff_t)((e >> LBW) - HALF) * sr + (e & (BW - 1)) - HALF;
            int acc = 0;
#pragma unroll
            for(int t = 0; t < TAPS; t++) acc += cf[t] * p[t];
            tmp[e] = (int16_t)(acc >> s1);
        }
        team_sync<T>();
#pragma unroll
        for(int t = 0; t < TAPS; t++) cf[t] = cv[t];
        for(int e = tt; e < BW * BH; e += T) {
            int acc = 0;
#pragma unroll
            for(int t = 0; t < TAPS; t++) acc += cf[t] * tmp[e + t * BW];
            dst[e] = (int16_t)clip3i(0, maxv, (acc + (1 << (s2 - 1))) >> s2);
        }
    }
    team_sync<T>();
}

template <int L2, int T>
XB_DEV void mc_item_t(const PicDev *__restrict__ pics, const xb200_mc_item &it, const SeqDev &sq, int16_t *pred, int16_t *aux,
                      int16_t *tmp, int tt)
{
    constexpr int N = 1 << L2, NY = N * N, NCH = NY >> 2;
    const int     x = it.x, y = it.y;
    int           mvt[2][2];
#pragma unroll
    for(int l = 0; l < 2; l++) {
        mvt[l][0] = it.mv[l][0]; mvt[l][1] = it.mv[l][1];
        if(it.refi[l] >= 0) { // xeve_mv_clip, reference src_base/xeve_mc.c:401-447
            const int lo = -(128 << 2), hx = (sq.w - 1 + 128) << 2, hy = (sq.h - 1 + 128) << 2;
            if((x << 2) + it.mv[l][0] < lo) mvt[l][0] = lo - (x << 2);
            if((y << 2) + it.mv[l][1] < lo) mvt[l][1] = lo - (y << 2);
            if((x << 2) + it.mv[l][0] + (N << 2) - 4 > hx) mvt[l][0] = hx - (x << 2) - (N << 2) + 4;
            if((y << 2) + it.mv[l][1] + (N << 2) - 4 > hy) mvt[l][1] = hy - (y << 2) - (N << 2) + 4;
            mvt[l][0] = (int16_t)mvt[l][0]; mvt[l][1] = (int16_t)mvt[l][1];
        }
    }
    int n = 0;
#pragma unroll
    for(int l = 0; l < 2; l++) {
        if(it.refi[l] < 0) continue;
        if(l == 1 && it.refi[0] >= 0 && it.ref_poc[0] == it.ref_poc[1] && mvt[0][0] == mvt[1][0] && mvt[0][1] == mvt[1][1]) break;
        const PicDev &rp  = pics[it.ref_pic[l]];
        int16_t      *dst = n == 0 ? pred : aux;
        const int     gx = (x << 2) + mvt[l][0], gy = (y << 2) + mvt[l][1];
        mc_plane_t<8, 4, N, N, T>(rp.p[0], rp.s[0], gx, gy, (it.mv[l][0] & 3) != 0, (it.mv[l][1] & 3) != 0, dst, sq.bd, tmp, tt);
        const bool hc = (it.mv[l][0] & 7) != 0, vc = (it.mv[l][1] & 7) != 0;
        mc_plane_t<4, 8, N / 2, N / 2, T>(rp.p[1], rp.s[1], gx, gy, hc, vc, dst + NY, sq.bd, tmp, tt);
        mc_plane_t<4, 8, N / 2, N / 2, T>(rp.p[2], rp.s[2], gx, gy, hc, vc, dst + NY + NCH, sq.bd, tmp, tt);
        n++;
    }
    if(n == 2) {
        for(int e = tt; e < NY + 2 * NCH; e += T) pred[e] = (int16_t)((pred[e] + aux[e] + 1) >> 1);
        team_sync<T>();
    }
}

// ---- transforms (LN = log2 of this plane's block size) ------------------------------------------------------------
template <int LN, int T> XB_DEV void fwd_dct_t(int16_t *blk, int32_t *TB, const int8_t *tm, const int8_t *tmT, int bd, int tt)
{
    constexpr int N = 1 << LN, K = N == 64 ? 32 : N, LK = N == 64 ? 5 : LN, ks = 6 - LN;
    const int     shift = (LN - 1 + bd - 8) + (LN + 6);
    for(int e = tt; e < N * K; e += T) { // TB[y][u] = sum_x tm[u][x] * X[y][x]
        const int y = e >> LK, u = e & (K - 1);
        int acc = 0;
#pragma unroll 8
        for(int x = 0; x < N; x++) acc += (int)tmT[x * 64 + (u << ks)] * (int)blk[y * N + x];
        TB[e] = acc;
    }
    team_sync<T>();
    for(int e = tt; e < N * N; e += T) { // C[v][u] = (sum_y tm[v][y] * TB[y][u] + rnd) >> shift
        const int v = e >> LN, u = e & (N - 1);
        int16_t   out = 0;
        if(v < K && u < K) {
            int64_t acc = 0;
#pragma unroll 8
            for(int y = 0; y < N; y++) acc += (int64_t)tm[(v << ks) * 64 + y] * (int64_t)TB[y * K + u];
            out = (int16_t)((acc + ((int64_t)1 << (shift - 1))) >> shift);
        }
        blk[e] = out;
    }
    team_sync<T>();
}

template <int LN, int T> XB_DEV void inv_dct_t(int16_t *blk, int32_t *TB, const int8_t *tm, int bd, int tt)
{
    constexpr int N = 1 << LN, ks = 6 - LN;
    const int     shift = 7 + 12 - (bd - 8);
    for(int e = tt; e < N * N; e += T) { // TB[y][u] = sum_v tm[v][y] * C[v][u]  (fits 32 bits)
        const int y = e >> LN, u = e & (N - 1);
        int acc = 0;
#pragma unroll 8
        for(int v = 0; v < N; v++) acc += (int)tm[(v << ks) * 64 + y] * (int)blk[v * N + u];
        TB[e] = acc;
    }
    team_sync<T>();
    for(int e = tt; e < N * N; e += T) { // X[y][x] = clip16((sum_u tm[u][x] * TB[y][u] + rnd) >> shift)
        const int y = e >> LN, x = e & (N - 1);
        int64_t acc = 0;
#pragma unroll 8
        for(int u = 0; u < N; u++) acc += (int64_t)tm[(u << ks) * 64 + x] * (int64_t)TB[y * N + u];
        acc = (acc + ((int64_t)1 << (shift - 1))) >> shift;
        blk[e] = (int16_t)max((int64_t)-32768, min((int64_t)32767, acc));
    }
    team_sync<T>();
}

// ---- RDOQ (see xb200_tq.cuh for the derivation of the two-state scan) -----------------------------------------------
XB_DEV int compose_map(int first, int second) { return ((second >> (first & 1)) & 1) | (((second >> ((first >> 1) & 1)) & 1) << 1); }

// INTRA: the coded-block flag of an intra luma block is cbf_luma, of an inter one cbf_all (src_base/xeve_tq.c:565-577)
template <int LN, int T, bool INTRA = false>
XB_DEV int quant_team(int16_t *blk, int32_t *TB, int qp, double d_lambda, int ch, int slice_type, const xb200_rates *__restrict__ rt,
                      int bd, int use_rdoq, int tt, TeamScratch &X)
{
    constexpr int N = 1 << LN, n = N * N;
    const int     q = c_quant_scale[qp % 6], qbits = 14 + (15 - bd - LN) + qp / 6;
    if(!use_rdoq) {
        const int32_t off = (int32_t)(slice_type == 2 ? 171 : 85) << (qbits - 9);
        int           cnt = 0;
        for(int e = tt; e < n; e += T) {
            const int     c   = blk[e];
            const int32_t lev = (int16_t)(((int32_t)abs(c) * q + off) >> qbits);
            const int16_t o   = (int16_t)(c < 0 ? -lev : lev);
            blk[e] = o;
            cnt += o != 0;
        }
        team_sync<T>();
        return team_sum_s32<T>(cnt, tt, X);
    }
    // zero-block pre-test + per-coefficient first pass, fused (the sums are only used if the block is coded)
    RdoqEnv E;
    E.lambda = (int64_t)(d_lambda * 32768.0 + 0.5);
    E.es     = c_err_scale[bd - 8][qp % 6][LN];
    E.qbits  = qbits;
    int16_t  *sc  = reinterpret_cast<int16_t *>(TB);
    uint16_t *pos = reinterpret_cast<uint16_t *>(TB) + n;
    const int64_t thr = ((int64_t)1 << qbits) - ((int64_t)(slice_type == 2 ? 201 : 153) << (qbits - 9));
    int64_t unc_part = 0;
    int     flags = 0; // bit 0: passes the zero-block threshold, bit 1: some max level != 0
    for(int e = tt; e < n; e += T) {
        const int x = e & (N - 1), y = e >> LN, d = x + y, c = blk[e];
        const int before = d < N ? (d * (d + 1)) >> 1 : n - (((2 * N - 1 - d) * (2 * N - d)) >> 1);
        const int mx = min(d, N - 1);
        const int sp = before + ((d & 1) ? mx - x : mx - y);
        sc[sp] = (int16_t)c; pos[sp] = (uint16_t)e;
        int64_t  ld; uint32_t maxl;
        rq_quant(c, q, qbits, ld, maxl);
        const int64_t e0 = (ld * E.es) >> 20;
        unc_part += e0 * e0;
        flags |= ((int64_t)abs(c) * q >= thr ? 1 : 0) | (maxl ? 2 : 0);
    }
    // OR-reduce the flags as a sum of two counters
    const int fsum = team_sum_s32<T>((flags & 1) | ((flags & 2) << 15), tt, X);
    if((fsum & 0xffff) == 0 || (fsum >> 16) == 0) {
        team_sync<T>();
        for(int e = tt; e < n; e += T) blk[e] = 0;
        team_sync<T>();
        return 0;
    }
    const int64_t unc_blk = team_sum_s64<T>(unc_part, tt, X);
    {
        const int ctx = ch == 0 ? 0 : 2;
        E.run[0][0] = rt->run[ctx][0]; E.run[0][1] = rt->run[ctx][1];
        E.run[1][0] = rt->run[ctx + 1][0]; E.run[1][1] = rt->run[ctx + 1][1];
        E.lev[0][0] = rt->level[ctx][0]; E.lev[0][1] = rt->level[ctx][1];
        E.lev[1][0] = rt->level[ctx + 1][0]; E.lev[1][1] = rt->level[ctx + 1][1];
    }
    const int32_t *cbf = ch == 0 ? (INTRA ? rt->cbf_luma : rt->cbf_all) : (ch == 1 ? rt->cbf_cb : rt->cbf_cr);
    const int64_t  best0 = unc_blk + (int64_t)cbf[0] * E.lambda, base0 = unc_blk + (int64_t)cbf[1] * E.lambda;
    const int64_t  last0 = (int64_t)rt->last[ch == 0 ? 0 : 1][0] * E.lambda, last1 = (int64_t)rt->last[ch == 0 ? 0 : 1][1] * E.lambda;
    const int64_t  zero_rate[2] = {(int64_t)E.run[0][1] * E.lambda, (int64_t)E.run[1][1] * E.lambda};
    team_sync<T>();
    constexpr int CH = n >= T ? n / T : 1;
    const int     s_beg = min(n, tt * CH), s_end = min(n, s_beg + CH), lane = tt & 31;
    // pass 1: state map of the chunk
    int map;
    {
        int sa = 0, sb = 1;
        for(int sp = s_beg; sp < s_end; sp++) {
            int64_t ld, dl; uint32_t maxl;
            rq_quant(sc[sp], q, qbits, ld, maxl);
            if(maxl == 0) { sa = sb = 1; continue; }
            const int na = rq_level(E, ld, maxl, sa, dl) ? 0 : 1;
            sb = (sb == sa) ? na : (rq_level(E, ld, maxl, sb, dl) ? 0 : 1);
            sa = na;
        }
        map = sa | (sb << 1);
    }
    int incl = map;
#pragma unroll
    for(int d = 1; d < 32; d <<= 1) {
        const int prev = __shfl_up_sync(0xffffffffu, incl, d);
        if(lane >= d) incl = compose_map(prev, incl);
    }
    int excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if(lane == 0) excl = 2; // identity
    int state = 0;
    if(T > 32) {
        if(lane == 31) X.w32[tt >> 5] = incl;
        team_bar<T>();
        for(int w = 0; w < (tt >> 5); w++) state = (X.w32[w] >> state) & 1;
        team_bar<T>();
    }
    state = (excl >> state) & 1;
    // pass 2: levels, increments, local best "last" candidate
    int64_t run_sum = 0, loc_best = 0;
    int     loc_idx = -1;
    for(int sp = s_beg; sp < s_end; sp++) {
        const int c = sc[sp];
        int64_t   ld, dl; uint32_t maxl, lev = 0;
        rq_quant(c, q, qbits, ld, maxl);
        if(maxl == 0) dl = zero_rate[state];
        else lev = rq_level(E, ld, maxl, state, dl);
        sc[sp] = (int16_t)(c > 0 ? (int)lev : -(int)lev);
        run_sum += dl;
        if(lev) {
            const int64_t cand = run_sum + last1;
            if(loc_idx < 0 || cand < loc_best) { loc_best = cand; loc_idx = sp; }
            run_sum += last0;
            state = 0;
        }
        else state = 1;
    }
    // exclusive prefix of the chunk totals
    int64_t inc = run_sum;
#pragma unroll
    for(int d = 1; d < 32; d <<= 1) {
        const int64_t prev = shfl_up_s64(inc, d);
        if(lane >= d) inc += prev;
    }
    int64_t pre = inc - run_sum;
    if(T > 32) {
        if(lane == 31) X.w64[tt >> 5] = inc;
        team_bar<T>();
        for(int w = 0; w < (tt >> 5); w++) pre += X.w64[w];
        team_bar<T>();
    }
    // arg-min of (value, scan index): first minimum in scan order
    int64_t bv = (loc_idx >= 0) ? base0 + pre + loc_best : INT64_MAX;
    int     bidx = loc_idx >= 0 ? loc_idx : 0x7fffffff;
#pragma unroll
    for(int m = 16; m > 0; m >>= 1) {
        const int64_t ov = (int64_t)shfl_xor_u64((uint64_t)bv, m);
        const int     oi = __shfl_xor_sync(0xffffffffu, bidx, m);
        if(ov < bv || (ov == bv && oi < bidx)) { bv = ov; bidx = oi; }
    }
    if(T > 32) {
        if(lane == 0) { X.w64[tt >> 5] = bv; X.w32[tt >> 5] = bidx; }
        team_bar<T>();
        bv = X.w64[0]; bidx = X.w32[0];
#pragma unroll
        for(int w = 1; w < T / 32; w++)
            if(X.w64[w] < bv || (X.w64[w] == bv && X.w32[w] < bidx)) { bv = X.w64[w]; bidx = X.w32[w]; }
        team_bar<T>();
    }
    const int best_last = (bidx != 0x7fffffff && bv < best0) ? bidx + 1 : 0;
    team_sync<T>(); // sc[] complete
    int cnt = 0;
    for(int sp = tt; sp < n; sp += T) {
        const int16_t v = sp < best_last ? sc[sp] : (int16_t)0;
        blk[pos[sp]]    = v;
        cnt += v != 0;
    }
    team_sync<T>();
    return team_sum_s32<T>(cnt, tt, X);
}

template <int LN, int T> XB_DEV void dequant_team(int16_t *blk, int qp, int bd, int tt)
{
    constexpr int n = 1 << (2 * LN);
    const int     shift = 20 - 14 - (15 - bd - LN);
    const int64_t scale = (int64_t)c_dequant_scale[qp % 6] << (qp / 6), off = shift ? (int64_t)1 << (shift - 1) : 0;
    for(int e = tt; e < n; e += T) {
        const int64_t v = ((int64_t)blk[e] * scale + off) >> shift;
        blk[e] = (int16_t)max((int64_t)-32768, min((int64_t)32767, v));
    }
    team_sync<T>();
}

// ---- one plane of one item: residual, SSD, TQ, ITDQ, recon, SSD ------------------------------------------------------------
// TCW / tmh: tensor-core work area and fp16 matrix of this plane's size, or null -> integer transform
template <int LN, int T, int LNMAX, bool USE_TC>
XB_DEV void residue_plane(const int16_t *__restrict__ org, int so, const int16_t *pr, int16_t *blk, int32_t *TB, const int8_t *tm,
                          const int8_t *tmT, int16_t *__restrict__ gco, int16_t *__restrict__ grec, int run, int qp, double lambda,
                          int ch, int slice_type, const xb200_rates *__restrict__ rt, const SeqDev &sq, int tt, TeamScratch &X,
                          int &nnz_out, int64_t &dist_pred, int64_t &dist_rec, TcWork<LNMAX> *TCW, const __half *tmh)
{
    constexpr int N = 1 << LN, nn = N * N;
    const int     maxv = (1 << sq.bd) - 1, sh = (sq.bd - 8) << 1;
    int64_t       dpart = 0;
    for(int e = tt; e < nn; e += T) {
        const int d = (int)org[(ptrdiff_t)(e >> LN) * so + (e & (N - 1))] - (int)pr[e];
        blk[e] = (int16_t)d;
        dpart += (d * d) >> sh;
    }
    team_sync<T>();
    dist_pred = team_sum_s64<T>(dpart, tt, X);
    int nnz = 0;
    if(run) {
        if constexpr(USE_TC && LN >= 5 && T >= 128 && LN <= LNMAX) {
            // residuals of <= 10-bit samples are fp16-exact: 32/64-point stages go to the tensor cores
            if(TCW != nullptr && sq.bd <= 10) tc_fwd_dct<LN, T, LNMAX>(*TCW, tmh, blk, sq.bd, tt);
            else fwd_dct_t<LN, T>(blk, TB, tm, tmT, sq.bd, tt);
        }
        else fwd_dct_t<LN, T>(blk, TB, tm, tmT, sq.bd, tt);
        nnz = quant_team<LN, T>(blk, TB, qp, lambda, ch, slice_type, rt, sq.bd, sq.rdoq, tt, X);
    }
    for(int e = tt; e < nn; e += T) gco[e] = blk[e];
    dist_rec = dist_pred;
    if(nnz) {
        team_sync<T>();
        dequant_team<LN, T>(blk, qp, sq.bd, tt);
        inv_dct_t<LN, T>(blk, TB, tm, sq.bd, tt);
        int64_t rpart = 0;
        for(int e = tt; e < nn; e += T) {
            const int16_t t = (int16_t)(blk[e] + pr[e]);
            const int     v = clip3i(0, maxv, t);
            if(grec) grec[e] = (int16_t)v;
            const int d     = v - (int)org[(ptrdiff_t)(e >> LN) * so + (e & (N - 1))];
            rpart += (d * d) >> sh;
        }
        dist_rec = team_sum_s64<T>(rpart, tt, X);
    }
    else if(grec) {
        for(int e = tt; e < nn; e += T) grec[e] = (int16_t)clip3i(0, maxv, pr[e]);
    }
    team_sync<T>();
    nnz_out = nnz;
}

template <int L2> struct Res2Cfg {
    static constexpr int T     = L2 <= 4 ? 32 : (L2 == 5 ? 128 : XB200_T64);
    static constexpr int CTA   = L2 <= 4 ? 128 : T;
    static constexpr int TEAMS = CTA / T;
    static constexpr int N     = 1 << L2;
    static constexpr int PRED  = N * N * 3 / 2;                               // samples
    static constexpr int TBW   = L2 == 6 ? 4096 : N * N;                      // int32 words: DCT stage / MC tmp / RDOQ scratch
    static constexpr int TEAM_BYTES = (2 * PRED + N * N) * 2 + TBW * 4 + (int)sizeof(TeamScratch);
    static constexpr bool TC   = L2 >= 5;                                     // 32/64-point luma (and 32-point chroma) on tcgen05
    static constexpr int TC_BYTES = TC ? (int)sizeof(TcWork<(L2 >= 5 ? L2 : 5)>) + 128 * N * 2 + (L2 == 6 ? 128 * 32 * 2 : 0) : 0;
    static constexpr int SMEM_INT = 8192 + TEAMS * TEAM_BYTES;                  // integer transform only
    static constexpr int SMEM  = SMEM_INT + TC_BYTES + (TC ? 128 : 0);           // with the tensor-core work area
};

template <int L2, bool USE_TC>
__global__ void __launch_bounds__(Res2Cfg<L2>::CTA) k_residue2(const PicDev *__restrict__ pics, xb200_residue_item *__restrict__ items,
                                                               const int32_t *__restrict__ order, int n,
                                                               const xb200_rates *__restrict__ rates, int16_t *__restrict__ coef,
                                                               int16_t *__restrict__ rec, const int8_t *__restrict__ g_tm64, SeqDev sq,
                                                               int16_t *__restrict__ pred_out)
{
    using Cf = Res2Cfg<L2>;
    constexpr int T = Cf::T, N = Cf::N, NY = N * N, NCH = NY >> 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int8_t *tm = reinterpret_cast<int8_t *>(smem_raw), *tmT = tm + 4096;
    const int team = threadIdx.x / T, tt = threadIdx.x % T;
    unsigned char *tb   = smem_raw + 8192 + team * Cf::TEAM_BYTES;
    int16_t       *pred = reinterpret_cast<int16_t *>(tb);
    int16_t       *aux  = pred + Cf::PRED;
    int16_t       *blk  = aux + Cf::PRED;
    int32_t       *TB   = reinterpret_cast<int32_t *>(blk + NY);
    TeamScratch   &X    = *reinterpret_cast<TeamScratch *>(TB + Cf::TBW);
    for(int e = threadIdx.x; e < 4096; e += Cf::CTA) {
        const int8_t v = g_tm64[e];
        tm[e] = v;
        tmT[(e & 63) * 64 + (e >> 6)] = v;
    }
    constexpr int LNMAX = L2 >= 5 ? L2 : 5;
    TcWork<LNMAX> *TCW = nullptr;
    __half        *tmh_y = nullptr, *tmh_c = nullptr;
    if constexpr(Cf::TC && USE_TC) {
        unsigned char *tcb = smem_raw + 8192 + Cf::TEAMS * Cf::TEAM_BYTES;
        tcb += (128 - (smem_u32(tcb) & 127)) & 127;
        TCW   = reinterpret_cast<TcWork<LNMAX> *>(tcb);
        tmh_y = reinterpret_cast<__half *>(tcb + sizeof(TcWork<LNMAX>));
        tc_fill_tm<L2, Cf::CTA>(tmh_y, g_tm64, threadIdx.x);
        if constexpr(L2 == 6) {
            tmh_c = tmh_y + 128 * N;
            tc_fill_tm<5, Cf::CTA>(tmh_c, g_tm64, threadIdx.x);
        }
        tc_setup<LNMAX, Cf::CTA>(*TCW, threadIdx.x);
    }
    __syncthreads();
    for(int i = blockIdx.x * Cf::TEAMS + team; i < n; i += gridDim.x * Cf::TEAMS) {
        xb200_residue_item *it = &items[order[i]];
        const xb200_mc_item mc = it->mc;
        mc_item_t<L2, T>(pics, mc, sq, pred, aux, reinterpret_cast<int16_t *>(TB), tt);
        const PicDev &o = pics[it->cur_pic];
        const int64_t oo = it->out_off;
        if(pred_out) // the CU pipeline keeps the prediction of every candidate mode (planes dropped by the cbf decision)
            for(int e = tt; e < Cf::PRED; e += T) pred_out[oo + e] = pred[e];
        const xb200_rates *rt = &rates[it->rate_idx];
        const int rs = it->run_stats, st = it->slice_type;
        int     nnz[3];
        int64_t dp[3], dr[3];
        residue_plane<L2, T, LNMAX, USE_TC>(o.p[0] + (ptrdiff_t)mc.y * o.s[0] + mc.x, o.s[0], pred, blk, TB, tm, tmT, coef + oo, rec ? rec + oo : nullptr, rs & 1,
                                    it->qp[0], it->lambda[0], 0, st, rt, sq, tt, X, nnz[0], dp[0], dr[0], TCW, tmh_y);
        residue_plane<L2 - 1, T, LNMAX, USE_TC>(o.p[1] + (ptrdiff_t)(mc.y >> 1) * o.s[1] + (mc.x >> 1), o.s[1], pred + NY, blk, TB, tm, tmT,
                                        coef + oo + NY, rec ? rec + oo + NY : nullptr, (rs >> 1) & 1, it->qp[1], it->lambda[1], 1, st, rt, sq, tt, X,
                                        nnz[1], dp[1], dr[1], tmh_c ? TCW : nullptr, tmh_c);
        residue_plane<L2 - 1, T, LNMAX, USE_TC>(o.p[2] + (ptrdiff_t)(mc.y >> 1) * o.s[2] + (mc.x >> 1), o.s[2], pred + NY + NCH, blk, TB, tm,
                                        tmT, coef + oo + NY + NCH, rec ? rec + oo + NY + NCH : nullptr, (rs >> 2) & 1, it->qp[2], it->lambda[2], 2, st,
                                        rt, sq, tt, X, nnz[2], dp[2], dr[2], tmh_c ? TCW : nullptr, tmh_c);
        if(tt == 0) {
#pragma unroll
            for(int c = 0; c < 3; c++) { it->nnz[c] = nnz[c]; it->dist_pred[c] = dp[c]; it->dist_rec[c] = dr[c]; }
        }
    }
    if constexpr(Cf::TC && USE_TC) tc_teardown<LNMAX>(*TCW, threadIdx.x);
}

#ifndef XB200_DEVICE_FUNCS_ONLY
// CK: the argument checks of a host-buffer call, done here instead of in a host loop over the records (see BinCheck, xb200_api.cu):
// bins[6] invalid argument, bins[7] unsupported shape
template <class CK>
__global__ void k_res_bin(const xb200_residue_item *__restrict__ items, int n, int32_t *__restrict__ order, int *__restrict__ bins, CK ck)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int       key = 5;
    if(i < n) {
        const xb200_residue_item &it = items[i];
        const int w = it.mc.w;
        key = (it.mc.h == w) ? (w == 8 ? 0 : w == 16 ? 1 : w == 32 ? 2 : w == 64 ? 3 : (w == 0 ? 5 : 4)) : 4; // w 0: empty slot
        if(ck.validate && w != 0) {
            auto pic_ok = [&](int h) { return h >= 0 && h < ck.n_pics && ck.pics[h].valid; };
            const int h = it.mc.h;
            bool unsup = w < 4 || h < 4 || w > 64 || h > 64 || (w & 3) || (h & 3) || w != h || w < 8 || (w & (w - 1));
            bool bad = it.mc.refi[0] < 0 && it.mc.refi[1] < 0;
#pragma unroll
            for(int l = 0; l < 2; l++)
                if(it.mc.refi[l] >= 0 && (!pic_ok(it.mc.ref_pic[l]) || ck.pics[it.mc.ref_pic[l]].pad_l == 0)) bad = true;
            // out_off must be even (planes are compacted as 32-bit words); the CU must lie inside the current picture
            bad = bad || !pic_ok(it.cur_pic) || it.rate_idx < 0 || it.rate_idx >= ck.n_rates || it.out_off < 0 || (it.out_off & 1) ||
                  it.out_off + (long long)w * h * 3 / 2 > ck.elems;
            if(!bad) bad = it.mc.x < 0 || it.mc.y < 0 || it.mc.x + w > ck.pics[it.cur_pic].w || it.mc.y + h > ck.pics[it.cur_pic].h;
            if(unsup) { key = 5; atomicOr(&bins[7], 1); }
            else if(bad) { key = 5; atomicOr(&bins[6], 1); }
        }
    }
    const int lane = threadIdx.x & 31;
#pragma unroll
    for(int k = 0; k < 5; k++) {
        const unsigned m = __ballot_sync(0xffffffffu, key == k);
        if(m == 0) continue;
        int base = 0;
        if(lane == __ffs(m) - 1) base = atomicAdd(&bins[k], __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if(key == k && k < 4) order[(size_t)k * n + base + __popc(m & ((1u << lane) - 1))] = i;
    }
}
#endif // XB200_DEVICE_FUNCS_ONLY
