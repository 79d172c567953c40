// xb200_tq.cuh -- forward DCT-II + (RDOQ) quantisation, dequantisation + inverse DCT,
// reconstruction.  Replaces reference src_base/xeve_tq.c:40-404 (tx_pb*b, xeve_trans),
// 425-649 (xeve_rdoq_run_length_cc), 651-730 (xeve_quant_nnz), 750-864 (xeve_sub_block_tq);
// src_base/xeve_itdq.c:34-580 (xeve_itx_pb*b, xeve_dquant, xeve_itdq); src_base/xeve_recon.c:35-57.
//
// Transforms are exact integer matrix products against the DCT matrix held in shared memory in
// both orientations (bank-conflict-free for either stage); the second forward stage and the
// second inverse stage accumulate in 64 bits (IMAD.WIDE) as the reference does.
//
// RDOQ is the reference's sequential zig-zag scan re-expressed as a parallel scan: in the
// Baseline context model the decision at a coefficient depends on the past only through
// "was the previous level zero?" (src_base/xeve_tq.c:492-495, 425-457), i.e. a two-state
// automaton.  Each thread composes the state map of a contiguous chunk of scan positions, the
// block resolves the entry states, a second pass produces levels and cost increments, and an
// exact s64 prefix sum + arg-min (first minimum in scan order) selects the last coefficient.
#pragma once
#include "xb200_common.cuh"

#define TQ_THREADS 128

template <int MAXN> struct TqSmemT {   // MAXN: largest transform size this instance works on
    int8_t   tm[64 * 64];       // tm[k][n]
    int8_t   tmT[64 * 64];      // tmT[n][k]
    int16_t  blk[MAXN * MAXN];  // working block (coefficients / residual), stride = N
    int32_t  T[MAXN * MAXN];    // stage buffer; RDOQ scratch aliases it
    int64_t  red64[TQ_THREADS];
    int32_t  red32[TQ_THREADS];
    int64_t  bcast64[4];
    int32_t  bcast32[8];
};
using TqSmem = TqSmemT<64>;

template <class SM> XB_DEV void tq_load_tm(SM &S, const int8_t *__restrict__ g_tm64, int tid, int nthr)
{
    for(int e = tid; e < 4096; e += nthr) {
        const int8_t v = g_tm64[e];
        S.tm[e] = v;
        S.tmT[(e & 63) * 64 + (e >> 6)] = v;
    }
    __syncthreads();
}

// forward transform of S.blk (N x N, N = 1 << l2), in place
template <class SM> XB_DEV void fwd_dct(SM &S, int l2, int bd, int tid, int nthr)
{
    const int N = 1 << l2, K = N == 64 ? 32 : N, ks = 6 - l2;
    const int shift = (l2 - 1 + bd - 8) + (l2 + 6);
    // stage 0: T[y][u] = sum_x tm[u][x] * X[y][x], exact in 32 bits
    for(int e = tid; e < N * K; e += nthr) {
        const int y = e / K, u = e % K;
        int acc = 0;
        for(int x = 0; x < N; x++) acc += (int)S.tmT[x * 64 + (u << ks)] * (int)S.blk[y * N + x];
        S.T[y * K + u] = acc;
    }
    __syncthreads();
    // stage 1: C[v][u] = (sum_y tm[v][y] * T[y][u] + rnd) >> shift, 64-bit accumulate
    for(int e = tid; e < N * N; e += nthr) {
        const int v = e / N, u = e % N;
        int16_t   out = 0;
        if(v < K && u < K) {
            int64_t acc = 0;
            for(int y = 0; y < N; y++) acc += (int64_t)S.tm[(v << ks) * 64 + y] * (int64_t)S.T[y * K + u];
            out = (int16_t)((acc + ((int64_t)1 << (shift - 1))) >> shift);
        }
        S.blk[e] = out;
    }
    __syncthreads();
}

// inverse transform of S.blk in place (input: dequantised coefficients)
template <class SM> XB_DEV void inv_dct(SM &S, int l2, int bd, int tid, int nthr)
{
    const int N = 1 << l2, ks = 6 - l2, shift = 7 + 12 - (bd - 8);
    // stage 0: T[y][u] = clip32(sum_v tm[v][y] * C[v][u])  (|sum| <= 64*32768*90 fits 32 bits)
    for(int e = tid; e < N * N; e += nthr) {
        const int y = e / N, u = e % N;
        int64_t acc = 0;
        for(int v = 0; v < N; v++) acc += (int64_t)((int)S.tm[(v << ks) * 64 + y] * (int)S.blk[v * N + u]);
        S.T[e] = (int32_t)max((int64_t)INT32_MIN, min((int64_t)INT32_MAX, acc));
    }
    __syncthreads();
    // stage 1: X[y][x] = clip16((sum_u tm[u][x] * T[y][u] + rnd) >> shift)
    for(int e = tid; e < N * N; e += nthr) {
        const int y = e / N, x = e % N;
        int64_t acc = 0;
        for(int u = 0; u < N; u++) acc += (int64_t)S.tm[(u << ks) * 64 + x] * (int64_t)S.T[y * N + u];
        acc = (acc + ((int64_t)1 << (shift - 1))) >> shift;
        S.blk[e] = (int16_t)max((int64_t)-32768, min((int64_t)32767, acc));
    }
    __syncthreads();
}

// ---- block reductions (all threads call) ----------------------------------------------------------
template <class SM> XB_DEV int64_t block_sum_s64(SM &S, int64_t v, int tid, int nthr)
{
    S.red64[tid] = v;
    __syncthreads();
    if(tid == 0) {
        uint64_t s = 0;
        for(int i = 0; i < nthr; i++) s += (uint64_t)S.red64[i];
        S.bcast64[0] = (int64_t)s;
    }
    __syncthreads();
    const int64_t r = S.bcast64[0];
    __syncthreads();
    return r;
}
template <class SM> XB_DEV int block_sum_s32(SM &S, int v, int tid, int nthr)
{
#pragma unroll
    for(int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
    if((tid & 31) == 0) S.red32[tid >> 5] = v;
    __syncthreads();
    int s = 0;
    for(int i = 0; i < nthr / 32; i++) s += S.red32[i];
    __syncthreads();
    return s;
}

struct RdoqEnv {
    int64_t lambda, es;
    int32_t run[2][2];   // [state: run==0 ? 0 : 1][bin]
    int32_t lev[2][2];   // level[ctx][.], level[ctx+1][.]
    int     qbits;
};

XB_DEV int64_t rq_rate(const RdoqEnv &E, uint32_t a, int s)
{
    int32_t rate;
    if(a == 0) rate = E.run[s][1];
    else {
        rate = 32768 + E.run[s][0];
        if(a == 1) rate += E.lev[0][0];
        else rate += E.lev[0][1] + E.lev[1][1] * (int32_t)(a - 2) + E.lev[1][0];
    }
    return (int64_t)rate * E.lambda;
}
// get_coded_level_rl for run-state s; returns level, *delta = coded - uncoded
XB_DEV uint32_t rq_level(const RdoqEnv &E, int64_t ld, uint32_t maxl, int s, int64_t &delta)
{
    const int64_t e0 = (ld * E.es) >> 20, unc = e0 * e0;
    int64_t       cod = unc + rq_rate(E, 0, s);
    uint32_t      lev = 0;
    if(maxl) {
        const uint32_t mn = maxl > 1 ? maxl - 1 : 1;
        for(uint32_t a = maxl; a >= mn; a--) {
            const int64_t d = ld - ((int64_t)a << E.qbits), er = (d * E.es) >> 20;
            const int64_t c = er * er + rq_rate(E, a, s);
            if(c < cod) { lev = a; cod = c; }
        }
    }
    delta = cod - unc;
    return lev;
}
XB_DEV void rq_quant(int c, int q, int qbits, int64_t &ld, uint32_t &maxl)
{
    const int64_t t = (int64_t)abs(c) * q, cap = (int64_t)INT32_MAX - ((int64_t)1 << (qbits - 1));
    ld   = (int32_t)min(t, cap);
    maxl = (uint32_t)(ld >> qbits);
    if(ld - ((int64_t)maxl << qbits) >= ((int64_t)1 << (qbits - 1))) maxl++;
}

// quantise S.blk (N x N transform coefficients) in place; returns nnz (uniform)
template <class SM> XB_DEV int quant_block(SM &S, int l2, int qp, double d_lambda, int is_intra, int ch, int slice_type,
                       const xb200_rates *__restrict__ rt, int bd, int use_rdoq, int tid, int nthr)
{
    const int N = 1 << l2, n = N * N;
    const int q = c_quant_scale[qp % 6], qbits = 14 + (15 - bd - l2) + qp / 6;
    if(!use_rdoq) { // plain quantiser, src_base/xeve_tq.c:704-727
        const int32_t off = (int32_t)(slice_type == 2 ? 171 : 85) << (qbits - 9);
        int           cnt = 0;
        for(int e = tid; e < n; e += nthr) {
            const int     c   = S.blk[e];
            const int32_t lev = (int16_t)(((int32_t)abs(c) * q + off) >> qbits);
            const int16_t o   = (int16_t)(c < 0 ? -lev : lev);
            S.blk[e] = o;
            cnt += o != 0;
        }
        return block_sum_s32(S, cnt, tid, nthr);
    }
    // ---- zero-block pre-test (src_base/xeve_tq.c:666-700) --------------------------------------------
    {
        const int64_t off = (int64_t)(slice_type == 2 ? 201 : 153) << (qbits - 9), thr = ((int64_t)1 << qbits) - off;
        int           coded = 0;
        for(int e = tid; e < n; e += nthr) coded |= ((int64_t)abs((int)S.blk[e]) * q >= thr);
        if(block_sum_s32(S, coded, tid, nthr) == 0) {
            for(int e = tid; e < n; e += nthr) S.blk[e] = 0;
            __syncthreads();
            return 0;
        }
    }
    // ---- scatter into scan order; block uncoded cost ----------------------------------------------------
    int16_t  *sc  = reinterpret_cast<int16_t *>(S.T);        // coefficient, later signed level, by scan pos
    uint16_t *pos = reinterpret_cast<uint16_t *>(S.T) + n;   // raster position by scan pos
    RdoqEnv   E;
    E.lambda = (int64_t)(d_lambda * 32768.0 + 0.5);
    E.es     = c_err_scale[bd - 8][qp % 6][l2];
    E.qbits  = qbits;
    {
        const int ctx = ch == 0 ? 0 : 2;
        E.run[0][0] = rt->run[ctx][0]; E.run[0][1] = rt->run[ctx][1];
        E.run[1][0] = rt->run[ctx + 1][0]; E.run[1][1] = rt->run[ctx + 1][1];
        E.lev[0][0] = rt->level[ctx][0]; E.lev[0][1] = rt->level[ctx][1];
        E.lev[1][0] = rt->level[ctx + 1][0]; E.lev[1][1] = rt->level[ctx + 1][1];
    }
    int64_t unc_part = 0;
    int     any = 0;
    for(int e = tid; e < n; e += nthr) {
        const int x = e & (N - 1), y = e >> l2, d = x + y, c = S.blk[e];
        const int before = d < N ? (d * (d + 1)) >> 1 : n - (((2 * N - 1 - d) * (2 * N - d)) >> 1);
        const int mx = min(d, N - 1);
        const int sp = before + ((d & 1) ? mx - x : mx - y);
        sc[sp] = (int16_t)c; pos[sp] = (uint16_t)e;
        int64_t  ld; uint32_t maxl;
        rq_quant(c, q, qbits, ld, maxl);
        const int64_t e0 = (ld * E.es) >> 20;
        unc_part += e0 * e0;
        any |= maxl != 0;
    }
    const int64_t unc_blk = block_sum_s64(S, unc_part, tid, nthr);
    if(block_sum_s32(S, any, tid, nthr) == 0) { // sum_all == 0
        for(int e = tid; e < n; e += nthr) S.blk[e] = 0;
        __syncthreads();
        return 0;
    }
    const int32_t *cbf = (!is_intra && ch == 0) ? rt->cbf_all : (ch == 0 ? rt->cbf_luma : (ch == 1 ? rt->cbf_cb : rt->cbf_cr));
    const int64_t  best0 = unc_blk + (int64_t)cbf[0] * E.lambda, base0 = unc_blk + (int64_t)cbf[1] * E.lambda;
    const int64_t  last0 = (int64_t)rt->last[ch == 0 ? 0 : 1][0] * E.lambda, last1 = (int64_t)rt->last[ch == 0 ? 0 : 1][1] * E.lambda;

    // ---- pass 1: state map of this thread's chunk (state 0: previous level non-zero / start) -------------
    const int CH = (n + nthr - 1) / nthr, s_beg = min(n, tid * CH), s_end = min(n, s_beg + CH);
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
        S.red32[tid] = sa | (sb << 1);
    }
    __syncthreads();
    int state = 0;
    for(int t = 0; t < tid; t++) state = (S.red32[t] >> state) & 1;
    __syncthreads();
    // ---- pass 2: levels, cost increments, chunk-local best "last" candidate --------------------------
    int64_t run_sum = 0, loc_best = 0;
    int     loc_idx = -1;
    for(int sp = s_beg; sp < s_end; sp++) {
        const int c = sc[sp];
        int64_t   ld, dl; uint32_t maxl;
        rq_quant(c, q, qbits, ld, maxl);
        const uint32_t lev = rq_level(E, ld, maxl, state, dl);
        // sign rule: tmp_coef = coef > 0 ? max : -max; out = tmp_coef < 0 ? -level : level
        sc[sp] = (int16_t)(((c > 0 ? (int)maxl : -(int)maxl) < 0) ? -(int)lev : (int)lev);
        run_sum += dl;
        if(lev) {
            const int64_t cand = run_sum + last1;
            if(loc_idx < 0 || cand < loc_best) { loc_best = cand; loc_idx = sp; }
            run_sum += last0;
            state = 0;
        }
        else state = 1;
    }
    S.red64[tid] = run_sum;
    __syncthreads();
    int64_t pre = base0;
    for(int t = 0; t < tid; t++) pre += S.red64[t];
    __syncthreads();
    // block arg-min over (value, scan index); first minimum wins
    S.red64[tid] = pre + loc_best;
    S.red32[tid] = loc_idx;
    __syncthreads();
    if(tid == 0) {
        int64_t bv = best0; int bi_ = -1;
        for(int t = 0; t < nthr; t++)
            if(S.red32[t] >= 0 && S.red64[t] < bv) { bv = S.red64[t]; bi_ = S.red32[t]; }
        S.bcast32[0] = bi_ + 1;
    }
    __syncthreads();
    const int best_last = S.bcast32[0];
    // ---- write back ----------------------------------------------------------------------------------------
    int cnt = 0;
    for(int sp = tid; sp < n; sp += nthr) {
        const int16_t v = sp < best_last ? sc[sp] : (int16_t)0;
        S.blk[pos[sp]]  = v;
        cnt += v != 0;
    }
    __syncthreads();
    return block_sum_s32(S, cnt, tid, nthr);
}

// dequantise S.blk in place (src_base/xeve_itdq.c:442-452, shift/offset from 454-497)
template <class SM> XB_DEV void dequant_block(SM &S, int l2, int qp, int bd, int tid, int nthr)
{
    const int     n = 1 << (2 * l2), shift = 20 - 14 - (15 - bd - l2);
    const int64_t scale = (int64_t)c_dequant_scale[qp % 6] << (qp / 6), off = shift ? (int64_t)1 << (shift - 1) : 0;
    for(int e = tid; e < n; e += nthr) {
        const int64_t v = ((int64_t)S.blk[e] * scale + off) >> shift;
        S.blk[e] = (int16_t)max((int64_t)-32768, min((int64_t)32767, v));
    }
    __syncthreads();
}

#ifndef XB200_DEVICE_FUNCS_ONLY
// ---- batched kernels --------------------------------------------------------------------------------------
// ctx->fn_tq: one CTA per item, planes processed in turn, in place in the global coefficient buffer
__global__ void __launch_bounds__(TQ_THREADS) k_tq(xb200_tq_item *__restrict__ items, int n, const xb200_rates *__restrict__ rates,
                                                    int16_t *__restrict__ coef, const int8_t *__restrict__ g_tm64, SeqDev sq)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TqSmem &S = *reinterpret_cast<TqSmem *>(smem_raw);
    const int tid = threadIdx.x;
    if(blockIdx.x >= n) return;
    xb200_tq_item *it = &items[blockIdx.x];
    tq_load_tm(S, g_tm64, tid, TQ_THREADS);
    const int l2y = it->log2_cuw, ny = 1 << (2 * l2y), nc = ny >> 2;
    int16_t  *base = coef + it->in_off;
    for(int c = 0; c < 3; c++) {
        int nnz = 0;
        if((it->run_stats >> c) & 1) {
            const int l2 = c ? l2y - 1 : l2y, nn = 1 << (2 * l2);
            int16_t  *g  = base + (c == 0 ? 0 : (c == 1 ? ny : ny + nc));
            for(int e = tid; e < nn; e += TQ_THREADS) S.blk[e] = g[e];
            __syncthreads();
            fwd_dct(S, l2, sq.bd, tid, TQ_THREADS);
            nnz = quant_block(S, l2, it->qp[c], it->lambda[c], it->is_intra, c, it->slice_type, &rates[it->rate_idx], sq.bd,
                              sq.rdoq, tid, TQ_THREADS);
            for(int e = tid; e < nn; e += TQ_THREADS) g[e] = S.blk[e];
            __syncthreads();
        }
        if(tid == 0) it->nnz[c] = nnz;
    }
}

// ctx->fn_itdp
__global__ void __launch_bounds__(TQ_THREADS) k_itdq(const xb200_tq_item *__restrict__ items, int n, int16_t *__restrict__ coef,
                                                      const int8_t *__restrict__ g_tm64, SeqDev sq)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TqSmem &S = *reinterpret_cast<TqSmem *>(smem_raw);
    const int tid = threadIdx.x;
    if(blockIdx.x >= n) return;
    const xb200_tq_item *it = &items[blockIdx.x];
    tq_load_tm(S, g_tm64, tid, TQ_THREADS);
    const int l2y = it->log2_cuw, ny = 1 << (2 * l2y), nc = ny >> 2;
    int16_t  *base = coef + it->in_off;
    for(int c = 0; c < 3; c++) {
        if(!it->nnz[c]) continue;
        const int l2 = c ? l2y - 1 : l2y, nn = 1 << (2 * l2);
        int16_t  *g  = base + (c == 0 ? 0 : (c == 1 ? ny : ny + nc));
        for(int e = tid; e < nn; e += TQ_THREADS) S.blk[e] = g[e];
        __syncthreads();
        dequant_block(S, l2, it->qp[c], sq.bd, tid, TQ_THREADS);
        inv_dct(S, l2, sq.bd, tid, TQ_THREADS);
        for(int e = tid; e < nn; e += TQ_THREADS) g[e] = S.blk[e];
        __syncthreads();
    }
}

// ctx->fn_recon over whole items: rec = clip(pred + resi) where nnz, else clip(pred)
__global__ void k_recon(const xb200_tq_item *__restrict__ items, int n, const int16_t *__restrict__ resi,
                        const int16_t *__restrict__ pred, int16_t *__restrict__ rec, SeqDev sq)
{
    if(blockIdx.x >= n) return;
    const xb200_tq_item *it = &items[blockIdx.x];
    const int ny = 1 << (it->log2_cuw + it->log2_cuh), nc = ny >> 2, maxv = (1 << sq.bd) - 1;
    const int64_t o = it->in_off;
    for(int e = threadIdx.x; e < ny + 2 * nc; e += blockDim.x) {
        const int     c = e < ny ? 0 : (e < ny + nc ? 1 : 2);
        const int16_t t = it->nnz[c] ? (int16_t)(resi[o + e] + pred[o + e]) : pred[o + e];
        rec[o + e]      = (int16_t)clip3i(0, maxv, t);
    }
}
#endif // XB200_DEVICE_FUNCS_ONLY
