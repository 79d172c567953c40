// xb200_tables.h -- codec constants of the EVC Baseline inter/transform path, generated from
// closed forms instead of stored tables (each generator is checked against the reference's
// table in tests/test_tables.py through the oracle harness).
//
//   DCT-II integer matrix  : reference src_base/xeve_tbl.c:83-236 (xeve_tbl_tm2..64)
//   MV bit-length table    : reference src_base/xeve_tbl.c:286-496 (xeve_tbl_mv_bits)
//   ref-index bit table    : reference src_base/xeve_tbl.c:498-517 (xeve_tbl_refi_bits)
//   zig-zag scan           : reference src_base/xeve_tbl.c:625-     (xeve_tbl_scan)
//   quant / dequant scales : reference src_base/xeve_tq.c:37-39, xeve_tbl.c:237
//   MC filter taps         : reference src_base/xeve_mc.c:39-93
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define XB_HD __host__ __device__ __forceinline__
#else
#define XB_HD static inline
#endif

// tm64[k][n] = round_half_away(64*sqrt(2)*cos(pi*(2n+1)*k/128)), row 0 = 64.
// The N-point matrix is tm64[k * 64/N][n], n < N.
static inline void xb200_gen_tm64(int8_t *tm /* [64*64] */)
{
    for(int k = 0; k < 64; k++)
        for(int n = 0; n < 64; n++) {
            double v = k == 0 ? 64.0 : 64.0 * sqrt(2.0) * cos(M_PI * (2 * n + 1) * k / 128.0);
            tm[k * 64 + n] = (int8_t)(v >= 0 ? floor(v + 0.5) : -floor(-v + 0.5));
        }
}

// zig-zag scan of a (1<<log2w) x (1<<log2h) block: anti-diagonals, first one running
// up-right -> down-left alternating, starting (0,0),(1,0),(0,1),(0,2),(1,1),(2,0)...
static inline void xb200_gen_scan(uint16_t *scan, int log2w, int log2h)
{
    const int w = 1 << log2w, h = 1 << log2h;
    int       pos = 0;
    for(int d = 0; d < w + h - 1; d++) {
        if(d & 1) { // odd diagonal: from top-right going down-left
            int x = d < w ? d : w - 1, y = d - x;
            while(x >= 0 && y < h) { scan[pos++] = (uint16_t)(y * w + x); x--; y++; }
        }
        else { // even diagonal: from bottom-left going up-right
            int y = d < h ? d : h - 1, x = d - y;
            while(y >= 0 && x < w) { scan[pos++] = (uint16_t)(y * w + x); x++; y--; }
        }
    }
}

// bits to code one MVD component; valid for -2048 < v <= 2048 (table range), exp-golomb beyond.
// The reference table is off-centre by one entry: its first element (v = -2047) holds 22.
XB_HD int xb200_mvd_bits(int v)
{
    if(v > 2048 || v <= -2048) {
        // src_base/xeve_pinter.c:74-92 (escape branch)
        unsigned a  = (unsigned)(v < 0 ? -v : v);
        int      nn = (int)((a + 1) >> 12), len = 11;
        while(len < 16 && nn != 0) { nn >>= 1; len++; }
        return (len << 1) + 2;
    }
    if(v == -2047) return 22;
    unsigned a = (unsigned)(v < 0 ? -v : v) + 1u;
#if defined(__CUDA_ARCH__)
    const int l = 31 - __clz(a); // floor(log2(|v|+1))
#else
    int l = 0;
    while(a >> (l + 1)) l++;
#endif
    return 2 * l + 1 + (v != 0);
}

XB_HD int xb200_refi_bits(int num_refp, int refi)
{
    if(num_refp < 2) return 0;
    return refi < num_refp - 1 ? refi + 1 : num_refp - 1;
}

// src_base/xeve_pinter.c:94-120
XB_HD int xb200_mv_bits(int mvd_x, int mvd_y, int num_refp, int refi)
{
    return xb200_mvd_bits(mvd_x) + xb200_mvd_bits(mvd_y) + xb200_refi_bits(num_refp, refi);
}

// luma 8-tap (quarter-pel phases 1..3; phase 0 = copy), chroma 4-tap (eighth-pel phases 1..7)
#define XB200_MC_L_TAPS \
    { {0, 0, 0, 64, 0, 0, 0, 0}, {0, 1, -5, 52, 20, -5, 1, 0}, {0, 2, -10, 40, 40, -10, 2, 0}, {0, 1, -5, 20, 52, -5, 1, 0} }
#define XB200_MC_C_TAPS                                                                                                 \
    { {0, 64, 0, 0}, {-2, 58, 10, -2}, {-4, 52, 20, -4}, {-6, 46, 30, -6}, {-8, 40, 40, -8}, {-6, 30, 46, -6},          \
      {-4, 20, 52, -4}, {-2, 10, 58, -2} }

#define XB200_QUANT_SCALE   { 26214, 23302, 20560, 18396, 16384, 14764 }
#define XB200_DEQUANT_SCALE { 40, 45, 51, 57, 64, 71 }

// RDOQ error scale, src_base/xeve_tq.c:406-423 (doubles, evaluated on the host in this order)
static inline int64_t xb200_err_scale(int qp_rem, int log2_size, int bit_depth)
{
    static const int q[6] = XB200_QUANT_SCALE;
    int    tr_shift = 15 - bit_depth - log2_size;
    double e        = (double)(1 << 15) * pow(2.0, -tr_shift);
    e               = e / q[qp_rem] / (1 << (bit_depth - 8));
    return (int64_t)(e * (double)(1 << 20));
}

// Deblocking filter strength (EVC Baseline table, src_base/xeve_tbl.c:239-257): the intra row is a step function of qp
// with steps at the qps below; the "luma cbf" and "motion differs" rows are the same curve lowered by 1 and 2 (floored
// at 0), the fourth row is all zero.  qp outside 0..51 never occurs for the values the encoder stores (clamped here).
XB_HD int xb200_df_strength(int idx, int qp)
{
    if(idx > 2) return 0;
    qp = qp < 0 ? 0 : (qp > 51 ? 51 : qp);
    //                        st:  1   2   3   4   5   6   7   8   9  10  11  12
    const unsigned char step[12] = {18, 27, 32, 35, 38, 40, 42, 43, 44, 45, 46, 47};
    int st = 0;
    for(int i = 0; i < 12; i++) st += qp >= step[i];
    st -= idx;
    return st < 0 ? 0 : st;
}

// Most-probable-mode symbol table of the Baseline intra mode syntax, [left mode + 1][upper mode + 1][mode] with index 0 = neighbour
// not intra / not available (EVC Baseline table; src_base/xeve_tbl.c:40-48)
#define XB200_MPM_TABLE { \
    { {0,2,3,1,4},{0,2,1,3,4},{0,2,1,3,4},{1,2,0,3,4},{0,2,1,3,4},{0,1,2,3,4} }, \
    { {1,0,2,3,4},{0,1,2,3,4},{0,1,2,3,4},{1,2,0,3,4},{0,1,3,2,4},{0,2,1,4,3} }, \
    { {1,0,2,3,4},{1,0,2,3,4},{1,0,2,3,4},{2,0,1,3,4},{1,0,3,2,4},{0,1,2,4,3} }, \
    { {1,0,2,3,4},{0,2,1,3,4},{1,0,2,3,4},{1,2,0,3,4},{0,1,2,3,4},{0,2,1,4,3} }, \
    { {0,1,2,3,4},{0,3,2,1,4},{1,0,2,3,4},{1,2,0,3,4},{1,2,3,0,4},{0,2,1,4,3} }, \
    { {0,1,2,3,4},{0,1,2,4,3},{0,1,2,4,3},{0,2,1,4,3},{0,1,2,3,4},{0,1,2,4,3} } }
