// xb200_deblock.cuh -- in-loop deblocking of a reconstructed picture.
//
// Reference: xeve_loop_filter (src_base/xeve_enc.c:2355-2414) runs two whole-picture passes -- vertical edges, then
// horizontal edges -- each a z-scan walk over the leaf CUs (xeve_deblock_tree, src_base/xeve_df.c:575-639) that filters
// the CU's left (resp. top) edge in 4-sample segments (xeve_deblock_cu_ver / _hor, :253-471) with a strength class
// derived from the two SCUs on either side (get_tbl_qp_to_st, :34-87).
//
// Here one pass is one grid, one thread per 4x4 SCU:
//   * k_df_mark turns the CU list into per-SCU edge flags (bit 0: the SCU's left side is a CU edge, bit 1: its top side);
//   * k_df_pass<HOR> filters the flagged luma segment of its SCU.  Luma segments never interact inside a pass: edges are
//     >= 4 samples apart and a filter reads/writes 2 samples on either side.
//   * Chroma (4:2:0) edges of 4-wide CUs are only 2 samples apart, and the filter READS 2 samples on either side while
//     writing 1: the edge at c reads sample c-2, which the edge at c-2 has just written.  The reference's order along such
//     a run is left-to-right (top-to-bottom) because the z-scan visits the left (upper) CU first.  The thread of the first
//     SCU of a run of consecutively flagged SCUs therefore walks the run sequentially; every other thread skips chroma.
//     Runs only exist where 4x4 CUs touch (intra), so almost all runs have length 1.
// Memory behaviour: each pass reads the flagged neighbourhood once and writes it once, with the 32 lanes of a warp on 32
// consecutive SCUs of one row (coalesced 8-byte segments).
#pragma once
#include "xb200_common.cuh"

#define DF_MCU_IF(m)   (((m) >> 15) & 1u) // src_base/xeve_def.h:591
#define DF_MCU_QP(m)   (int)(((m) >> 16) & 0x7Fu)
#define DF_MCU_CBFL(m) (((m) >> 24) & 1u)
#define DF_MCU_IBC(m)  (((m) >> 26) & 1u)

__global__ void k_df_mark(const xb200_df_cu *__restrict__ cus, int64_t n, int w_scu, int h_scu, uint8_t *__restrict__ flags)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const xb200_df_cu cu = cus[i];
    const int xs = cu.x >> 2, ys = cu.y >> 2;
    if(cu.x < 0 || cu.y < 0 || cu.log2_cuw < 2 || cu.log2_cuh < 2 || cu.log2_cuw > 7 || cu.log2_cuh > 7) return;
    const int cw = min((1 << cu.log2_cuw) >> 2, w_scu - xs), ch = min((1 << cu.log2_cuh) >> 2, h_scu - ys);
    if(cw <= 0 || ch <= 0) return;
    const uint8_t left = xs > 0 ? 1 : 0, top = ys > 0 ? 2 : 0;
    uint8_t *f = flags + (size_t)ys * w_scu + xs;
    f[0] = left | top;
    for(int k = 1; k < cw; k++) f[k] = top;
    for(int k = 1; k < ch; k++) f[(size_t)k * w_scu] = left;
}

// strength class of the edge between SCU p (the one whose QP is used) and SCU q, src_base/xeve_df.c:34-87
XB_DEV int df_class(const uint32_t *__restrict__ scu, const int8_t *__restrict__ refi, const int16_t *__restrict__ mv, int p, int q)
{
    const uint32_t mp = scu[p], mq = scu[q];
    if(DF_MCU_IF(mp) | DF_MCU_IF(mq)) return 0;
    if(DF_MCU_CBFL(mp) | DF_MCU_CBFL(mq)) return 1;
    if(DF_MCU_IBC(mp) | DF_MCU_IBC(mq)) return 2;
    const char2  rp = reinterpret_cast<const char2 *>(refi)[p], rq = reinterpret_cast<const char2 *>(refi)[q];
    const short4 vp = reinterpret_cast<const short4 *>(mv)[p], vq = reinterpret_cast<const short4 *>(mv)[q];
    // invalid lists count as zero motion
    const int p0x = rp.x >= 0 ? vp.x : 0, p0y = rp.x >= 0 ? vp.y : 0, p1x = rp.y >= 0 ? vp.z : 0, p1y = rp.y >= 0 ? vp.w : 0;
    int       q0x = rq.x >= 0 ? vq.x : 0, q0y = rq.x >= 0 ? vq.y : 0, q1x = rq.y >= 0 ? vq.z : 0, q1y = rq.y >= 0 ? vq.w : 0;
    if(rp.x == rq.x && rp.y == rq.y) {}
    else if(rp.x == rq.y && rp.y == rq.x) {
        int t;
        t = q0x; q0x = q1x; q1x = t;
        t = q0y; q0y = q1y; q1y = t;
    }
    else return 2;
    return (abs(p0x - q0x) >= 4 || abs(p0y - q0y) >= 4 || abs(p1x - q1x) >= 4 || abs(p1y - q1y) >= 4) ? 2 : 3;
}

// the 4-tap edge filter (src_base/xeve_df.c:89-251): all values stay far inside s16 for <= 12-bit samples, so plain
// int arithmetic with C's truncating division reproduces the reference's s16 intermediates
template <bool LUMA> XB_DEV void df_taps(int &A, int &B, int &C, int &D, int st, int maxv)
{
    const int d    = (A - (B << 2) + (C << 2) - D) / 8;
    const int ad   = abs(d);
    const int t16  = max(0, (ad - st) << 1);
    int       clip = max(0, ad - t16);
    const int d1   = d < 0 ? -clip : clip;
    if(LUMA) {
        clip >>= 1;
        const int d2 = clip3i(-clip, clip, (A - D) / 4);
        A = clip3i(0, maxv, A - d2);
        D = clip3i(0, maxv, D + d2);
    }
    B = clip3i(0, maxv, B + d1);
    C = clip3i(0, maxv, C - d1);
}

XB_DEV int lo16(uint32_t v) { return (int)(int16_t)(v & 0xffffu); }
XB_DEV int hi16(uint32_t v) { return (int)(int16_t)(v >> 16); }
XB_DEV uint32_t pack16(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }

struct DfArgs {
    int16_t        *pl[3];
    int             s[3];
    int             w_scu, h_scu, bd;
    const uint32_t *scu;
    const int8_t   *refi;
    const int16_t  *mv;
    const uint8_t  *flags;
    xb200_df_pic    pp;
};

template <bool HOR> XB_DEV void df_chroma_segment(const DfArgs &a, int xs, int ys, int cls, int qp)
{
    const int maxv = (1 << a.bd) - 1, bdo = 6 * (a.bd - 8);
#pragma unroll
    for(int c = 1; c < 3; c++) {
        const int qc = clip3i(-bdo, 57, qp + (c == 1 ? a.pp.qp_u_offset : a.pp.qp_v_offset));
        const int st = xb200_df_strength(cls, a.pp.chroma_qp[c - 1][qc + bdo]) << (a.bd - 8);
        if(!st) continue;
        int16_t *b = a.pl[c] + (ptrdiff_t)(ys * 2) * a.s[c] + xs * 2;
        if(HOR) { // rows -2 -1 | 0 1, two columns
            const ptrdiff_t s = a.s[c];
            uint32_t r0 = *reinterpret_cast<uint32_t *>(b - 2 * s), r1 = *reinterpret_cast<uint32_t *>(b - s);
            uint32_t r2 = *reinterpret_cast<uint32_t *>(b), r3 = *reinterpret_cast<uint32_t *>(b + s);
            int A0 = lo16(r0), B0 = lo16(r1), C0 = lo16(r2), D0 = lo16(r3), A1 = hi16(r0), B1 = hi16(r1), C1 = hi16(r2), D1 = hi16(r3);
            df_taps<false>(A0, B0, C0, D0, st, maxv);
            df_taps<false>(A1, B1, C1, D1, st, maxv);
            *reinterpret_cast<uint32_t *>(b - s) = pack16(B0, B1);
            *reinterpret_cast<uint32_t *>(b)     = pack16(C0, C1);
        }
        else {    // columns -2 -1 | 0 1, two rows
#pragma unroll
            for(int i = 0; i < 2; i++) {
                int16_t *r = b + (ptrdiff_t)i * a.s[c];
                uint32_t ab = *reinterpret_cast<uint32_t *>(r - 2), cd = *reinterpret_cast<uint32_t *>(r);
                int A = lo16(ab), B = hi16(ab), C = lo16(cd), D = hi16(cd);
                df_taps<false>(A, B, C, D, st, maxv);
                r[-1] = (int16_t)B;
                r[0]  = (int16_t)C;
            }
        }
    }
}

template <bool HOR> __global__ void __launch_bounds__(256) k_df_pass(const DfArgs a)
{
    const int xs = blockIdx.x * 32 + (threadIdx.x & 31), ys = blockIdx.y * 8 + (threadIdx.x >> 5);
    if(xs >= a.w_scu || ys >= a.h_scu) return;
    const int     bit = HOR ? 2 : 1, nb = HOR ? a.w_scu : 1; // distance to the SCU across the edge
    const int     p = ys * a.w_scu + xs;
    if(!(a.flags[p] & bit)) return;
    const int maxv = (1 << a.bd) - 1;
    // the luma samples of the segment are fetched BEFORE the strength is known: the loads do not depend on the map reads, so
    // both round trips to memory overlap (a segment whose strength turns out to be 0 costs one wasted read, nothing is written)
    int16_t        *b = a.pl[0] + (ptrdiff_t)(ys * 4) * a.s[0] + xs * 4;
    const ptrdiff_t s = a.s[0];
    uint2           r0, r1, r2, r3; // HOR: rows -2 -1 | 0 1 (4 columns each); VER: rows 0..3 (columns -2 -1 | 0 1)
    if(HOR) {
        r0 = *reinterpret_cast<uint2 *>(b - 2 * s); r1 = *reinterpret_cast<uint2 *>(b - s);
        r2 = *reinterpret_cast<uint2 *>(b); r3 = *reinterpret_cast<uint2 *>(b + s);
    }
    else {
        r0.x = *reinterpret_cast<uint32_t *>(b - 2); r0.y = *reinterpret_cast<uint32_t *>(b);
        r1.x = *reinterpret_cast<uint32_t *>(b + s - 2); r1.y = *reinterpret_cast<uint32_t *>(b + s);
        r2.x = *reinterpret_cast<uint32_t *>(b + 2 * s - 2); r2.y = *reinterpret_cast<uint32_t *>(b + 2 * s);
        r3.x = *reinterpret_cast<uint32_t *>(b + 3 * s - 2); r3.y = *reinterpret_cast<uint32_t *>(b + 3 * s);
    }
    int       cls = df_class(a.scu, a.refi, a.mv, p, p - nb), qp = DF_MCU_QP(a.scu[p]);
    const int st = xb200_df_strength(cls, qp) << (a.bd - 8);
    if(st) {
        if(HOR) { // 4 columns; rows -2 -1 | 0 1 (8-byte aligned row segments)
            int A[4] = {lo16(r0.x), hi16(r0.x), lo16(r0.y), hi16(r0.y)}, B[4] = {lo16(r1.x), hi16(r1.x), lo16(r1.y), hi16(r1.y)};
            int C[4] = {lo16(r2.x), hi16(r2.x), lo16(r2.y), hi16(r2.y)}, D[4] = {lo16(r3.x), hi16(r3.x), lo16(r3.y), hi16(r3.y)};
#pragma unroll
            for(int i = 0; i < 4; i++) df_taps<true>(A[i], B[i], C[i], D[i], st, maxv);
            *reinterpret_cast<uint2 *>(b - 2 * s) = make_uint2(pack16(A[0], A[1]), pack16(A[2], A[3]));
            *reinterpret_cast<uint2 *>(b - s)     = make_uint2(pack16(B[0], B[1]), pack16(B[2], B[3]));
            *reinterpret_cast<uint2 *>(b)         = make_uint2(pack16(C[0], C[1]), pack16(C[2], C[3]));
            *reinterpret_cast<uint2 *>(b + s)     = make_uint2(pack16(D[0], D[1]), pack16(D[2], D[3]));
        }
        else {    // 4 rows; columns -2 -1 | 0 1
            const uint2 rr[4] = {r0, r1, r2, r3};
#pragma unroll
            for(int i = 0; i < 4; i++) {
                int16_t *r = b + i * s;
                int A = lo16(rr[i].x), B = hi16(rr[i].x), C = lo16(rr[i].y), D = hi16(rr[i].y);
                df_taps<true>(A, B, C, D, st, maxv);
                *reinterpret_cast<uint32_t *>(r - 2) = pack16(A, B);
                *reinterpret_cast<uint32_t *>(r)     = pack16(C, D);
            }
        }
    }
    // chroma: only the head of a run of consecutively flagged SCUs (along the filtering direction) works, sequentially
    const int along = HOR ? ys : xs, lim = HOR ? a.h_scu : a.w_scu, stride = HOR ? a.w_scu : 1;
    if(along > 0 && (a.flags[p - stride] & bit)) return; // along == 0 cannot be flagged, so p - stride exists
    int cx = xs, cy = ys, pp = p;
    for(int k = along;;) {
        df_chroma_segment<HOR>(a, cx, cy, cls, qp);
        k++; pp += stride;
        if(k >= lim || !(a.flags[pp] & bit)) break;
        if(HOR) cy++; else cx++;
        cls = df_class(a.scu, a.refi, a.mv, pp, pp - nb);
        qp  = DF_MCU_QP(a.scu[pp]);
    }
}
