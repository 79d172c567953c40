// xb200_misc.cuh -- picture preparation, distortion probes, MV-predictor inputs.
// (The fused candidate-evaluation kernel lives in xb200_residue2.cuh.)
#pragma once
#include "xb200_common.cuh"
#include "xb200_mc.cuh"
#include "xb200_tq.cuh"
#include "xb200_had.cuh"

// ---- picture upload: input depth -> internal depth (reference src_base/xeve_util.c:1552-1571, 1670-1704)
//      and border replication (reference src_base/xeve_util.c:190-248) ----------------------------------
template <typename T>
__global__ void k_convert(const T *__restrict__ src, int src_stride_elems, int16_t *__restrict__ dst, int dst_stride, int w, int h,
                          int shift)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if(x < w && y < h) dst[(size_t)y * dst_stride + x] = (int16_t)((int)src[(size_t)y * src_stride_elems + x] << shift);
}

// Border replication of all three planes in ONE launch that only touches the border (xeve_picbuf_expand,
// src_base/xeve_util.c:190-248).  One thread writes one aligned group of 4 samples (an 8-byte store): groups of the rows above /
// below the picture copy (or, in the corners, broadcast) from the first / last active row, groups of the left / right strips
// broadcast the edge sample of their row.  Plane widths and paddings are multiples of 4 samples, strides multiples of 64 and the
// active origin is 16-byte aligned (xb200_pic_create), so every group is aligned.
struct PadArgs {
    int16_t *act[3];
    int      s[3], w[3], h[3], pad[3];
    int      first[4]; // first group index of plane 0, 1, 2 and the total
};
__host__ __device__ inline int pad_plane_groups(int w, int h, int pad) { return 2 * pad * ((w + 2 * pad) >> 2) + h * 2 * (pad >> 2); }
__global__ void __launch_bounds__(256) k_pad3(const PadArgs a)
{
    int g = blockIdx.x * 256 + threadIdx.x;
    if(g >= a.first[3]) return;
    const int k = g >= a.first[2] ? 2 : (g >= a.first[1] ? 1 : 0);
    g -= a.first[k];
    const int w = a.w[k], h = a.h[k], pad = a.pad[k], stride = a.s[k], gpr = (w + 2 * pad) >> 2, gps = pad >> 2;
    int16_t  *act = a.act[k];
    int       x, y;
    if(g < 2 * pad * gpr) { // rows above / below the picture
        const int row = g / gpr;
        y = row < pad ? row - pad : h + (row - pad);
        x = (g % gpr) * 4 - pad;
    }
    else {                  // left / right strips of an active row
        g -= 2 * pad * gpr;
        y = g / (2 * gps);
        const int c = g % (2 * gps);
        x = c < gps ? c * 4 - pad : w + (c - gps) * 4;
    }
    const int16_t *srow = act + (ptrdiff_t)min(max(y, 0), h - 1) * stride;
    uint2          v;
    if(x >= 0 && x < w) v = *reinterpret_cast<const uint2 *>(srow + x); // above / below the active columns: copy
    else {
        const uint32_t e = (uint16_t)srow[x < 0 ? 0 : w - 1];
        v.x = v.y = e | (e << 16);
    }
    *reinterpret_cast<uint2 *>(act + (ptrdiff_t)y * stride + x) = v;
}

// ---- distortion probes: one warp per item, blocks addressed inside device pictures -------------------------
XB_DEV const int16_t *blk_ptr(const PicDev *pics, int pic, int plane, int x, int y, int &stride)
{
    const PicDev p = pics[pic];
    stride         = p.s[plane];
    return p.p[plane] + (ptrdiff_t)y * stride + x;
}

// XEVE_FN_SAD (src_base/xeve_sad.c:40-61); the shift is applied to the sum
__global__ void k_sad(const PicDev *__restrict__ pics, const xb200_blk_item *__restrict__ items, int64_t n, int32_t *__restrict__ out, int bd)
{
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if(i >= n) return;
    const xb200_blk_item it = items[i];
    int            s1, s2;
    const int16_t *a = blk_ptr(pics, it.pic1, it.plane1, it.x1, it.y1, s1), *b = blk_ptr(pics, it.pic2, it.plane2, it.x2, it.y2, s2);
    const int      w = 1 << it.log2w, h = 1 << it.log2h, lane = threadIdx.x & 31;
    int            sum = 0;
    for(int e = lane; e < w * h; e += 32) {
        const int d = (int)a[(ptrdiff_t)(e / w) * s1 + e % w] - (int)b[(ptrdiff_t)(e / w) * s2 + e % w];
        sum += (d ^ (d >> 15)) - (d >> 15); // XEVE_ABS16 applied to the int difference
    }
#pragma unroll
    for(int m = 16; m > 0; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
    if(lane == 0) out[i] = sum >> (bd - 8);
}

// XEVE_FN_SSD (src_base/xeve_sad.c:275-297); the shift is applied to every squared difference
__global__ void k_ssd(const PicDev *__restrict__ pics, const xb200_blk_item *__restrict__ items, int64_t n, int64_t *__restrict__ out, int bd)
{
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if(i >= n) return;
    const xb200_blk_item it = items[i];
    int            s1, s2;
    const int16_t *a = blk_ptr(pics, it.pic1, it.plane1, it.x1, it.y1, s1), *b = blk_ptr(pics, it.pic2, it.plane2, it.x2, it.y2, s2);
    const int      w = 1 << it.log2w, h = 1 << it.log2h, lane = threadIdx.x & 31, sh = (bd - 8) << 1;
    int64_t        sum = 0;
    for(int e = lane; e < w * h; e += 32) {
        const int d = (int)a[(ptrdiff_t)(e / w) * s1 + e % w] - (int)b[(ptrdiff_t)(e / w) * s2 + e % w];
        sum += (d * d) >> sh;
    }
#pragma unroll
    for(int m = 16; m > 0; m >>= 1) sum += (int64_t)shfl_xor_u64((uint64_t)sum, m);
    if(lane == 0) out[i] = sum;
}


// XEVE_FN_SATD = xeve_had for square blocks (src_base/xeve_sad.c:1043-1140)
__global__ void k_satd(const PicDev *__restrict__ pics, const xb200_blk_item *__restrict__ items, int64_t n, int32_t *__restrict__ out, int bd)
{
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if(i >= n) return;
    const xb200_blk_item it = items[i];
    int            s1, s2;
    const int16_t *a = blk_ptr(pics, it.pic1, it.plane1, it.x1, it.y1, s1), *b = blk_ptr(pics, it.pic2, it.plane2, it.x2, it.y2, s2);
    const int      w = 1 << it.log2w, h = 1 << it.log2h, lane = threadIdx.x & 31;
    int            sum = 0;
    if((w & 7) == 0 && (h & 7) == 0) {
        const int tw = w >> 3, nt = tw * (h >> 3);
        for(int t = lane; t < nt; t += 32) {
            const int tx = (t % tw) * 8, ty = (t / tw) * 8;
            sum += had_tile_dev<8>(a + (ptrdiff_t)ty * s1 + tx, s1, b + (ptrdiff_t)ty * s2 + tx, s2);
        }
    }
    else {
        const int tw = w >> 2, nt = tw * (h >> 2);
        for(int t = lane; t < nt; t += 32) {
            const int tx = (t % tw) * 4, ty = (t / tw) * 4;
            sum += had_tile_dev<4>(a + (ptrdiff_t)ty * s1 + tx, s1, b + (ptrdiff_t)ty * s2 + tx, s2);
        }
    }
#pragma unroll
    for(int m = 16; m > 0; m >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, m);
    if(lane == 0) out[i] = sum >> (bd - 8);
}

// ---- MV-predictor inputs ---------------------------------------------------------------------
// xeve_get_avail_inter (src_base/xeve_util.c:652-715, one tile), xeve_get_motion (:526-573), xeve_get_mv_dir (:619-650)
__global__ void k_mvp(xb200_mvp_item *__restrict__ items, int64_t n, xb200_mvp_pic pp, const uint32_t *__restrict__ map_scu,
                      const int16_t *__restrict__ map_mv, const int16_t *__restrict__ col0, const int16_t *__restrict__ col1)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    xb200_mvp_item it = items[i];
    const int x = it.x_scu, y = it.y_scu, w = pp.w_scu, h = pp.h_scu, scuw = (1 << it.log2_cuw) >> 2, scuh = (1 << it.log2_cuh) >> 2;
    const int scup = x + y * w, l = it.lidx;
    auto COD = [&](int p) { return (int)((map_scu[p] >> 31) & 1); };
    auto IF  = [&](int p) { return (int)((map_scu[p] >> 15) & 1); };
    auto IBC = [&](int p) { return (int)((map_scu[p] >> 26) & 1); };
    unsigned av = 0;
    if(x > 0 && !IF(scup - 1) && COD(scup - 1) && !IBC(scup - 1)) {
        av |= 1u << 1; // AVAIL_LE
        if(y + scuh < h && COD(scup + scuh * w - 1) && !IF(scup + scuh * w - 1) && !IBC(scup + scuh * w - 1)) av |= 1u << 7; // LO_LE
    }
    if(y > 0) {
        if(!IF(scup - w) && !IBC(scup - w)) av |= 1u << 0;                                   // AVAIL_UP
        if(!IF(scup - w + scuw - 1) && !IBC(scup - w + scuw - 1)) av |= 1u << 9;             // AVAIL_RI_UP
        if(x > 0 && !IF(scup - w - 1) && COD(scup - w - 1) && !IBC(scup - w - 1)) av |= 1u << 5; // AVAIL_UP_LE
        if(x + scuw < w && ((map_scu[scup - w + scuw] >> 15) & 0x10001) == 0x10000 && COD(scup - w + scuw)) av |= 1u << 6; // UP_RI
    }
    if(x + scuw < w && !IF(scup + scuw) && COD(scup + scuw) && !IBC(scup + scuw)) {
        av |= 1u << 3; // AVAIL_RI
        if(y + scuh < h && COD(scup + scuh * w + scuw) && !IF(scup + scuh * w + scuw) && !IBC(scup + scuh * w + scuw)) av |= 1u << 8;
    }
    it.avail = (uint16_t)av;
    auto mvat = [&](const int16_t *m, int p, int list, int c) { return m[((size_t)p * 2 + list) * 2 + c]; };
    const int nb[3] = {scup - 1, scup - w, scup - w + scuw};
    const unsigned need[3] = {1u << 1, 1u << 0, 1u << 6};
#pragma unroll
    for(int k = 0; k < 3; k++) {
        it.refi[k] = 0;
        const bool ok = (av & need[k]) != 0;
        it.mvp[k][0] = ok ? mvat(map_mv, nb[k], l, 0) : (int16_t)1;
        it.mvp[k][1] = ok ? mvat(map_mv, nb[k], l, 1) : (int16_t)1;
    }
    const int16_t *col = l ? col1 : col0;
    it.refi[3] = 0;
    it.mvp[3][0] = mvat(col, scup, 0, 0);
    it.mvp[3][1] = mvat(col, scup, 0, 1);
    // temporal direct: colocated MV of the list-1 reference at the CU's bottom-right SCU, scaled by POC distances
    {
        const int br = scup + (scuw - 1) + (scuh - 1) * w;
        const int mx = mvat(col1, br, 0, 0), my = mvat(col1, br, 0, 1);
        const int dco = pp.ref_poc[1] - pp.col_list_poc0, d0 = pp.poc - pp.ref_poc[0], d1 = pp.ref_poc[1] - pp.poc;
        if(dco == 0) { it.mv_dir[0][0] = it.mv_dir[0][1] = it.mv_dir[1][0] = it.mv_dir[1][1] = 0; }
        else {
            it.mv_dir[0][0] = (int16_t)(d0 * mx / dco); it.mv_dir[0][1] = (int16_t)(d0 * my / dco);
            it.mv_dir[1][0] = (int16_t)(-d1 * mx / dco); it.mv_dir[1][1] = (int16_t)(-d1 * my / dco);
        }
    }
    items[i] = it;
}
