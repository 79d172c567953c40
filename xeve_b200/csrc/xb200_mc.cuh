// xb200_mc.cuh -- batched motion compensation (8-tap luma / 4-tap chroma, uni + bi), one CTA per
// pi->fn_mc call.  Replaces reference src_base/xeve_mc.c:99-381 (xeve_mc_l_* / xeve_mc_c_*),
// 401-447 (xeve_mv_clip), 449-463 (xeve_average_16b_no_clip), 465-610 (xeve_mc).
#pragma once
#include "xb200_common.cuh"

#define MC_THREADS 128

// One plane, one list.  (gx, gy) = absolute position in 1/FR sample units (FR = 4 luma, 8 chroma),
// already formed from the CLIPPED mv; want_h / want_v come from the UNCLIPPED mv fraction
// (reference quirk: src_base/xeve_mc.c:490-516).  dst has stride bw.  tmp holds (bh+TAPS-1)*bw.
// Must be called by all threads of the CTA (contains barriers).
template <int TAPS, int FR>
XB_DEV void mc_plane(const int16_t *__restrict__ ref, int sr, int gx, int gy, bool want_h, bool want_v,
                     int16_t *__restrict__ dst, int bw, int bh, int bd, int16_t *__restrict__ tmp, int tid, int nthr)
{
    constexpr int HALF = TAPS / 2 - 1;
    constexpr int FSH  = FR == 4 ? 2 : 3;
    const int     ix = gx >> FSH, iy = gy >> FSH, fx = gx & (FR - 1), fy = gy & (FR - 1);
    const int     maxv = (1 << bd) - 1;
    const int16_t *ch = TAPS == 8 ? c_mc_l[fx] : c_mc_c[fx];
    const int16_t *cv = TAPS == 8 ? c_mc_l[fy] : c_mc_c[fy];
    const int16_t *base = ref + (ptrdiff_t)iy * sr + ix;
    if(!want_h && !want_v) {
        for(int e = tid; e < bw * bh; e += nthr) dst[e] = base[(ptrdiff_t)(e / bw) * sr + e % bw];
    }
    else if(want_h != want_v) {
        const ptrdiff_t stp = want_h ? 1 : sr;
        const int16_t  *cf  = want_h ? ch : cv;
        for(int e = tid; e < bw * bh; e += nthr) {
            const int16_t *p = base + (ptrdiff_t)(e / bw) * sr + e % bw - HALF * stp;
            int acc = 0;
#pragma unroll
            for(int t = 0; t < TAPS; t++) acc += cf[t] * p[t * stp];
            dst[e] = (int16_t)clip3i(0, maxv, acc >> 6);
        }
    }
    else {
        const int s1 = min(4, bd - 8), s2 = max(8, 20 - bd);
        for(int e = tid; e < bw * (bh + TAPS - 1); e += nthr) {
            const int16_t *p = base + (ptrdiff_t)(e / bw - HALF) * sr + e % bw - HALF;
            int acc = 0;
#pragma unroll
            for(int t = 0; t < TAPS; t++) acc += ch[t] * p[t];
            tmp[e] = (int16_t)(acc >> s1);
        }
        __syncthreads();
        for(int e = tid; e < bw * bh; e += nthr) {
            int acc = 0;
#pragma unroll
            for(int t = 0; t < TAPS; t++) acc += cv[t] * tmp[e + t * bw];
            dst[e] = (int16_t)clip3i(0, maxv, (acc + (1 << (s2 - 1))) >> s2);
        }
    }
    __syncthreads();
}

// Full xeve_mc for one item into `pred` (Y | U | V, stride = block width).  `aux` = second
// prediction buffer of the same size (bi), `tmp` = interpolation scratch (71*64 samples).
XB_DEV void mc_item(const PicDev *__restrict__ pics, const xb200_mc_item &it, const SeqDev &sq, int16_t *pred, int16_t *aux,
                    int16_t *tmp, int tid, int nthr, bool luma_only = false)
{
    const int w = it.w, h = it.h, x = it.x, y = it.y, ny = w * h, nc = ny >> 2;
    int       mvt[2][2];
#pragma unroll
    for(int l = 0; l < 2; l++) {
        mvt[l][0] = it.mv[l][0]; mvt[l][1] = it.mv[l][1];
        if(it.refi[l] >= 0) { // xeve_mv_clip
            const int lo = -(128 << 2), hx = (sq.w - 1 + 128) << 2, hy = (sq.h - 1 + 128) << 2;
            if((x << 2) + it.mv[l][0] < lo) mvt[l][0] = lo - (x << 2);
            if((y << 2) + it.mv[l][1] < lo) mvt[l][1] = lo - (y << 2);
            if((x << 2) + it.mv[l][0] + (w << 2) - 4 > hx) mvt[l][0] = hx - (x << 2) - (w << 2) + 4;
            if((y << 2) + it.mv[l][1] + (h << 2) - 4 > hy) mvt[l][1] = hy - (y << 2) - (h << 2) + 4;
            mvt[l][0] = (int16_t)mvt[l][0]; mvt[l][1] = (int16_t)mvt[l][1];
        }
    }
    int n = 0;
    for(int l = 0; l < 2; l++) {
        if(it.refi[l] < 0) continue;
        if(l == 1 && it.refi[0] >= 0 && it.ref_poc[0] == it.ref_poc[1] && mvt[0][0] == mvt[1][0] && mvt[0][1] == mvt[1][1])
            break; // identical motion: list 0 prediction stands (src_base/xeve_mc.c:545-551)
        const PicDev rp  = pics[it.ref_pic[l]];
        int16_t     *dst = n == 0 ? pred : aux;
        const int    gx = (x << 2) + mvt[l][0], gy = (y << 2) + mvt[l][1];
        const bool   hl = (it.mv[l][0] & 3) != 0, vl = (it.mv[l][1] & 3) != 0;
        const bool   hc = (it.mv[l][0] & 7) != 0, vc = (it.mv[l][1] & 7) != 0;
        mc_plane<8, 4>(rp.p[0], rp.s[0], gx, gy, hl, vl, dst, w, h, sq.bd, tmp, tid, nthr);
        if(!luma_only) {
            mc_plane<4, 8>(rp.p[1], rp.s[1], gx, gy, hc, vc, dst + ny, w >> 1, h >> 1, sq.bd, tmp, tid, nthr);
            mc_plane<4, 8>(rp.p[2], rp.s[2], gx, gy, hc, vc, dst + ny + nc, w >> 1, h >> 1, sq.bd, tmp, tid, nthr);
        }
        n++;
    }
    if(n == 2) {
        for(int e = tid; e < (luma_only ? ny : ny + 2 * nc); e += nthr) pred[e] = (int16_t)((pred[e] + aux[e] + 1) >> 1);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(MC_THREADS) k_mc(const PicDev *__restrict__ pics, const xb200_mc_item *__restrict__ items,
                                                    int n, const int64_t *__restrict__ pred_off, int16_t *__restrict__ out, SeqDev sq)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int16_t *pred = reinterpret_cast<int16_t *>(smem_raw); // 6144
    int16_t *aux  = pred + 6144;                           // 6144
    int16_t *tmp  = aux + 6144;                            // 71 * 64
    const int i = blockIdx.x;
    if(i >= n) return;
    const xb200_mc_item it = items[i];
    mc_item(pics, it, sq, pred, aux, tmp, threadIdx.x, MC_THREADS);
    const int total = it.w * it.h * 3 / 2;
    int16_t  *o     = out + pred_off[i];
    for(int e = threadIdx.x; e < total; e += MC_THREADS) o[e] = pred[e];
}
// get_org_bi after fn_mc: org_bi = 2*org - pred (luma), contiguous w*h block per item
__global__ void __launch_bounds__(MC_THREADS) k_bi_org(const PicDev *__restrict__ pics, const xb200_mc_item *__restrict__ items,
                                                        int n, const int32_t *__restrict__ cur_pic, const int64_t *__restrict__ off,
                                                        int16_t *__restrict__ side, SeqDev sq)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int16_t *pred = reinterpret_cast<int16_t *>(smem_raw);
    int16_t *aux  = pred + 6144;
    int16_t *tmp  = aux + 6144;
    const int i = blockIdx.x;
    if(i >= n) return;
    const xb200_mc_item it = items[i];
    if(it.w == 0) return; // slot left empty by the CU pipeline
    mc_item(pics, it, sq, pred, aux, tmp, threadIdx.x, MC_THREADS, true);
    const PicDev   o   = pics[cur_pic[i]];
    const int16_t *org = o.p[0] + (ptrdiff_t)it.y * o.s[0] + it.x;
    int16_t       *dst = side + off[i];
    for(int e = threadIdx.x; e < it.w * it.h; e += MC_THREADS)
        dst[e] = (int16_t)(((int16_t)org[(ptrdiff_t)(e / it.w) * o.s[0] + e % it.w] << 1) - pred[e]);
}
#define MC_SMEM_BYTES ((6144 * 2 + 71 * 64) * 2)
