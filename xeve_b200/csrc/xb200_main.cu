// xb200_main.cu -- Main-profile operators.  First kernel of that row: the two-stage 16-bit transforms, forward and
// inverse -- the "IQT" DCT-II of sps.tool_iqt (xeve_trans / xeve_itrans, src_main/xevem_tq.c:58-334, 709-716 and
// src_main/xevem_itdq.c:302-557) and the ATS pair DST-VII / DCT-VIII (xeve_t_MxN_ats_intra / xeve_it_MxN_ats_intra,
// src_main/xevem_tq.c:336-707, src_main/xevem_itdq.c:63-300) -- over a list of blocks of any shape the Main-profile tree produces
// (2..64 samples per side, square or not).  Each stage is an int16 x int8 -> int32 matrix product followed by a rounding shift and a
// 16-bit store (forward: truncation; inverse: saturation), so the intermediate between the stages is what makes the result differ from
// a single fused 2-D transform: it has to be materialised exactly.
//
// One CTA per block, both stages in shared memory (block + intermediate: 16 KB for 64x64); the matrix rows are read through the
// read-only path (the 64x64 DCT-II matrix is 4 KB, the eight ATS matrices 2.7 KB: L1-resident).  HBM traffic is the block in and out,
// 4 bytes per sample.  The index mapping keeps every shared-memory access of the inner loops either a broadcast or conflict-free (see
// the stage comments); it has not been profiled yet -- the kernel was written after the round's GPU budget was spent.
#define XB200_NO_CONSTANTS // the __constant__ tables belong to xb200_api.cu
#include "xb200_ctx.h"
#include <math.h>

struct TrmArgs {
    const xb200_trm_item *items;
    int64_t               n, elems;
    int16_t              *blocks;
    const int8_t         *tm64, *ats; // ats: [2: DCT-VIII, DST-VII][4: 4, 8, 16, 32 points][32 * 32]
    int                   bd;
    int                  *err;
};

// the n-point matrix of one direction, staged in shared memory as int8 rows of n + 4 bytes (the padding makes a warp that walks the
// rows at a fixed column hit 32 different banks); 64 points: only rows 0..31 exist (outputs 32..63 are zero)
struct TrmMat {
    const int8_t *src;
    int           src_stride, k_step, kmax;
};
__device__ __forceinline__ TrmMat trm_matrix(const TrmArgs &a, int log2n, int ats, int dct8)
{
    TrmMat t;
    const int n = 1 << log2n;
    if(ats) { t.src = a.ats + ((dct8 ? 0 : 1) * 4 + (log2n - 2)) * 1024; t.src_stride = n; t.k_step = 1; t.kmax = n; }
    else    { t.src = a.tm64; t.src_stride = 64; t.k_step = 64 >> log2n; t.kmax = n == 64 ? 32 : n; }
    return t;
}
__device__ __forceinline__ void trm_stage_matrix(const TrmMat &t, int n, int8_t *dst)
{
    for(int e = threadIdx.x; e < t.kmax * n; e += 256) {
        const int k = e / n, x = e - k * n;
        dst[k * (n + 4) + x] = t.src[(k * t.k_step) * t.src_stride + x];
    }
}

__global__ void __launch_bounds__(256) k_transform_main(TrmArgs a)
{
    __shared__ int16_t s_a[4096], s_b[4096];
    __shared__ int8_t  s_mw[64 * 68], s_mh[64 * 68];
    for(int64_t i = blockIdx.x; i < a.n; i += gridDim.x) {
        const xb200_trm_item it = a.items[i];
        const int lw = it.log2_w, lh = it.log2_h;
        const int lo = it.ats ? 2 : 1, hi = it.ats ? 5 : 6;
        if(lw < lo || lw > hi || lh < lo || lh > hi || it.inverse > 1 || it.ats > 1 || it.tridx > 3 || it.off < 0 ||
           it.off + ((int64_t)1 << (lw + lh)) > a.elems) {
            if(threadIdx.x == 0) atomicExch(a.err, 1);
            continue;
        }
        const int w = 1 << lw, h = 1 << lh, n = w * h, pw = w + 4, ph = h + 4;
        int16_t  *blk = a.blocks + it.off;
        const TrmMat mw = trm_matrix(a, lw, it.ats, it.tridx >> 1), mh = trm_matrix(a, lh, it.ats, it.tridx & 1);
        __syncthreads(); // the previous block's write-back reads s_a, its last stage the matrices
        for(int e = threadIdx.x; e < n; e += 256) s_a[e] = blk[e];
        trm_stage_matrix(mw, w, s_mw);
        trm_stage_matrix(mh, h, s_mh);
        __syncthreads();
        if(!it.inverse) {
            const int sh1 = lw - 1 + a.bd - 8, add1 = sh1 ? 1 << (sh1 - 1) : 0, sh2 = lh + 6, add2 = 1 << (sh2 - 1);
            // rows: t[j][k] = sum_x M_w[k][x] * blk[j][x]; lanes walk k: the block sample is a broadcast, the matrix column conflict-free
            for(int e = threadIdx.x; e < n; e += 256) {
                const int k = e & (w - 1), j = e >> lw;
                int       acc = 0;
                if(k < mw.kmax) {
                    for(int x = 0; x < w; x++) acc += s_mw[k * pw + x] * s_a[j * w + x];
                    acc = (acc + add1) >> sh1;
                }
                s_b[e] = (int16_t)acc;                              // [j][k], row-major
            }
            __syncthreads();
            // columns: out[k][j] = sum_y M_h[k][y] * t[y][j]; lanes walk j: contiguous reads of t, the matrix entry is a broadcast
            for(int e = threadIdx.x; e < n; e += 256) {
                const int j = e & (w - 1), k = e >> lw;
                int       acc = 0;
                if(k < mh.kmax) {
                    for(int y = 0; y < h; y++) acc += s_mh[k * ph + y] * s_b[y * w + j];
                    acc = (acc + add2) >> sh2;
                }
                s_a[e] = (int16_t)acc;                              // [k][j]: vertical frequency major, as the reference stores it
            }
        }
        else {
            const int sh1 = 7, add1 = 64, sh2 = 12 - (a.bd - 8), add2 = 1 << (sh2 - 1);
            // columns: t[y][j] = sat16(sum_k M_h[k][y] * coef[k][j]); lanes walk j
            for(int e = threadIdx.x; e < n; e += 256) {
                const int j = e & (w - 1), y = e >> lw;
                int       acc = 0;
                for(int k = 0; k < mh.kmax; k++) acc += s_mh[k * ph + y] * s_a[k * w + j];
                acc = (acc + add1) >> sh1;
                s_b[e] = (int16_t)max(-32768, min(32767, acc));   // [y][j]
            }
            __syncthreads();
            // rows: out[y][x] = sat16(sum_k M_w[k][x] * t[y][k]); lanes walk x: t is a broadcast, the matrix row contiguous
            for(int e = threadIdx.x; e < n; e += 256) {
                const int x = e & (w - 1), y = e >> lw;
                int       acc = 0;
                for(int k = 0; k < mw.kmax; k++) acc += s_mw[k * pw + x] * s_b[y * w + k];
                acc = (acc + add2) >> sh2;
                s_a[e] = (int16_t)max(-32768, min(32767, acc));
            }
        }
        __syncthreads();
        for(int e = threadIdx.x; e < n; e += 256) blk[e] = s_a[e];
    }
}

// the 8-bit ATS matrices: rounded basis functions scaled by 64 sqrt(N) (they equal the reference's table, tests/test_oracle.py)
static void gen_ats(int8_t *out)
{
    const double pi = 3.14159265358979323846;
    for(int type = 0; type < 2; type++)
        for(int l = 2; l <= 5; l++) {
            const int    n_ = 1 << l;
            const double scale = 64.0 * sqrt((double)n_) * sqrt(4.0 / (2 * n_ + 1));
            int8_t      *m = out + (type * 4 + (l - 2)) * 1024;
            for(int k = 0; k < n_; k++)
                for(int n = 0; n < n_; n++) {
                    const double f = type == 0 ? cos(pi * (2 * k + 1) * (2 * n + 1) / (4 * n_ + 2)) : sin(pi * (2 * k + 1) * (n + 1) / (2 * n_ + 1));
                    m[k * n_ + n] = (int8_t)floor(scale * f + 0.5);
                }
        }
}

int xb200_transform_main(xb200_ctx *c, const xb200_trm_item *items, int64_t n, int16_t *blocks, int64_t elems, int mem)
{
    if(!c || n < 0 || n > (1 << 28) || elems < 0 || (n && (!items || !blocks)) || (mem != XB200_MEM_HOST && mem != XB200_MEM_DEVICE))
        return XB200_ERR_INVALID_ARGUMENT;
    if(n == 0) return XB200_OK;
    CK(cudaSetDevice(c->device));
    int r;
    if(!c->b_ats.p) {
        static int8_t h_ats[2 * 4 * 1024];
        gen_ats(h_ats);
        if((r = xb200_ensure(c->b_ats, sizeof(h_ats)))) return r;
        CK(cudaMemcpyAsync(c->b_ats.p, h_ats, sizeof(h_ats), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    xb200_trm_item *d_items;
    int16_t        *d_blocks;
    if((r = to_dev(c, c->b_items, items, (size_t)n, mem, &d_items))) return r;
    if((r = to_dev(c, c->b_aux0, blocks, (size_t)elems, mem, &d_blocks))) return r;
    CK(cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream));
    CK(cudaEventRecord(c->ev0, c->stream));
    TrmArgs a;
    a.items = d_items; a.n = n; a.elems = elems; a.blocks = d_blocks; a.tm64 = c->d_tm64;
    a.ats = static_cast<const int8_t *>(c->b_ats.p); a.bd = c->seq.bit_depth; a.err = c->d_err;
    const unsigned grid = (unsigned)(n < c->sms * 8 ? n : c->sms * 8);
    k_transform_main<<<grid, 256, 0, c->stream>>>(a);
    c->launches += 1;
    if(mem == XB200_MEM_HOST) CK(cudaMemcpyAsync(blocks, d_blocks, (size_t)elems * sizeof(int16_t), cudaMemcpyDeviceToHost, c->stream));
    int err = 0;
    CK(cudaMemcpyAsync(&err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if((r = xb200_finish(c))) return r;
    if(err) { cudaMemset(c->d_err, 0, sizeof(int)); return XB200_ERR_INVALID_ARGUMENT; }
    return XB200_OK;
}
