// xb200_frame.cu -- whole-picture operators of the library (second translation unit; the per-CU operators live in
// xb200_api.cu): in-loop deblocking.
#define XB200_NO_CONSTANTS // the __constant__ tables belong to xb200_api.cu
#include "xb200_ctx.h"
#include "xb200_deblock.cuh"

#define finish xb200_finish
#define ensure xb200_ensure

int xb200_preload_frame_kernels()   // see xb200_preload_api_kernels
{
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k_df_pass<false>));
    CK(cudaFuncGetAttributes(&fa, k_df_pass<true>));
    CK(cudaFuncGetAttributes(&fa, k_df_mark));
    return XB200_OK;
}

int xb200_deblock_dev(xb200_ctx *c, Pic &p, const xb200_df_pic *pp, const uint32_t *d_scu, const int8_t *d_refi, const int16_t *d_mv,
                      const uint8_t *d_flags, cudaStream_t stream)
{
    DfArgs a;
    for(int k = 0; k < 3; k++) { a.pl[k] = p.buf[k] + (size_t)p.pad[k] * p.s[k] + p.pad[k]; a.s[k] = p.s[k]; }
    a.w_scu = pp->w_scu; a.h_scu = pp->h_scu; a.bd = c->seq.bit_depth;
    a.scu = d_scu; a.refi = d_refi; a.mv = d_mv; a.flags = d_flags; a.pp = *pp;
    const dim3 grid((pp->w_scu + 31) / 32, (pp->h_scu + 7) / 8);
    k_df_pass<false><<<grid, 256, 0, stream>>>(a);
    k_df_pass<true><<<grid, 256, 0, stream>>>(a);
    c->launches += 2;
    CK(cudaGetLastError());
    return XB200_OK;
}

// ---- in-loop deblocking --------------------------------------------------------------------------------------
int xb200_deblock(xb200_ctx *c, int32_t pic, const xb200_df_cu *cus, int64_t n, const xb200_df_pic *pp, const uint32_t *map_scu,
                  const int8_t *map_refi, const int16_t *map_mv, int expand, int mem)
{
    if(!c || !pic_ok(c, pic) || n < 0 || n > (1 << 28) || !pp || (n && (!cus || !map_scu || !map_refi || !map_mv)))
        return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    Pic &p = c->pics[pic];
    if(pp->w_scu != (p.w[0] + 3) / 4 || pp->h_scu != (p.h[0] + 3) / 4 || (p.w[0] & 7) || (p.h[0] & 7)) return XB200_ERR_INVALID_ARGUMENT;
    if(mem == XB200_MEM_HOST)
        for(int64_t i = 0; i < n; i++) {
            const xb200_df_cu &u = cus[i];
            if(u.x < 0 || u.y < 0 || (u.x & 3) || (u.y & 3) || u.log2_cuw < 2 || u.log2_cuh < 2 || u.log2_cuw > 7 || u.log2_cuh > 7 ||
               u.x + (1 << u.log2_cuw) > p.w[0] || u.y + (1 << u.log2_cuh) > p.h[0])
                return XB200_ERR_INVALID_ARGUMENT;
        }
    const size_t f = (size_t)pp->w_scu * pp->h_scu;
    int r;
    xb200_df_cu *d_cus;
    uint32_t    *d_scu;
    int8_t      *d_refi;
    int16_t     *d_mv;
    if((r = to_dev(c, c->b_items, cus, (size_t)n, mem, &d_cus))) return r;
    if((r = to_dev(c, c->b_aux0, map_scu, f, mem, &d_scu))) return r;
    if((r = to_dev(c, c->b_aux1, map_refi, f * 2, mem, &d_refi))) return r;
    if((r = to_dev(c, c->b_aux2, map_mv, f * 4, mem, &d_mv))) return r;
    if((r = ensure(c->b_df, f))) return r;
    CK(cudaEventRecord(c->ev0, c->stream));
    if(n) {
        uint8_t *d_flags = static_cast<uint8_t *>(c->b_df.p);
        CK(cudaMemsetAsync(d_flags, 0, f, c->stream));
        k_df_mark<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(d_cus, n, pp->w_scu, pp->h_scu, d_flags);
        DfArgs a;
        for(int k = 0; k < 3; k++) { a.pl[k] = p.buf[k] + (size_t)p.pad[k] * p.s[k] + p.pad[k]; a.s[k] = p.s[k]; }
        a.w_scu = pp->w_scu; a.h_scu = pp->h_scu; a.bd = c->seq.bit_depth;
        a.scu = d_scu; a.refi = d_refi; a.mv = d_mv; a.flags = d_flags; a.pp = *pp;
        const dim3 grid((pp->w_scu + 31) / 32, (pp->h_scu + 7) / 8);
        k_df_pass<false><<<grid, 256, 0, c->stream>>>(a); // every vertical edge of the picture ...
        k_df_pass<true><<<grid, 256, 0, c->stream>>>(a);  // ... then every horizontal edge (src_base/xeve_enc.c:2363)
        c->launches += 3;
    }
    if(expand && (r = xb200_pad_planes(c, p))) return r;
    return finish(c);
}

