// xb200_intra.cu -- intra analysis operator (third translation unit).  Keeps its own copies of the
// quantiser tables (static __constant__) so it does not have to be compiled together with the inter operators.
#define XB200_CONST_LINKAGE static
#include "xb200_ctx.h"
#include "xb200_intra.cuh"
#include <math.h>
#include <vector>

namespace {

struct IntraCheck {
    const PicDev *pics;
    int           n_pics, w, h, n_rates;
    long long     n_states, side_elems, elems;
};
// size-class binning (order[k * n + ...] = indices of the CUs with log2 size k + 2, warp-aggregated) and argument checks:
// bins[5] unsupported shape, bins[6] invalid argument
__global__ void k_intra_bin(const xb200_intra_item *__restrict__ items, int n, int32_t *__restrict__ order, int *__restrict__ bins, IntraCheck ck)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    int       key = 7;
    if(i < n) {
        const xb200_intra_item &it = items[i];
        const int l2 = it.log2_cuw;
        if(l2 != it.log2_cuh || l2 < 2 || l2 > 6) { key = 7; atomicOr(&bins[5], 1); }
        else {
            const long long sz = (3ll << (2 * l2)) >> 1;
            const bool bad = it.cur_pic < 0 || it.cur_pic >= ck.n_pics || !ck.pics[it.cur_pic].valid || it.x < 0 || it.y < 0 ||
                             it.x + (1 << l2) > ck.w || it.y + (1 << l2) > ck.h || it.slice_type > 2 || it.rate_idx < 0 ||
                             it.rate_idx >= ck.n_rates || it.state_in < 0 || it.state_in >= ck.n_states || it.state_out < 0 ||
                             it.state_out >= ck.n_states || it.nb_off < 0 || it.nb_off + 8 * (1 << l2) + 6 > ck.side_elems || it.out_off < 0 ||
                             it.out_off + sz > ck.elems || it.ctx_skip > 1 || it.ctx_pred_mode > 2 || it.mpm[0] > 4 || it.mpm[1] > 4 ||
                             it.mpm[2] > 4 || it.mpm[3] > 4 || it.mpm[4] > 4;
            if(bad) { key = 7; atomicOr(&bins[6], 1); }
            else key = l2 - 2;
        }
    }
#pragma unroll
    for(int k = 0; k < 5; k++) {
        const unsigned m = __ballot_sync(0xffffffffu, key == k);
        if(m == 0) continue;
        const int leader = __ffs(m) - 1;
        int       base = 0;
        if(lane == leader) base = atomicAdd(&bins[k], __popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if(key == k) order[(size_t)k * n + base + __popc(m & ((1u << lane) - 1))] = i;
    }
}

template <int L2> constexpr int intra_smem() { return 8192 + IntraCfg<L2>::TEAMS * (int)sizeof(IntraTeam<L2>); }

template <int L2>
int launch_intra(xb200_ctx *c, xb200_intra_item *d_items, const int32_t *d_order, int cnt, const xb200_rates *d_rates, const xb200_sbac *d_st0,
                 xb200_sbac *d_st1, const int16_t *d_side, int16_t *d_coef, int16_t *d_rec)
{
    if(cnt == 0) return XB200_OK;
    using Cf = IntraCfg<L2>;
    const int teams_needed = (cnt + Cf::TEAMS - 1) / Cf::TEAMS;
    const int resident = L2 <= 4 ? c->sms * 8 : c->sms * 2;   // persistent CTAs: a few per SM
    const int grid = teams_needed < resident ? teams_needed : resident;
    k_intra<L2><<<grid, Cf::CTA, intra_smem<L2>(), c->side[L2 <= 3 ? 0 : L2 - 3]>>>(c->d_pics, d_items, d_order, cnt, d_rates, d_st0,
                                                                                                     d_st1, d_side, d_coef, d_rec, c->d_tm64, c->sq);
    c->launches++;
    CK(cudaGetLastError());
    return XB200_OK;
}

int intra_init(xb200_ctx *c)
{
    if(c->intra_ready) return XB200_OK;
    const int32_t qs[6] = XB200_QUANT_SCALE, dq[6] = XB200_DEQUANT_SCALE;
    static int64_t es[7][6][7];
    for(int b = 0; b < 7; b++)
        for(int q = 0; q < 6; q++)
            for(int l2 = 0; l2 < 7; l2++) es[b][q][l2] = xb200_err_scale(q, l2, b + 8);
    static int8_t tm[64 * 64];
    xb200_gen_tm64(tm);
    CK(cudaMemcpyToSymbol(c_tm64, tm, sizeof(tm)));
    static const uint8_t mpm[6][6][5] = XB200_MPM_TABLE;
    CK(cudaMemcpyToSymbol(c_mpm_tbl, mpm, sizeof(mpm)));
    CK(cudaMemcpyToSymbol(c_quant_scale, qs, sizeof(qs)));
    CK(cudaMemcpyToSymbol(c_dequant_scale, dq, sizeof(dq)));
    CK(cudaMemcpyToSymbol(c_err_scale, es, sizeof(es)));
    CK(cudaFuncSetAttribute(k_intra<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, intra_smem<2>()));
    CK(cudaFuncSetAttribute(k_intra<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, intra_smem<3>()));
    CK(cudaFuncSetAttribute(k_intra<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, intra_smem<4>()));
    CK(cudaFuncSetAttribute(k_intra<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, intra_smem<5>()));
    CK(cudaFuncSetAttribute(k_intra<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, intra_smem<6>()));
    c->intra_ready = true;
    return XB200_OK;
}

} // namespace

int xb200_analyze_intra(xb200_ctx *c, xb200_intra_item *items, int64_t n, const xb200_rates *rates, int64_t n_rates, xb200_sbac *states,
                        int64_t n_states, const int16_t *side, int64_t side_elems, int16_t *coef, int16_t *rec, int64_t elems)
{
    if(!c || n < 0 || n > (1 << 26) || n_rates < 0 || n_states < 0 || side_elems < 0 || elems < 0 ||
       (n && (!items || !rates || !states || !side || !coef)))
        return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    if(n == 0) return XB200_OK;
    int r;
    if((r = intra_init(c))) return r;
    if((r = xb200_sync_pics(c))) return r;
    xb200_intra_item *d_items;
    xb200_rates      *d_rates;
    xb200_sbac       *d_st0, *d_st1;
    int16_t          *d_side;
    int32_t          *d_order;
    if((r = to_dev(c, c->b_in_items, items, (size_t)n, XB200_MEM_HOST, &d_items))) return r;
    if((r = to_dev(c, c->b_in_rates, rates, (size_t)n_rates, XB200_MEM_HOST, &d_rates))) return r;
    if((r = to_dev(c, c->b_in_st0, states, (size_t)n_states, XB200_MEM_HOST, &d_st0))) return r;
    if((r = xb200_ensure(c->b_in_st1, (size_t)n_states * sizeof(xb200_sbac) + 64))) return r;   // output states start as a copy of the input
    d_st1 = static_cast<xb200_sbac *>(c->b_in_st1.p);
    CK(cudaMemcpyAsync(d_st1, d_st0, (size_t)n_states * sizeof(xb200_sbac), cudaMemcpyDeviceToDevice, c->stream));
    if((r = to_dev(c, c->b_in_side, side, (size_t)side_elems, XB200_MEM_HOST, &d_side))) return r;
    if((r = xb200_ensure(c->b_in_order, (size_t)5 * n * sizeof(int32_t) + 64))) return r;
    d_order = static_cast<int32_t *>(c->b_in_order.p);
    if((r = xb200_ensure(c->b_in_coef, (size_t)elems * 2 + 64))) return r;
    if(rec && (r = xb200_ensure(c->b_in_rec, (size_t)elems * 2 + 64))) return r;
    int16_t *d_coef = static_cast<int16_t *>(c->b_in_coef.p), *d_rec = rec ? static_cast<int16_t *>(c->b_in_rec.p) : nullptr;
    // the records are checked and binned by CU size on the device (a host pass over 170 k records costs more than the 4x4 kernel)
    CK(cudaMemsetAsync(c->d_bins, 0, sizeof(int) * 16, c->stream));
    const IntraCheck ck = {c->d_pics, (int)c->pics.size(), c->seq.w, c->seq.h, (int)n_rates, (long long)n_states, (long long)side_elems,
                           (long long)elems};
    k_intra_bin<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_items, (int)n, d_order, c->d_bins, ck);
    c->launches++;
    int bins[16];
    CK(cudaMemcpyAsync(bins, c->d_bins, sizeof(bins), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if(bins[5]) return XB200_ERR_UNSUPPORTED;
    if(bins[6]) return XB200_ERR_INVALID_ARGUMENT;
    const int cnt[5] = {bins[0], bins[1], bins[2], bins[3], bins[4]};
    const int first[5] = {0, (int)n, 2 * (int)n, 3 * (int)n, 4 * (int)n};
    CK(cudaEventRecord(c->ev0, c->stream));
    CK(cudaMemsetAsync(d_coef, 0, (size_t)elems * 2, c->stream));
    // the five size classes run concurrently on the side streams (fork / join around the caller-visible stream)
    CK(cudaEventRecord(c->ev_fork, c->stream));
    for(int i = 0; i < 4; i++) CK(cudaStreamWaitEvent(c->side[i], c->ev_fork, 0));
    {   // 4x4 and 8x8 CUs: one thread per CU (XB200_INTRA_SMALL=team selects the warp-per-CU kernels instead; identical results)
        const char *e = getenv("XB200_INTRA_SMALL");
        if(e && e[0] == 't') {
            if((r = launch_intra<2>(c, d_items, d_order + first[0], cnt[0], d_rates, d_st0, d_st1, d_side, d_coef, d_rec))) return r;
            if((r = launch_intra<3>(c, d_items, d_order + first[1], cnt[1], d_rates, d_st0, d_st1, d_side, d_coef, d_rec))) return r;
        }
        else {
            if(cnt[0]) {
                k_intra_thr<2><<<(cnt[0] + I4_THREADS - 1) / I4_THREADS, I4_THREADS, 0, c->side[0]>>>(c->d_pics, d_items, d_order + first[0], cnt[0],
                                                                                                     d_rates, d_st0, d_st1, d_side, d_coef, d_rec, c->sq);
                c->launches++;
            }
            if(cnt[1]) {
                k_intra_thr<3><<<(cnt[1] + I4_THREADS - 1) / I4_THREADS, I4_THREADS, 0, c->side[3]>>>(c->d_pics, d_items, d_order + first[1], cnt[1],
                                                                                                     d_rates, d_st0, d_st1, d_side, d_coef, d_rec, c->sq);
                c->launches++;
            }
            CK(cudaGetLastError());
        }
    }
    if((r = launch_intra<4>(c, d_items, d_order + first[2], cnt[2], d_rates, d_st0, d_st1, d_side, d_coef, d_rec))) return r;
    if((r = launch_intra<5>(c, d_items, d_order + first[3], cnt[3], d_rates, d_st0, d_st1, d_side, d_coef, d_rec))) return r;
    if((r = launch_intra<6>(c, d_items, d_order + first[4], cnt[4], d_rates, d_st0, d_st1, d_side, d_coef, d_rec))) return r;
    for(int i = 0; i < 4; i++) {
        CK(cudaEventRecord(c->ev_join[i], c->side[i]));
        CK(cudaStreamWaitEvent(c->stream, c->ev_join[i], 0));
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    if((r = to_host(c, items, d_items, (size_t)n, XB200_MEM_HOST))) return r;
    if((r = to_host(c, states, d_st1, (size_t)n_states, XB200_MEM_HOST))) return r;
    if((r = to_host(c, coef, d_coef, (size_t)elems, XB200_MEM_HOST))) return r;
    if(rec && (r = to_host(c, rec, d_rec, (size_t)elems, XB200_MEM_HOST))) return r;
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_ms = ms;
    CK(cudaGetLastError());
    return XB200_OK;
}

int xb200_intra_nbr(xb200_ctx *c, int32_t pic, xb200_nbr_item *items, int64_t n, const uint32_t *map_scu, const int8_t *map_ipm, int w_scu,
                    int h_scu, int constrained_intra_pred, int16_t *side, int64_t side_elems)
{
    if(!c || !pic_ok(c, pic) || n < 0 || n > (1 << 26) || side_elems < 0 || (n && (!items || !map_scu || !map_ipm || !side)))
        return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    if(n == 0) return XB200_OK;
    const Pic &p = c->pics[pic];
    if(w_scu != (p.w[0] + 3) / 4 || h_scu != (p.h[0] + 3) / 4) return XB200_ERR_INVALID_ARGUMENT;
    for(int64_t i = 0; i < n; i++) {
        const xb200_nbr_item &it = items[i];
        if(it.log2_cuw < 2 || it.log2_cuw > 6 || it.log2_cuh != it.log2_cuw) return XB200_ERR_UNSUPPORTED;
        if(it.x < 0 || it.y < 0 || (it.x & 3) || (it.y & 3) || it.x + (1 << it.log2_cuw) > p.w[0] || it.y + (1 << it.log2_cuh) > p.h[0] ||
           it.nb_off < 0 || it.nb_off + 8 * (1 << it.log2_cuw) + 6 > side_elems)
            return XB200_ERR_INVALID_ARGUMENT;
    }
    int r;
    if((r = intra_init(c))) return r;
    if((r = xb200_sync_pics(c))) return r;
    const size_t    f = (size_t)w_scu * h_scu;
    xb200_nbr_item *d_items;
    uint32_t       *d_scu;
    int8_t         *d_ipm;
    if((r = to_dev(c, c->b_in_items, items, (size_t)n, XB200_MEM_HOST, &d_items))) return r;
    if((r = to_dev(c, c->b_in_st0, map_scu, f, XB200_MEM_HOST, &d_scu))) return r;
    if((r = to_dev(c, c->b_in_st1, map_ipm, f, XB200_MEM_HOST, &d_ipm))) return r;
    if((r = xb200_ensure(c->b_in_side, (size_t)side_elems * 2 + 64))) return r;
    int16_t *d_side = static_cast<int16_t *>(c->b_in_side.p);
    CK(cudaEventRecord(c->ev0, c->stream));
    k_intra_nbr<<<(unsigned)((n + 3) / 4), 128, 0, c->stream>>>(c->d_pics, pic, d_items, n, d_scu, d_ipm, w_scu, h_scu, constrained_intra_pred != 0,
                                                               c->seq.bit_depth, d_side);
    c->launches++;
    CK(cudaEventRecord(c->ev1, c->stream));
    if((r = to_host(c, items, d_items, (size_t)n, XB200_MEM_HOST))) return r;
    if((r = to_host(c, side, d_side, (size_t)side_elems, XB200_MEM_HOST))) return r;
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_ms = ms;
    CK(cudaGetLastError());
    return XB200_OK;
}
