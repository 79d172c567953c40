// xb200_api.cu -- host side of the C ABI declared in include/xeve_b200.h: context, device
// picture pool, staging, kernel launches.  Single translation unit (the kernels share
// __constant__ tables).  There is deliberately no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <mutex>
#include <chrono>

#include "xb200_ctx.h"
#include "xb200_me.cuh"
#include "xb200_mc.cuh"
#include "xb200_tq.cuh"
#include "xb200_misc.cuh"
#include "xb200_residue2.cuh"
#include "xb200_dct_tc.cuh"
#include "xb200_rate.cuh"
#include "xb200_analyze.cuh"
#include "xb200_pipeline.cuh"
#include "xb200_lanecoder.cuh"
#include <math.h>

int xb200_ensure(DevBuf &b, size_t bytes)
{
    if(bytes <= b.cap) return XB200_OK;
    if(b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    CK(cudaMalloc(&b.p, want));
    b.cap = want;
    return XB200_OK;
}

int xb200_finish(xb200_ctx *c)
{
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_ms = ms;
    CK(cudaGetLastError());
    return XB200_OK;
}

namespace {

int ensure(DevBuf &b, size_t bytes) { return xb200_ensure(b, bytes); }

int sync_pics(xb200_ctx *c)
{
    if(!c->pics_dirty) return XB200_OK;
    const int n = (int)c->pics.size();
    if(n > c->d_pics_cap) {
        if(c->d_pics) cudaFree(c->d_pics);
        c->d_pics_cap = n + 16 < 4096 ? 4096 : n + 16;   // roomy: growing the table means cudaFree, which waits for every running kernel
        CK(cudaMalloc(&c->d_pics, sizeof(PicDev) * c->d_pics_cap));
    }
    std::vector<PicDev> h(n);
    for(int i = 0; i < n; i++) {
        const Pic &p = c->pics[i];
        memset(&h[i], 0, sizeof(PicDev));
        if(!p.used) continue;
        for(int k = 0; k < 3; k++) {
            h[i].p[k] = p.buf[k] + (size_t)p.pad[k] * p.s[k] + p.pad[k];
            h[i].s[k] = p.s[k];
        }
        h[i].w = p.w[0]; h[i].h = p.h[0]; h[i].pad_l = p.pad[0]; h[i].pad_c = p.pad[1]; h[i].valid = 1;
    }
    if(n) CK(cudaMemcpyAsync(c->d_pics, h.data(), sizeof(PicDev) * n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream)); // h goes out of scope
    c->pics_dirty = false;
    return XB200_OK;
}

int finish(xb200_ctx *c) { return xb200_finish(c); }

// Arguments of the caller's records are checked here, on the device, when `validate` is set (host-buffer calls): a host loop over
// the records costs more than the search of a small picture (1.3 ms per 20 MB of records).  bins[6] != 0: invalid argument.
struct BinCheck {
    const PicDev *pics;
    int           n_pics, validate, have_side, n_rates;
    long long     side_elems, elems;
};
XB_DEV bool bin_pic_ok(const BinCheck &ck, int h) { return h >= 0 && h < ck.n_pics && ck.pics[h].valid; }

__global__ void k_me_bin(const xb200_me_item *__restrict__ items, int n, int32_t *__restrict__ order, int *__restrict__ bins, BinCheck ck)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    int       key = 5, margin = 0; // key 4 = unsupported shape, 5 = out of range
    if(i < n) {
        const xb200_me_item &it = items[i];
        const int            l2 = it.log2_cuw;
        key = (l2 >= 3 && l2 <= 6 && it.log2_cuh == l2) ? l2 - 3 : (l2 == 0 ? 5 : 4); // log2 0: slot left empty by the CU pipeline
        if(ck.validate && l2 != 0) {
            // the CU must lie inside the (unpadded) current picture; the search range bounds the staged window
            const bool bad = !bin_pic_ok(ck, it.cur_pic) || !bin_pic_ok(ck, it.ref_pic) || ck.pics[it.ref_pic].pad_l == 0 || it.gop_size <= 0 ||
                             it.max_search_range < 1 || it.max_search_range > 256 || it.x < 0 || it.y < 0 ||
                             it.x + (1 << l2) > ck.pics[it.cur_pic].w || it.y + (1 << l2) > ck.pics[it.cur_pic].h ||
                             (it.bi && (it.org_bi_off < 0 || (it.org_bi_off & 3) || !ck.have_side ||
                                        (long long)it.org_bi_off + (1ll << (it.log2_cuw + it.log2_cuh)) > ck.side_elems));
            if(bad) { key = 5; atomicOr(&bins[6], 1); }
        }
        int d = it.poc - it.ref_poc;
        d     = d < 0 ? -d : d;
        int dyn = (it.max_search_range * d + (it.gop_size >> 1)) / max(1, it.gop_size);
        dyn     = max(it.max_search_range >> 2, min(it.max_search_range, dyn));
        margin  = it.bi ? 6 : dyn + 2; // bi: window radius 5, +1 for the integer refinement pass (me_level 1)
    }
#pragma unroll
    for(int k = 0; k < 5; k++) { // warp-aggregated: one atomic per warp and bin
        const unsigned m = __ballot_sync(0xffffffffu, key == k);
        if(m == 0) continue;
        const int leader = __ffs(m) - 1;
        int       mx = key == k ? margin : 0;
#pragma unroll
        for(int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        int base = 0;
        if(lane == leader) {
            base = atomicAdd(&bins[k], __popc(m));
            if(k < 4) atomicMax(&bins[8 + k], mx);
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        if(key == k && k < 4) order[(size_t)k * n + base + __popc(m & ((1u << lane) - 1))] = i;
    }
}

// Coefficient read-back of xb200_residue with host buffers: only the planes that hold a non-zero level travel.  A warp per item
// reserves room for each non-zero plane in the compact buffer (atomic bump), records {destination in the caller's buffer, place in
// the compact buffer, size} and copies the plane.
struct PlaneRef {
    long long dst;          // element offset in the caller's coefficient buffer
    unsigned  src, elems;   // element offset in the compact buffer, plane size
};
__global__ void k_compact_planes(const xb200_residue_item *__restrict__ items, int n, const int16_t *__restrict__ coef,
                                 int16_t *__restrict__ compact, PlaneRef *__restrict__ list, unsigned long long *__restrict__ total,
                                 int *__restrict__ count)
{
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if(i >= n) return;
    const xb200_residue_item &it = items[i];
    const int ny = it.mc.w * it.mc.h, nc = ny >> 2;
    if(ny == 0) return; // empty slot
#pragma unroll
    for(int c = 0; c < 3; c++) {
        if(it.nnz[c] == 0) continue;
        const long long dst = it.out_off + (c == 0 ? 0 : (c == 1 ? ny : ny + nc));
        const int       sz = c ? nc : ny;
        unsigned        at = 0;
        if(lane == 0) {
            at = (unsigned)atomicAdd(total, (unsigned long long)sz);
            const int k = atomicAdd(count, 1);
            list[k].dst = dst; list[k].src = at; list[k].elems = (unsigned)sz;
        }
        at = __shfl_sync(0xffffffffu, at, 0);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(coef + dst);   // plane sizes and offsets are even
        uint32_t       *out = reinterpret_cast<uint32_t *>(compact + at);
        for(int e = lane; e < (sz >> 1); e += 32) out[e] = src[e];
    }
}

// the four size-binned grids of one operator call run concurrently on side streams
int fork_streams(xb200_ctx *c)
{
    CK(cudaEventRecord(c->ev_fork, c->stream));
    for(int i = 0; i < 4; i++) CK(cudaStreamWaitEvent(c->side[i], c->ev_fork, 0));
    return XB200_OK;
}
int join_streams(xb200_ctx *c)
{
    for(int i = 0; i < 4; i++) {
        CK(cudaEventRecord(c->ev_join[i], c->side[i]));
        CK(cudaStreamWaitEvent(c->stream, c->ev_join[i], 0));
    }
    return XB200_OK;
}

template <int L2>
int launch_me(xb200_ctx *c, xb200_me_item *d_items, const int32_t *order, int cnt, const int16_t *d_side, int margin)
{
    if(cnt == 0) return XB200_OK;
    const int W = 1 << L2, ext = W + 2 * margin + 7;
    const int cap = (align_up(ext, 8) + 8) * ext + 16;
    const size_t smem = me_smem_bytes(L2, cap);
    if(smem > 227 * 1024) return XB200_ERR_UNSUPPORTED;
    CK(cudaFuncSetAttribute(k_me<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_me<L2><<<(cnt + MeGeom<L2>::TEAMS - 1) / MeGeom<L2>::TEAMS, MeGeom<L2>::CTA, smem, c->side[L2 - 3]>>>(c->d_pics, d_items, order, cnt, d_side, c->sq, cap, c->d_err);
    c->launches++;
    CK(cudaGetLastError());
    return XB200_OK;
}

template <int L2, bool USE_TC>
int launch_residue2_v(xb200_ctx *c, xb200_residue_item *d_items, const int32_t *order, int cnt, const xb200_rates *d_rates,
                      int16_t *d_coef, int16_t *d_rec, int16_t *d_pred)
{
    using Cf = Res2Cfg<L2>;
    constexpr int smem = USE_TC ? Cf::SMEM : Cf::SMEM_INT;
    int &blocks_per_sm = c->res2_blocks[L2 - 3][USE_TC ? 1 : 0];   // per context (= per device), not per process
    const int sms = c->sms;
    if(!blocks_per_sm) {
        CK(cudaFuncSetAttribute(k_residue2<L2, USE_TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_residue2<L2, USE_TC>, Cf::CTA, smem));
        if(blocks_per_sm < 1) blocks_per_sm = 1;
    }
    const int want = (cnt + Cf::TEAMS - 1) / Cf::TEAMS, grid = want < sms * blocks_per_sm ? want : sms * blocks_per_sm;
    k_residue2<L2, USE_TC><<<grid, Cf::CTA, smem, c->side[L2 - 3]>>>(c->d_pics, d_items, order, cnt, d_rates, d_coef, d_rec, c->d_tm64, c->sq, d_pred);
    c->launches++;
    CK(cudaGetLastError());
    return XB200_OK;
}
template <int L2>
int launch_residue2(xb200_ctx *c, xb200_residue_item *d_items, const int32_t *order, int cnt, const xb200_rates *d_rates,
                    int16_t *d_coef, int16_t *d_rec, int16_t *d_pred = nullptr)
{
    if(cnt == 0) return XB200_OK;
    if(Res2Cfg<L2>::TC && c->sq.tc_dct && c->sq.bd <= 10) return launch_residue2_v<L2, true>(c, d_items, order, cnt, d_rates, d_coef, d_rec, d_pred);
    return launch_residue2_v<L2, false>(c, d_items, order, cnt, d_rates, d_coef, d_rec, d_pred);
}

template <int L2, int TEAMS>
int launch_analyze_v(xb200_ctx *c, xb200_cu_item *d_items, const int32_t *order, int cnt, const xb200_rates *d_rates, const xb200_sbac *d_st_in,
                     xb200_sbac *d_st_out, int16_t *d_coef, int16_t *d_rec, int cap)
{
    using Cf = CuCfg<L2, TEAMS>;
    const size_t smem = Cf::smem_bytes(cap);
    int blocks_per_sm = 0, sms = 0;
    CK(cudaFuncSetAttribute(k_analyze_cu<L2, TEAMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, k_analyze_cu<L2, TEAMS>, Cf::CTA, smem));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    if(blocks_per_sm < 1) blocks_per_sm = 1;
    const int want = (cnt + TEAMS - 1) / TEAMS, grid = want < sms * blocks_per_sm ? want : sms * blocks_per_sm;
    int r = ensure(c->b_scr[L2 - 3], (size_t)grid * TEAMS * Cf::SCRATCH * sizeof(int16_t));
    if(r) return r;
    k_analyze_cu<L2, TEAMS><<<grid, Cf::CTA, smem, c->side[L2 - 3]>>>(c->d_pics, d_items, order, cnt, d_rates, d_st_in, d_st_out, d_coef, d_rec,
                                                                     static_cast<int16_t *>(c->b_scr[L2 - 3].p), c->d_tm64, c->sq, cap, c->d_err);
    c->launches++;
    CK(cudaGetLastError());
    return XB200_OK;
}
template <int L2>
int launch_analyze(xb200_ctx *c, xb200_cu_item *d_items, const int32_t *order, int cnt, const xb200_rates *d_rates, const xb200_sbac *d_st_in,
                   xb200_sbac *d_st_out, int16_t *d_coef, int16_t *d_rec, int margin)
{
    if(cnt == 0) return XB200_OK;
    const int    W = 1 << L2, ext = W + 2 * margin + 7;
    const int    cap = (align_up(ext, 8) + 8) * ext + 16;
    const size_t lim = 227 * 1024;
    if constexpr(L2 <= 4) { // a warp per CU: as many CUs per CTA as the staged search windows leave room for
        if(CuCfg<L2, 4>::smem_bytes(cap) <= lim) return launch_analyze_v<L2, 4>(c, d_items, order, cnt, d_rates, d_st_in, d_st_out, d_coef, d_rec, cap);
        if(CuCfg<L2, 2>::smem_bytes(cap) <= lim) return launch_analyze_v<L2, 2>(c, d_items, order, cnt, d_rates, d_st_in, d_st_out, d_coef, d_rec, cap);
    }
    if(CuCfg<L2, 1>::smem_bytes(cap) <= lim) return launch_analyze_v<L2, 1>(c, d_items, order, cnt, d_rates, d_st_in, d_st_out, d_coef, d_rec, cap);
    return XB200_ERR_UNSUPPORTED;
}

} // namespace

int xb200_sync_pics(xb200_ctx *c) { return sync_pics(c); }

extern "C" {

const char *xb200_version(void) { return "xeve_b200 0.1 (sm_100a)"; }

int xb200_create(xb200_ctx **out, int device, const xb200_seq *seq)
{
    if(!out || !seq) return XB200_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if(seq->w <= 0 || seq->h <= 0 || (seq->w & 7) || (seq->h & 7) || seq->bit_depth < 8 || seq->bit_depth > 14)
        return XB200_ERR_INVALID_ARGUMENT;
    int ndev = 0;
    if(cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        fprintf(stderr, "xeve_b200: no usable CUDA device (this library has no CPU path)\n");
        return XB200_ERR_UNSUPPORTED;
    }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if(prop.major != 10) {
        fprintf(stderr, "xeve_b200: device %d is sm_%d%d; this build targets sm_100a only\n", device, prop.major, prop.minor);
        return XB200_ERR_UNSUPPORTED;
    }
    CK(cudaSetDevice(device));
    xb200_ctx *c = new xb200_ctx();
    c->device = device;
    c->sms    = prop.multiProcessorCount;
    c->seq    = *seq;
    c->sq.w = seq->w; c->sq.h = seq->h; c->sq.bd = seq->bit_depth;
    c->sq.me_level = seq->me_level; c->sq.hpel_cnt = seq->hpel_cnt; c->sq.qpel_cnt = seq->qpel_cnt;
    c->sq.me_complexity = seq->me_complexity;
    for(int i = 0; i < 2; i++) { c->sq.min_clip[i] = seq->min_clip[i]; c->sq.max_clip[i] = seq->max_clip[i]; }
    c->sq.rdoq = seq->rdoq; c->sq.merge_num = seq->merge_num; c->sq.gop_size = seq->gop_size;
    {   // opt-in: tensor-core transform stages inside xb200_residue (bit-identical to the integer stages)
        const char *e = getenv("XB200_TC_DCT");
        c->sq.tc_dct = (e && e[0] == '1') ? 1 : 0;
    }
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    for(int i = 0; i < 4; i++) {
        CK(cudaStreamCreateWithFlags(&c->side[i], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
    }
    // constant tables: device-wide state, uploaded by the first context of a device only -- a later context may be created while the
    // chain server of another one is alive, and then nothing here may wait for the device (xb200_chain.cu: chain_init)
    static std::mutex once_mu;
    static bool       once_done[64] = {};
    std::lock_guard<std::mutex> once_lock(once_mu);
    const bool first = !once_done[device & 63];
    {
        static int8_t tm[64 * 64];
        xb200_gen_tm64(tm);
        if(first) CK(cudaMemcpyToSymbol(c_tm64, tm, sizeof(tm)));
        CK(cudaMalloc(&c->d_tm64, sizeof(tm)));
        CK(cudaMemcpyAsync(c->d_tm64, tm, sizeof(tm), cudaMemcpyHostToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        const int16_t l[4][8] = XB200_MC_L_TAPS;
        const int16_t ch[8][4] = XB200_MC_C_TAPS;
        const int32_t qs[6] = XB200_QUANT_SCALE, dq[6] = XB200_DEQUANT_SCALE;
        static int64_t es[7][6][7];
        for(int b = 0; b < 7; b++)
            for(int q = 0; q < 6; q++)
                for(int l2 = 0; l2 < 7; l2++) es[b][q][l2] = xb200_err_scale(q, l2, b + 8);
        if(first) {
            CK(cudaMemcpyToSymbol(c_mc_l, l, sizeof(l)));
            CK(cudaMemcpyToSymbol(c_mc_c, ch, sizeof(ch)));
            CK(cudaMemcpyToSymbol(c_quant_scale, qs, sizeof(qs)));
            CK(cudaMemcpyToSymbol(c_dequant_scale, dq, sizeof(dq)));
            CK(cudaMemcpyToSymbol(c_err_scale, es, sizeof(es)));
        }
    }
    if(first) {   // rate estimation tables: zig-zag scans (closed form) and xeve_init_bits_est (src_base/xeve_mode.c:304-313, host libm)
        static uint16_t scan[16 + 64 + 256 + 1024 + 4096];
        static int32_t  eb[1024];
        int             off = 0;
        for(int l2 = 2; l2 <= 6; l2++) { xb200_gen_scan(scan + off, l2, l2); off += 1 << (2 * l2); }
        for(int i = 0; i < 1024; i++) {
            const double p = (512 * (i + 0.5)) / 1024;
            eb[i] = (int32_t)(-32768 * (log(p) / log(2.0) - 9));
        }
        CK(cudaMemcpyToSymbol(g_scan, scan, sizeof(scan)));
        CK(cudaMemcpyToSymbol(g_entropy_bits, eb, sizeof(eb)));
    }
    CK(cudaMalloc(&c->d_err, sizeof(int)));
    CK(cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMalloc(&c->d_bins, sizeof(int) * 16));
    if(first) {
        CK(cudaFuncSetAttribute(k_mc, cudaFuncAttributeMaxDynamicSharedMemorySize, MC_SMEM_BYTES));
        CK(cudaFuncSetAttribute(k_bi_org, cudaFuncAttributeMaxDynamicSharedMemorySize, MC_SMEM_BYTES));
        CK(cudaFuncSetAttribute(k_tq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TqSmem)));
        CK(cudaFuncSetAttribute(k_itdq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TqSmem)));
    }
    once_done[device & 63] = true;
    *out = c;
    return XB200_OK;
}

void xb200_destroy(xb200_ctx *c)
{
    if(!c) return;
    cudaSetDevice(c->device);
    const bool had_chain = c->chain != nullptr;
    const int  device = c->device;
    xb200_chain_free(c);          // returns with the device's worker grid gone and held off (cudaFree waits for the device)
    cudaStreamSynchronize(c->stream);
    for(auto &p : c->pics)
        for(int k = 0; k < 3; k++)
            if(p.buf[k]) cudaFree(p.buf[k]);
    for(void *g : c->garbage) cudaFree(g);
    for(DevBuf *b : {&c->b_items, &c->b_side, &c->b_aux0, &c->b_aux1, &c->b_aux2, &c->b_order, &c->b_stage, &c->b_df, &c->b_in_items,
                     &c->b_in_rates, &c->b_in_st0, &c->b_in_st1, &c->b_in_side, &c->b_in_coef, &c->b_in_rec, &c->b_in_order, &c->b_scr[0],
                     &c->b_scr[1], &c->b_scr[2], &c->b_scr[3], &c->b_st0, &c->b_st1, &c->b_cu_items, &c->b_cu_rates, &c->b_cu_state,
                     &c->b_cu_me, &c->b_cu_res, &c->b_cu_mc, &c->b_cu_cur, &c->b_cu_off, &c->b_cu_side, &c->b_cu_order, &c->b_cu_coef,
                     &c->b_cu_rec, &c->b_cu_nzr, &c->b_cu_nzl, &c->b_cu_meta})
        if(b->p) cudaFree(b->p);
    for(DevBuf *b : {&c->b_compact, &c->b_coff, &c->b_ats})
        if(b->p) cudaFree(b->p);
    if(c->h_pin) cudaFreeHost(c->h_pin);
    if(c->d_pics) cudaFree(c->d_pics);
    if(c->d_tm64) cudaFree(c->d_tm64);
    if(c->d_err) cudaFree(c->d_err);
    if(c->d_bins) cudaFree(c->d_bins);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaEventDestroy(c->ev_fork);
    for(int i = 0; i < 4; i++) { cudaStreamDestroy(c->side[i]); cudaEventDestroy(c->ev_join[i]); }
    cudaStreamDestroy(c->stream);
    delete c;
    if(had_chain) xb200_chain_drain_end(device);
}

int64_t xb200_launch_count(const xb200_ctx *c) { return c ? c->launches : 0; }
double  xb200_last_kernel_ms(const xb200_ctx *c) { return c ? c->last_ms : 0.0; }

// ---- pictures ---------------------------------------------------------------------------------------------
int xb200_pic_create(xb200_ctx *c, int padded, int32_t *handle)
{
    if(!c || !handle) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    int idx = -1;
    for(size_t i = 0; i < c->pics.size(); i++)    // prefer a free slot that already holds buffers of this kind
        if(!c->pics[i].used && c->pics[i].buf[0] && c->pics[i].padded == (padded != 0)) { idx = (int)i; break; }
    for(size_t i = 0; i < c->pics.size() && idx < 0; i++)
        if(!c->pics[i].used) { idx = (int)i; break; }
    if(idx < 0) { c->pics.emplace_back(); idx = (int)c->pics.size() - 1; }
    Pic &p = c->pics[idx];
    // a destroyed picture keeps its buffers (cudaFree waits for the device, and the chain server's worker grid lives as long as any
    // context has work): a slot of the same kind is reused as it is, zeroed like a new one
    const bool reuse = p.buf[0] && p.padded == (padded != 0);
    if(!reuse) {
        for(int k = 0; k < 3; k++) if(p.buf[k]) c->garbage.push_back(p.buf[k]);   // other kind: freed with the context
        p = Pic();
        p.padded = padded != 0;
    }
    for(int k = 0; k < 3; k++) {
        p.w[k]   = k ? c->seq.w / 2 : c->seq.w;
        p.h[k]   = k ? c->seq.h / 2 : c->seq.h;
        p.pad[k] = padded ? (k ? XB200_PAD_C : XB200_PAD_L) : 0;
        p.s[k]   = align_up(p.w[k] + 2 * p.pad[k], 64);
        const size_t elems = (size_t)p.s[k] * (p.h[k] + 2 * p.pad[k] + 1) + 64;
        if(!reuse) CK(cudaMalloc(&p.buf[k], elems * sizeof(int16_t)));
        CK(cudaMemsetAsync(p.buf[k], 0, elems * sizeof(int16_t), c->stream));
    }
    p.used = true;
    c->pics_dirty = true;
    *handle = idx;
    return XB200_OK;
}

int xb200_pic_destroy(xb200_ctx *c, int32_t handle)
{
    if(!c || !pic_ok(c, handle)) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    Pic &p = c->pics[handle];
    p.used = false;               // the buffers stay with the slot (see xb200_pic_create)
    c->pics_dirty = true;
    return XB200_OK;
}

// Process-level CUDA settings this library depends on, applied when it is loaded BEFORE the CUDA runtime initialises (an application
// linked against it; a host that has already initialised CUDA, like the Python bench, sets them itself):
//   CUDA_MODULE_LOADING=EAGER       a kernel's first launch must never have to load code while the chain server is alive (see below)
//   CUDA_DEVICE_MAX_CONNECTIONS=32  hardware work queues: short kernels of many streams next to long-running ones
// Never overrides what the user has set.
__attribute__((constructor)) static void xb200_process_env()
{
    setenv("CUDA_MODULE_LOADING", "EAGER", 0);
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
}

// Load (not launch) the short kernels that run NEXT TO the chain server.  With lazy module loading (the CUDA 12 default) the first launch
// of a kernel loads its code, and that can synchronise the device -- which never happens while the long-lived worker grid waits for
// work: the first loop filter or upload conversion after the server had started dead-locked.  cudaFuncGetAttributes forces the load.
extern "C++" int xb200_preload_api_kernels()
{
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k_pad3));
    CK(cudaFuncGetAttributes(&fa, k_convert<uint8_t>));
    CK(cudaFuncGetAttributes(&fa, k_convert<uint16_t>));
    return XB200_OK;
}

extern "C++" int xb200_pad_planes(xb200_ctx *c, Pic &p, cudaStream_t stream)
{
    if(!p.padded) return XB200_OK;
    if(!stream) stream = c->stream;
    PadArgs a;
    int     total = 0;
    for(int k = 0; k < 3; k++) {
        a.act[k] = p.buf[k] + (size_t)p.pad[k] * p.s[k] + p.pad[k];
        a.s[k] = p.s[k]; a.w[k] = p.w[k]; a.h[k] = p.h[k]; a.pad[k] = p.pad[k];
        a.first[k] = total;
        total += pad_plane_groups(p.w[k], p.h[k], p.pad[k]);
    }
    a.first[3] = total;
    k_pad3<<<(total + 255) / 256, 256, 0, stream>>>(a);
    c->launches++;
    CK(cudaGetLastError());
    return XB200_OK;
}

int xb200_pic_upload(xb200_ctx *c, int32_t handle, const void *const planes[3], const int32_t stride_bytes[3], int in_bit_depth,
                     int mem)
{
    if(!c || !pic_ok(c, handle) || !planes || !stride_bytes || in_bit_depth < 8 || in_bit_depth > c->seq.bit_depth)
        return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    Pic      &p = c->pics[handle];
    const int bps = in_bit_depth > 8 ? 2 : 1, shift = c->seq.bit_depth - in_bit_depth;
    CK(cudaEventRecord(c->ev0, c->stream));
    size_t off[3], total = 0;
    for(int k = 0; k < 3; k++) { off[k] = total; total += (size_t)align_up(p.w[k] * bps, 256) * p.h[k]; }
    if(mem == XB200_MEM_HOST) {
        int r = ensure(c->b_stage, total);
        if(r) return r;
    }
    for(int k = 0; k < 3; k++) {
        const uint8_t *src;
        int            sstride;
        if(mem == XB200_MEM_HOST) {
            sstride = align_up(p.w[k] * bps, 256);
            uint8_t *d = static_cast<uint8_t *>(c->b_stage.p) + off[k];
            CK(cudaMemcpy2DAsync(d, sstride, planes[k], stride_bytes[k], (size_t)p.w[k] * bps, p.h[k], cudaMemcpyHostToDevice,
                                 c->stream));
            src = d;
        }
        else { src = static_cast<const uint8_t *>(planes[k]); sstride = stride_bytes[k]; }
        int16_t *act = p.buf[k] + (size_t)p.pad[k] * p.s[k] + p.pad[k];
        dim3     grid((p.w[k] + 255) / 256, p.h[k]);
        if(bps == 1) k_convert<uint8_t><<<grid, 256, 0, c->stream>>>(src, sstride, act, p.s[k], p.w[k], p.h[k], shift);
        else k_convert<uint16_t><<<grid, 256, 0, c->stream>>>(reinterpret_cast<const uint16_t *>(src), sstride / 2, act, p.s[k], p.w[k], p.h[k], shift);
        c->launches++;
    }
    CK(cudaGetLastError());
    int r = xb200_pad_planes(c, p);
    if(r) return r;
    return finish(c);
}

int xb200_pic_upload_s16(xb200_ctx *c, int32_t handle, const int16_t *const planes[3], const int32_t stride_elems[3], int mem)
{
    if(!c || !pic_ok(c, handle) || !planes || !stride_elems) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    Pic &p = c->pics[handle];
    CK(cudaEventRecord(c->ev0, c->stream));
    for(int k = 0; k < 3; k++) {
        int16_t *act = p.buf[k] + (size_t)p.pad[k] * p.s[k] + p.pad[k];
        CK(cudaMemcpy2DAsync(act, (size_t)p.s[k] * 2, planes[k], (size_t)stride_elems[k] * 2, (size_t)p.w[k] * 2, p.h[k],
                             mem == XB200_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, c->stream));
    }
    int r = xb200_pad_planes(c, p);
    if(r) return r;
    return finish(c);
}

int xb200_pic_download(xb200_ctx *c, int32_t handle, int with_padding, int16_t *const planes[3], const int32_t stride_elems[3])
{
    if(!c || !pic_ok(c, handle) || !planes || !stride_elems) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    Pic &p = c->pics[handle];
    for(int k = 0; k < 3; k++) {
        const int      pad = with_padding ? p.pad[k] : 0;
        const int16_t *src = p.buf[k] + (size_t)(p.pad[k] - pad) * p.s[k] + (p.pad[k] - pad);
        CK(cudaMemcpy2DAsync(planes[k], (size_t)stride_elems[k] * 2, src, (size_t)p.s[k] * 2, (size_t)(p.w[k] + 2 * pad) * 2,
                             p.h[k] + 2 * pad, cudaMemcpyDeviceToHost, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    return XB200_OK;
}

// ---- probes -----------------------------------------------------------------------------------------------
// ---- whole inter mode decision of a list of CUs ---------------------------------------------------------------------
static int analyze_pipeline(xb200_ctx *c, xb200_cu_item *items, int64_t n, const xb200_rates *rates, int64_t n_rates, xb200_sbac *states,
                            int64_t n_states, int16_t *coef, int16_t *rec, int64_t elems, const int32_t *order, const int *cnt);

int xb200_analyze_cu(xb200_ctx *c, xb200_cu_item *items, int64_t n, const xb200_rates *rates, int64_t n_rates, xb200_sbac *states,
                     int64_t n_states, int16_t *coef, int16_t *rec, int64_t elems)
{
    if(!c || n < 0 || n_rates < 0 || n_states < 0 || elems < 0 || (n && (!items || !rates || !states || !coef)) || n > (1 << 26))
        return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    if(n == 0) return XB200_OK;
    if(c->seq.merge_num < 1 || c->seq.merge_num > 4 || c->seq.gop_size < 1) return XB200_ERR_INVALID_ARGUMENT;
    std::vector<int32_t> order((size_t)4 * n);
    int                  cnt[4] = {0, 0, 0, 0}, margin[4] = {0, 0, 0, 0};
    for(int64_t i = 0; i < n; i++) {
        const xb200_cu_item &it = items[i];
        if(it.log2_cuw != it.log2_cuh || it.log2_cuw < 3 || it.log2_cuw > 6) return XB200_ERR_UNSUPPORTED;
        const int w = 1 << it.log2_cuw, k = it.log2_cuw - 3;
        if(!pic_ok(c, it.cur_pic) || it.slice_type > 1 || it.x < 0 || it.y < 0 || it.x + w > c->seq.w || it.y + w > c->seq.h ||
           it.rate_idx < 0 || it.rate_idx >= n_rates || it.state_in < 0 || it.state_in >= n_states || it.state_out >= n_states ||
           it.out_off < 0 || it.out_off + w * w * 3 / 2 > elems || it.max_search_range < 1 || it.max_search_range > 256)
            return XB200_ERR_INVALID_ARGUMENT;
        for(int l = 0; l < (it.slice_type == 0 ? 2 : 1); l++) {
            if(it.num_refp[l] < 1 || it.num_refp[l] > XB200_MAX_REFP) return XB200_ERR_UNSUPPORTED;
            for(int q = 0; q < it.num_refp[l]; q++) {
                if(!pic_ok(c, it.ref_pic[l][q]) || !c->pics[it.ref_pic[l][q]].padded) return XB200_ERR_INVALID_ARGUMENT;
                int d = it.poc - it.ref_poc[l][q];
                d = d < 0 ? -d : d;
                int dyn = (it.max_search_range * d + (c->seq.gop_size >> 1)) / c->seq.gop_size;
                dyn = dyn < (it.max_search_range >> 2) ? (it.max_search_range >> 2) : (dyn > it.max_search_range ? it.max_search_range : dyn);
                if(dyn + 2 > margin[k]) margin[k] = dyn + 2;
            }
            for(int q = 0; q < 4; q++)
                if(it.refi_pred[l][q] >= (int)it.num_refp[l]) return XB200_ERR_INVALID_ARGUMENT;
        }
        order[(size_t)k * n + cnt[k]++] = (int32_t)i;
    }
    int r = sync_pics(c);
    if(r) return r;
    {   // default: the pipeline of frame-wide grids; XB200_ANALYZE=fused selects the one-team-per-CU kernel (same results)
        const char *e = getenv("XB200_ANALYZE");
        if(!(e && e[0] == 'f')) return analyze_pipeline(c, items, n, rates, n_rates, states, n_states, coef, rec, elems, order.data(), cnt);
    }
    xb200_cu_item *d_items;
    xb200_rates   *d_rates;
    xb200_sbac    *d_in, *d_out;
    int32_t       *d_order;
    if((r = to_dev(c, c->b_items, items, (size_t)n, XB200_MEM_HOST, &d_items))) return r;
    if((r = to_dev(c, c->b_aux0, rates, (size_t)n_rates, XB200_MEM_HOST, &d_rates))) return r;
    if((r = to_dev(c, c->b_st0, states, (size_t)n_states, XB200_MEM_HOST, &d_in))) return r;
    if((r = to_dev(c, c->b_st1, states, (size_t)n_states, XB200_MEM_HOST, &d_out))) return r;
    if((r = to_dev(c, c->b_order, order.data(), order.size(), XB200_MEM_HOST, &d_order))) return r;
    if((r = ensure(c->b_aux1, (size_t)elems * 2 + 64))) return r;
    int16_t *d_coef = static_cast<int16_t *>(c->b_aux1.p), *d_rec = nullptr;
    if(rec) {
        if((r = ensure(c->b_aux2, (size_t)elems * 2 + 64))) return r;
        d_rec = static_cast<int16_t *>(c->b_aux2.p);
    }
    CK(cudaStreamSynchronize(c->stream)); // `order` is pageable host memory
    CK(cudaEventRecord(c->ev0, c->stream));
    if((r = fork_streams(c))) return r;
    if((r = launch_analyze<6>(c, d_items, d_order + 3 * n, cnt[3], d_rates, d_in, d_out, d_coef, d_rec, margin[3]))) return r;
    if((r = launch_analyze<5>(c, d_items, d_order + 2 * n, cnt[2], d_rates, d_in, d_out, d_coef, d_rec, margin[2]))) return r;
    if((r = launch_analyze<4>(c, d_items, d_order + 1 * n, cnt[1], d_rates, d_in, d_out, d_coef, d_rec, margin[1]))) return r;
    if((r = launch_analyze<3>(c, d_items, d_order + 0 * n, cnt[0], d_rates, d_in, d_out, d_coef, d_rec, margin[0]))) return r;
    if((r = join_streams(c))) return r;
    CK(cudaEventRecord(c->ev1, c->stream));
    if((r = to_host(c, items, d_items, (size_t)n, XB200_MEM_HOST))) return r;
    if((r = to_host(c, states, d_out, (size_t)n_states, XB200_MEM_HOST))) return r;
    if((r = to_host(c, coef, d_coef, (size_t)elems, XB200_MEM_HOST))) return r;
    if(rec && (r = to_host(c, rec, d_rec, (size_t)elems, XB200_MEM_HOST))) return r;
    int err = 0;
    CK(cudaMemcpyAsync(&err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_ms = ms;
    CK(cudaGetLastError());
    if(err) { cudaMemset(c->d_err, 0, sizeof(int)); return XB200_ERR_UNEXPECTED; }
    return XB200_OK;
}

} // extern "C"
template <typename OutT, typename K>
static int run_probe(xb200_ctx *c, const xb200_blk_item *items, int64_t n, OutT *out, int mem, K kernel)
{
    if(!c || (n && (!items || !out)) || n < 0) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    if(mem == XB200_MEM_HOST)
        for(int64_t i = 0; i < n; i++)
            if(!pic_ok(c, items[i].pic1) || !pic_ok(c, items[i].pic2) || items[i].plane1 > 2 || items[i].plane2 > 2)
                return XB200_ERR_INVALID_ARGUMENT;
    int r = sync_pics(c);
    if(r) return r;
    xb200_blk_item *d_items;
    OutT           *d_out = out;
    if((r = to_dev(c, c->b_items, items, (size_t)n, mem, &d_items))) return r;
    if(mem == XB200_MEM_HOST) {
        if((r = ensure(c->b_aux0, (size_t)n * sizeof(OutT) + 64))) return r;
        d_out = static_cast<OutT *>(c->b_aux0.p);
    }
    CK(cudaEventRecord(c->ev0, c->stream));
    if(n) {
        kernel<<<(unsigned)((n + 3) / 4), 128, 0, c->stream>>>(c->d_pics, d_items, n, d_out, c->sq.bd);
        c->launches++;
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    if((r = to_host(c, out, d_out, (size_t)n, mem))) return r;
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_ms = ms;
    CK(cudaGetLastError());
    return XB200_OK;
}
extern "C" {
int xb200_sad(xb200_ctx *c, const xb200_blk_item *items, int64_t n, int32_t *out, int mem) { return run_probe(c, items, n, out, mem, k_sad); }
int xb200_ssd(xb200_ctx *c, const xb200_blk_item *items, int64_t n, int64_t *out, int mem) { return run_probe(c, items, n, out, mem, k_ssd); }
int xb200_satd(xb200_ctx *c, const xb200_blk_item *items, int64_t n, int32_t *out, int mem) { return run_probe(c, items, n, out, mem, k_satd); }

} // extern "C"
template <int LN> static int run_dct_tc(xb200_ctx *c, const int16_t *d_in, int16_t *d_out, int n)
{
    const size_t  smem = tc_probe_smem<LN>();
    CK(cudaFuncSetAttribute(k_dct_tc<LN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = n < 296 ? n : 296;
    k_dct_tc<LN><<<grid, 128, smem, c->stream>>>(d_in, d_out, n, c->d_tm64, c->sq.bd);
    c->launches++;
    CK(cudaGetLastError());
    return XB200_OK;
}

extern "C" {
int xb200_fwd_dct_tc(xb200_ctx *c, const int16_t *in, int16_t *out, int64_t n, int log2n)
{
    if(!c || n < 0 || (n && (!in || !out)) || n > (1 << 24)) return XB200_ERR_INVALID_ARGUMENT;
    if(log2n != 5 && log2n != 6) return XB200_ERR_UNSUPPORTED;
    CK(cudaSetDevice(c->device));
    if(n == 0) return XB200_OK;
    const size_t elems = (size_t)n << (2 * log2n);
    int r;
    int16_t *d_in;
    if((r = to_dev(c, c->b_aux0, in, elems, XB200_MEM_HOST, &d_in))) return r;
    if((r = ensure(c->b_aux1, elems * 2 + 64))) return r;
    int16_t *d_out = static_cast<int16_t *>(c->b_aux1.p);
    CK(cudaEventRecord(c->ev0, c->stream));
    r = log2n == 6 ? run_dct_tc<6>(c, d_in, d_out, (int)n) : run_dct_tc<5>(c, d_in, d_out, (int)n);
    if(r) return r;
    CK(cudaEventRecord(c->ev1, c->stream));
    if((r = to_host(c, out, d_out, elems, XB200_MEM_HOST))) return r;
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_ms = ms;
    CK(cudaGetLastError());
    return XB200_OK;
}

int xb200_mvp(xb200_ctx *c, xb200_mvp_item *items, int64_t n, const xb200_mvp_pic *pic, const uint32_t *map_scu, const int16_t *map_mv,
              const int16_t *col_mv0, const int16_t *col_mv1)
{
    if(!c || n < 0 || (n && (!items || !pic || !map_scu || !map_mv || !col_mv0 || !col_mv1)) || n > (1 << 28)) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    if(n == 0) return XB200_OK;
    if(pic->w_scu <= 0 || pic->h_scu <= 0) return XB200_ERR_INVALID_ARGUMENT;
    const size_t f = (size_t)pic->w_scu * pic->h_scu;
    for(int64_t i = 0; i < n; i++) {
        const xb200_mvp_item &it = items[i];
        const int sw = (1 << it.log2_cuw) >> 2, sh = (1 << it.log2_cuh) >> 2;
        if(it.x_scu < 0 || it.y_scu < 0 || it.log2_cuw < 2 || it.log2_cuh < 2 || it.log2_cuw > 7 || it.log2_cuh > 7 || it.lidx > 1 ||
           it.x_scu + sw > pic->w_scu || it.y_scu + sh > pic->h_scu)
            return XB200_ERR_INVALID_ARGUMENT;
    }
    int r;
    xb200_mvp_item *d_items;
    uint32_t       *d_scu;
    int16_t        *d_mv, *d_c0, *d_c1;
    if((r = to_dev(c, c->b_items, items, (size_t)n, XB200_MEM_HOST, &d_items))) return r;
    if((r = to_dev(c, c->b_aux0, map_scu, f, XB200_MEM_HOST, &d_scu))) return r;
    if((r = to_dev(c, c->b_aux1, map_mv, f * 4, XB200_MEM_HOST, &d_mv))) return r;
    if((r = to_dev(c, c->b_aux2, col_mv0, f * 4, XB200_MEM_HOST, &d_c0))) return r;
    if((r = to_dev(c, c->b_side, col_mv1, f * 4, XB200_MEM_HOST, &d_c1))) return r;
    CK(cudaEventRecord(c->ev0, c->stream));
    k_mvp<<<(unsigned)((n + 127) / 128), 128, 0, c->stream>>>(d_items, n, *pic, d_scu, d_mv, d_c0, d_c1);
    c->launches++;
    CK(cudaEventRecord(c->ev1, c->stream));
    if((r = to_host(c, items, d_items, (size_t)n, XB200_MEM_HOST))) return r;
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_ms = ms;
    CK(cudaGetLastError());
    return XB200_OK;
}

// ---- rate estimation ----------------------------------------------------------------------------------------
int xb200_rdo_bits(xb200_ctx *c, xb200_bits_item *items, int64_t n, xb200_sbac *states, int64_t n_states, const int16_t *coef,
                   int64_t coef_elems)
{
    if(!c || n < 0 || n_states < 0 || coef_elems < 0 || (n && (!items || !states)) || n > (1 << 28)) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    if(n == 0) return XB200_OK;
    for(int64_t i = 0; i < n; i++) {
        const xb200_bits_item &it = items[i];
        if(it.kind > 3 || it.slice_type > 2 || it.state_in < 0 || it.state_in >= n_states || it.state_out >= n_states || it.ch > 2 ||
           it.mvp_idx[0] > 3 || it.mvp_idx[1] > 3)
            return XB200_ERR_INVALID_ARGUMENT;
        if(it.kind == 1 || it.kind == 3) {
            const int64_t ny = (int64_t)1 << (it.log2_cuw + it.log2_cuh);
            if(it.log2_cuw != it.log2_cuh || it.log2_cuw < 3 || it.log2_cuw > 6 || !coef || it.coef_off < 0 ||
               it.coef_off + ny + (ny >> 1) > coef_elems)
                return XB200_ERR_INVALID_ARGUMENT;
        }
    }
    int r;
    xb200_bits_item *d_items;
    xb200_sbac      *d_in, *d_out;
    int16_t         *d_coef;
    if((r = to_dev(c, c->b_items, items, (size_t)n, XB200_MEM_HOST, &d_items))) return r;
    if((r = to_dev(c, c->b_aux0, states, (size_t)n_states, XB200_MEM_HOST, &d_in))) return r;
    if((r = to_dev(c, c->b_aux1, states, (size_t)n_states, XB200_MEM_HOST, &d_out))) return r;
    if((r = to_dev(c, c->b_side, coef, (size_t)coef_elems, XB200_MEM_HOST, &d_coef))) return r;
    CK(cudaEventRecord(c->ev0, c->stream));
    const long long blocks = (n + RATE_WARPS - 1) / RATE_WARPS;
    k_rdo_bits<<<(unsigned)(blocks < c->sms * 16 ? blocks : c->sms * 16), RATE_WARPS * 32, 0, c->stream>>>(d_items, n, d_in, d_out, d_coef);
    c->launches++;
    if((r = to_host(c, items, d_items, (size_t)n, XB200_MEM_HOST))) return r;
    if((r = to_host(c, states, d_out, (size_t)n_states, XB200_MEM_HOST))) return r;
    return finish(c);
}

int xb200_rdoq_rates(xb200_ctx *c, const xb200_sbac *states, int64_t n, xb200_rates *rates)
{
    if(!c || n < 0 || (n && (!states || !rates)) || n > (1 << 24)) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    if(n == 0) return XB200_OK;
    int r;
    xb200_sbac *d_st;
    if((r = to_dev(c, c->b_aux0, states, (size_t)n, XB200_MEM_HOST, &d_st))) return r;
    if((r = ensure(c->b_aux1, (size_t)n * sizeof(xb200_rates)))) return r;
    xb200_rates *d_rt = static_cast<xb200_rates *>(c->b_aux1.p);
    CK(cudaEventRecord(c->ev0, c->stream));
    CK(cudaMemsetAsync(d_rt, 0, (size_t)n * sizeof(xb200_rates), c->stream));
    k_rdoq_rates<<<(unsigned)((n * 64 + 127) / 128), 128, 0, c->stream>>>(d_st, n, d_rt);
    c->launches++;
    if((r = to_host(c, rates, d_rt, (size_t)n, XB200_MEM_HOST))) return r;
    return finish(c);
}

// ---- motion search ------------------------------------------------------------------------------------------
int xb200_me(xb200_ctx *c, xb200_me_item *items, int64_t n, const int16_t *side, int64_t side_elems, int mem)
{
    if(!c || n < 0 || (n && !items) || n > (1 << 28)) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    if(c->sq.me_complexity > 1) return XB200_ERR_UNSUPPORTED; // me_raster (placebo) is not offloaded
    if(n == 0) return XB200_OK;
    int r = sync_pics(c);
    if(r) return r;
    xb200_me_item *d_items;
    int16_t       *d_side;
    if((r = to_dev(c, c->b_items, items, (size_t)n, mem, &d_items))) return r;
    if((r = to_dev(c, c->b_side, side, (size_t)side_elems, mem, &d_side))) return r;
    if((r = ensure(c->b_order, sizeof(int32_t) * 4 * (size_t)n))) return r;
    int32_t *order = static_cast<int32_t *>(c->b_order.p);
    CK(cudaEventRecord(c->ev0, c->stream));
    CK(cudaMemsetAsync(c->d_bins, 0, sizeof(int) * 16, c->stream));
    // host-buffer calls: the records are validated by the binning kernel (invalid picture handles, org_bi ranges -> bins[6])
    const BinCheck ck = {c->d_pics, (int)c->pics.size(), mem == XB200_MEM_HOST, side != nullptr, 0, (long long)side_elems, 0};
    k_me_bin<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_items, (int)n, order, c->d_bins, ck);
    c->launches++;
    int bins[16];
    CK(cudaMemcpyAsync(bins, c->d_bins, sizeof(bins), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if(bins[6]) return XB200_ERR_INVALID_ARGUMENT;
    if(bins[4]) return XB200_ERR_UNSUPPORTED;
    // largest blocks first: they are the long poles of the launch sequence
    if((r = fork_streams(c))) return r;
    if((r = launch_me<6>(c, d_items, order + 3 * n, bins[3], d_side, bins[8 + 3]))) return r;
    if((r = launch_me<5>(c, d_items, order + 2 * n, bins[2], d_side, bins[8 + 2]))) return r;
    if((r = launch_me<4>(c, d_items, order + 1 * n, bins[1], d_side, bins[8 + 1]))) return r;
    if((r = launch_me<3>(c, d_items, order + 0 * n, bins[0], d_side, bins[8 + 0]))) return r;
    if((r = join_streams(c))) return r;
    CK(cudaEventRecord(c->ev1, c->stream));
    if((r = to_host(c, items, d_items, (size_t)n, mem))) return r;
    int err = 0;
    CK(cudaMemcpyAsync(&err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_ms = ms;
    CK(cudaGetLastError());
    if(err) { cudaMemset(c->d_err, 0, sizeof(int)); return XB200_ERR_UNEXPECTED; }
    return XB200_OK;
}

// ---- motion compensation ------------------------------------------------------------------------------------
static int check_mc(const xb200_ctx *c, const xb200_mc_item &it)
{
    if(it.w < 4 || it.h < 4 || it.w > 64 || it.h > 64 || (it.w & 3) || (it.h & 3)) return XB200_ERR_UNSUPPORTED;
    if(it.refi[0] < 0 && it.refi[1] < 0) return XB200_ERR_INVALID_ARGUMENT;
    for(int l = 0; l < 2; l++)
        if(it.refi[l] >= 0 && (!pic_ok(c, it.ref_pic[l]) || !c->pics[it.ref_pic[l]].padded)) return XB200_ERR_INVALID_ARGUMENT;
    return XB200_OK;
}

int xb200_mc(xb200_ctx *c, const xb200_mc_item *items, int64_t n, const int64_t *pred_off, int16_t *pred, int64_t pred_elems,
             int mem)
{
    if(!c || n < 0 || (n && (!items || !pred_off || !pred)) || n > (1 << 28)) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    if(mem == XB200_MEM_HOST)
        for(int64_t i = 0; i < n; i++) {
            int r = check_mc(c, items[i]);
            if(r) return r;
            if(pred_off[i] < 0 || pred_off[i] + items[i].w * items[i].h * 3 / 2 > pred_elems) return XB200_ERR_INVALID_ARGUMENT;
        }
    if(n == 0) return XB200_OK;
    int r = sync_pics(c);
    if(r) return r;
    xb200_mc_item *d_items;
    int64_t       *d_off;
    int16_t       *d_pred = pred;
    if((r = to_dev(c, c->b_items, items, (size_t)n, mem, &d_items))) return r;
    if((r = to_dev(c, c->b_aux0, pred_off, (size_t)n, mem, &d_off))) return r;
    if(mem == XB200_MEM_HOST) {
        if((r = ensure(c->b_aux1, (size_t)pred_elems * 2 + 64))) return r;
        d_pred = static_cast<int16_t *>(c->b_aux1.p);
    }
    CK(cudaEventRecord(c->ev0, c->stream));
    k_mc<<<(unsigned)n, MC_THREADS, MC_SMEM_BYTES, c->stream>>>(c->d_pics, d_items, (int)n, d_off, d_pred, c->sq);
    c->launches++;
    CK(cudaEventRecord(c->ev1, c->stream));
    if((r = to_host(c, pred, d_pred, (size_t)pred_elems, mem))) return r;
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_ms = ms;
    CK(cudaGetLastError());
    return XB200_OK;
}

int xb200_bi_org(xb200_ctx *c, const xb200_mc_item *items, int64_t n, const int32_t *cur_pic, const int64_t *off, int16_t *side,
                 int64_t side_elems, int mem)
{
    if(!c || n < 0 || (n && (!items || !cur_pic || !off || !side)) || n > (1 << 28)) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    if(mem == XB200_MEM_HOST)
        for(int64_t i = 0; i < n; i++) {
            int r = check_mc(c, items[i]);
            if(r) return r;
            if(!pic_ok(c, cur_pic[i]) || off[i] < 0 || (off[i] & 3) || off[i] + items[i].w * items[i].h > side_elems)
                return XB200_ERR_INVALID_ARGUMENT;
        }
    if(n == 0) return XB200_OK;
    int r = sync_pics(c);
    if(r) return r;
    xb200_mc_item *d_items;
    int32_t       *d_cur;
    int64_t       *d_off;
    int16_t       *d_side = side;
    if((r = to_dev(c, c->b_items, items, (size_t)n, mem, &d_items))) return r;
    if((r = to_dev(c, c->b_aux0, off, (size_t)n, mem, &d_off))) return r;
    if((r = to_dev(c, c->b_aux2, cur_pic, (size_t)n, mem, &d_cur))) return r;
    if(mem == XB200_MEM_HOST) {
        if((r = ensure(c->b_aux1, (size_t)side_elems * 2 + 64))) return r;
        d_side = static_cast<int16_t *>(c->b_aux1.p);
        CK(cudaMemsetAsync(d_side, 0, (size_t)side_elems * 2, c->stream));
    }
    CK(cudaEventRecord(c->ev0, c->stream));
    k_bi_org<<<(unsigned)n, MC_THREADS, MC_SMEM_BYTES, c->stream>>>(c->d_pics, d_items, (int)n, d_cur, d_off, d_side, c->sq);
    c->launches++;
    CK(cudaEventRecord(c->ev1, c->stream));
    if((r = to_host(c, side, d_side, (size_t)side_elems, mem))) return r;
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_ms = ms;
    CK(cudaGetLastError());
    return XB200_OK;
}

// ---- transform / quantisation ---------------------------------------------------------------------------------
static int check_tq(const xb200_tq_item &it, int64_t n_rates, int64_t elems)
{
    if(it.log2_cuw != it.log2_cuh || it.log2_cuw < 3 || it.log2_cuw > 6) return XB200_ERR_UNSUPPORTED;
    if(it.rate_idx < 0 || it.rate_idx >= n_rates) return XB200_ERR_INVALID_ARGUMENT;
    if(it.in_off < 0 || it.in_off + ((int64_t)3 << (2 * it.log2_cuw)) / 2 > elems) return XB200_ERR_INVALID_ARGUMENT;
    return XB200_OK;
}

int xb200_tq(xb200_ctx *c, xb200_tq_item *items, int64_t n, const xb200_rates *rates, int64_t n_rates, int16_t *coef,
             int64_t coef_elems, int mem)
{
    if(!c || n < 0 || (n && (!items || !coef || !rates)) || n > (1 << 28)) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    if(mem == XB200_MEM_HOST)
        for(int64_t i = 0; i < n; i++) {
            int r = check_tq(items[i], n_rates, coef_elems);
            if(r) return r;
        }
    if(n == 0) return XB200_OK;
    int            r;
    xb200_tq_item *d_items;
    xb200_rates   *d_rates;
    int16_t       *d_coef;
    if((r = to_dev(c, c->b_items, items, (size_t)n, mem, &d_items))) return r;
    if((r = to_dev(c, c->b_aux0, rates, (size_t)n_rates, mem, &d_rates))) return r;
    if((r = to_dev(c, c->b_aux1, coef, (size_t)coef_elems, mem, &d_coef))) return r;
    CK(cudaEventRecord(c->ev0, c->stream));
    k_tq<<<(unsigned)n, TQ_THREADS, sizeof(TqSmem), c->stream>>>(d_items, (int)n, d_rates, d_coef, c->d_tm64, c->sq);
    c->launches++;
    CK(cudaEventRecord(c->ev1, c->stream));
    if((r = to_host(c, items, d_items, (size_t)n, mem))) return r;
    if((r = to_host(c, coef, d_coef, (size_t)coef_elems, mem))) return r;
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_ms = ms;
    CK(cudaGetLastError());
    return XB200_OK;
}

int xb200_itdq(xb200_ctx *c, const xb200_tq_item *items, int64_t n, int16_t *coef, int64_t coef_elems, int mem)
{
    if(!c || n < 0 || (n && (!items || !coef)) || n > (1 << 28)) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    if(mem == XB200_MEM_HOST)
        for(int64_t i = 0; i < n; i++) {
            int r = check_tq(items[i], INT64_MAX, coef_elems);
            if(r) return r;
        }
    if(n == 0) return XB200_OK;
    int            r;
    xb200_tq_item *d_items;
    int16_t       *d_coef;
    if((r = to_dev(c, c->b_items, items, (size_t)n, mem, &d_items))) return r;
    if((r = to_dev(c, c->b_aux1, coef, (size_t)coef_elems, mem, &d_coef))) return r;
    CK(cudaEventRecord(c->ev0, c->stream));
    k_itdq<<<(unsigned)n, TQ_THREADS, sizeof(TqSmem), c->stream>>>(d_items, (int)n, d_coef, c->d_tm64, c->sq);
    c->launches++;
    CK(cudaEventRecord(c->ev1, c->stream));
    if((r = to_host(c, coef, d_coef, (size_t)coef_elems, mem))) return r;
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_ms = ms;
    CK(cudaGetLastError());
    return XB200_OK;
}

int xb200_recon(xb200_ctx *c, const xb200_tq_item *items, int64_t n, const int16_t *resi, const int16_t *pred, int16_t *rec,
                int64_t elems, int mem)
{
    if(!c || n < 0 || (n && (!items || !resi || !pred || !rec)) || n > (1 << 28)) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    if(mem == XB200_MEM_HOST)
        for(int64_t i = 0; i < n; i++) {
            int r = check_tq(items[i], INT64_MAX, elems);
            if(r) return r;
        }
    if(n == 0) return XB200_OK;
    int            r;
    xb200_tq_item *d_items;
    int16_t       *d_resi, *d_pred, *d_rec = rec;
    if((r = to_dev(c, c->b_items, items, (size_t)n, mem, &d_items))) return r;
    if((r = to_dev(c, c->b_aux0, resi, (size_t)elems, mem, &d_resi))) return r;
    if((r = to_dev(c, c->b_aux1, pred, (size_t)elems, mem, &d_pred))) return r;
    if(mem == XB200_MEM_HOST) {
        if((r = ensure(c->b_aux2, (size_t)elems * 2 + 64))) return r;
        d_rec = static_cast<int16_t *>(c->b_aux2.p);
        CK(cudaMemsetAsync(d_rec, 0, (size_t)elems * 2, c->stream));
    }
    CK(cudaEventRecord(c->ev0, c->stream));
    k_recon<<<(unsigned)n, 128, 0, c->stream>>>(d_items, (int)n, d_resi, d_pred, d_rec, c->sq);
    c->launches++;
    CK(cudaEventRecord(c->ev1, c->stream));
    if((r = to_host(c, rec, d_rec, (size_t)elems, mem))) return r;
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_ms = ms;
    CK(cudaGetLastError());
    return XB200_OK;
}

static int residue_impl(xb200_ctx *c, xb200_residue_item *items, int64_t n, const xb200_rates *rates, int64_t n_rates, int16_t *coef,
                        int16_t *rec, int64_t elems, int mem, int16_t *pred_dev);
int xb200_residue(xb200_ctx *c, xb200_residue_item *items, int64_t n, const xb200_rates *rates, int64_t n_rates, int16_t *coef,
                  int16_t *rec, int64_t elems, int mem)
{
    return residue_impl(c, items, n, rates, n_rates, coef, rec, elems, mem, nullptr);
}
// pred_dev (device pointer or null): the prediction of every item is stored next to its coefficients (CU pipeline)
static int residue_impl(xb200_ctx *c, xb200_residue_item *items, int64_t n, const xb200_rates *rates, int64_t n_rates, int16_t *coef,
                        int16_t *rec, int64_t elems, int mem, int16_t *pred_dev)
{
    if(!c || n < 0 || (n && (!items || !coef || !rates)) || n > (1 << 28)) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    static const bool timing = getenv("XB200_TIMING") != nullptr;   // per-phase wall times of the host-buffer path on stderr
    auto              now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double            t_ph[8] = {now(), 0, 0, 0, 0, 0, 0, 0};
    if(n == 0) return XB200_OK;
    t_ph[1] = now();
    int r = sync_pics(c);
    if(r) return r;
    xb200_residue_item *d_items;
    xb200_rates        *d_rates;
    int16_t            *d_coef = coef, *d_rec = rec;
    if((r = to_dev(c, c->b_items, items, (size_t)n, mem, &d_items))) return r;
    if((r = to_dev(c, c->b_aux0, rates, (size_t)n_rates, mem, &d_rates))) return r;
    if(mem == XB200_MEM_HOST) {
        if((r = ensure(c->b_aux1, (size_t)elems * 2 + 64))) return r;
        d_coef = static_cast<int16_t *>(c->b_aux1.p);
        if(rec) {
            if((r = ensure(c->b_aux2, (size_t)elems * 2 + 64))) return r;
            d_rec = static_cast<int16_t *>(c->b_aux2.p);
        }
    }
    if((r = ensure(c->b_order, sizeof(int32_t) * 4 * (size_t)n))) return r;
    int32_t *order = static_cast<int32_t *>(c->b_order.p);
    CK(cudaEventRecord(c->ev0, c->stream));
    CK(cudaMemsetAsync(c->d_bins, 0, sizeof(int) * 16, c->stream));
    // host-buffer calls: the records are validated by the binning kernel (bins[6]: invalid argument, bins[7] / bins[4]: unsupported)
    const BinCheck ck = {c->d_pics, (int)c->pics.size(), mem == XB200_MEM_HOST, 0, (int)n_rates, 0, (long long)elems};
    k_res_bin<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(d_items, (int)n, order, c->d_bins, ck);
    c->launches++;
    int bins[16];
    CK(cudaMemcpyAsync(bins, c->d_bins, sizeof(bins), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if(bins[6]) return XB200_ERR_INVALID_ARGUMENT;
    if(bins[4] || bins[7]) return XB200_ERR_UNSUPPORTED;
    if((r = fork_streams(c))) return r;
    if((r = launch_residue2<6>(c, d_items, order + 3 * n, bins[3], d_rates, d_coef, d_rec, pred_dev))) return r;
    if((r = launch_residue2<5>(c, d_items, order + 2 * n, bins[2], d_rates, d_coef, d_rec, pred_dev))) return r;
    if((r = launch_residue2<4>(c, d_items, order + 1 * n, bins[1], d_rates, d_coef, d_rec, pred_dev))) return r;
    if((r = launch_residue2<3>(c, d_items, order + 0 * n, bins[0], d_rates, d_coef, d_rec, pred_dev))) return r;
    if((r = join_streams(c))) return r;
    CK(cudaEventRecord(c->ev1, c->stream));
    if((r = to_host(c, items, d_items, (size_t)n, mem))) return r;
    if(rec && (r = to_host(c, rec, d_rec, (size_t)elems, mem))) return r;
    if(mem == XB200_MEM_HOST) {
        // coefficient read-back, compacted: a plane without a non-zero level (94 % of the bytes at the bench's QP) is not copied and
        // the caller's buffer is left untouched there -- the reference never reads coefficients of a plane whose nnz is 0.  The
        // device packs the non-zero planes and lists them; the host only walks that list (no pass over the records).
        if((r = ensure(c->b_compact, (size_t)elems * 2 + 64))) return r;
        if((r = ensure(c->b_coff, (size_t)3 * n * sizeof(PlaneRef) + 64))) return r;
        int16_t  *d_compact = static_cast<int16_t *>(c->b_compact.p);
        PlaneRef *d_list = static_cast<PlaneRef *>(c->b_coff.p);
        CK(cudaMemsetAsync(c->d_bins + 12, 0, sizeof(int) * 4, c->stream));
        k_compact_planes<<<(unsigned)((n + 7) / 8), 256, 0, c->stream>>>(d_items, (int)n, d_coef, d_compact, d_list,
                                                                        reinterpret_cast<unsigned long long *>(c->d_bins + 12), c->d_bins + 14);
        c->launches++;
        struct { unsigned long long total; int cnt, pad; } hdr;
        CK(cudaMemcpyAsync(&hdr, c->d_bins + 12, sizeof(hdr), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        t_ph[2] = now();
        if(hdr.cnt) {
            const size_t need = (size_t)hdr.total * 2 + (size_t)hdr.cnt * sizeof(PlaneRef) + 64;
            if(c->h_pin_cap < need) {
                if(c->h_pin) cudaFreeHost(c->h_pin);
                c->h_pin = nullptr; c->h_pin_cap = 0;
                const size_t want = need + need / 2 + 4096;
                CK(cudaHostAlloc(&c->h_pin, want, cudaHostAllocDefault));
                c->h_pin_cap = want;
            }
            PlaneRef *h_list = static_cast<PlaneRef *>(c->h_pin);
            int16_t  *h_data = reinterpret_cast<int16_t *>(h_list + hdr.cnt);
            CK(cudaMemcpyAsync(h_list, d_list, (size_t)hdr.cnt * sizeof(PlaneRef), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaMemcpyAsync(h_data, d_compact, (size_t)hdr.total * 2, cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            t_ph[3] = now();
            for(int k = 0; k < hdr.cnt; k++) memcpy(coef + h_list[k].dst, h_data + h_list[k].src, (size_t)h_list[k].elems * 2);
        }
        t_ph[4] = now();
        if(timing)
            fprintf(stderr, "xb200_residue host path: upload + kernels + records back %.3f | list + compact copy %.3f | scatter of %d planes %.3f ms\n",
                    t_ph[2] - t_ph[1], t_ph[3] > 0 ? t_ph[3] - t_ph[2] : 0.0, hdr.cnt, t_ph[3] > 0 ? t_ph[4] - t_ph[3] : 0.0);
    }
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->last_ms = ms;
    CK(cudaGetLastError());
    return XB200_OK;
}

} // extern "C"

namespace {
template <int L2> int launch_skip(xb200_ctx *c, const xb200_cu_item *d_items, const int32_t *order, int cnt, const xb200_sbac *d_in,
                                         CuState *d_state, int16_t *d_scr, int64_t elems)
{
    if(cnt == 0) return XB200_OK;
    using Cf = SkipCfg<L2>;
    CK(cudaFuncSetAttribute(k_cu_skip<L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cf::SMEM));
    const int want = (cnt + Cf::TEAMS - 1) / Cf::TEAMS, grid = want < c->sms * 8 ? want : c->sms * 8;
    k_cu_skip<L2><<<grid, Cf::CTA, Cf::SMEM, c->side[L2 - 3]>>>(c->d_pics, d_items, order, cnt, d_in, d_state, d_scr, elems, c->sq);
    c->launches++;
    CK(cudaGetLastError());
    return XB200_OK;
}

} // namespace

extern "C" {

// cbf decisions of all residue slots.  Small CUs (8x8: short bin streams, three quarters of the slots): one coder per LANE; large CUs (long
// streams: the serial critical path matters, not throughput): one coder per warp with ballot-parallel coefficient fetch and
// register-resident unary runs.  The two grids run concurrently.  XB200_DECIDE=warp / lanes forces one kernel for all sizes.
static int run_decide(xb200_ctx *c, const xb200_cu_item *d_items, int n_slots, int per_cu, const xb200_sbac *d_in, CuState *d_state,
                      const xb200_residue_item *d_res, const int16_t *d_scr, int64_t elems)
{
    const char *e = getenv("XB200_DECIDE");
    int         split = (e && e[0] == 'w') ? 0 : ((e && e[0] == 'l') ? 64 : 8);    // widths <= split go to the lane coder (8: measured best)
    if(e && e[0] >= '0' && e[0] <= '9') split = atoi(e);
    int r;
    if((r = fork_streams(c))) return r;
    if(split < 64) {
        k_cu_decide<<<(n_slots + PIPE_WARPS - 1) / PIPE_WARPS, PIPE_WARPS * 32, 0, c->side[0]>>>(d_items, n_slots, per_cu, d_in, d_state, d_res, d_scr,
                                                                                         split + 1, 64);
        c->launches++;
    }
    if(split > 0) {
        if((r = ensure(c->b_cu_nzr, (size_t)5 * elems * 2 + 64))) return r;
        if((r = ensure(c->b_cu_nzl, (size_t)5 * elems * 2 + 64))) return r;
        if((r = ensure(c->b_cu_meta, (size_t)n_slots + 64))) return r;
        uint16_t *nzr = static_cast<uint16_t *>(c->b_cu_nzr.p);
        int16_t  *nzl = static_cast<int16_t *>(c->b_cu_nzl.p);
        uint8_t  *meta = static_cast<uint8_t *>(c->b_cu_meta.p);
        k_cu_nzlist<<<(n_slots + 3) / 4, 128, 0, c->side[1]>>>(d_res, n_slots, d_scr, nzr, nzl, meta, elems, 1, split);
        const int smem = 4 * (int)sizeof(LcShared);
        CK(cudaFuncSetAttribute(k_cu_decide_lanes, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        k_cu_decide_lanes<<<(n_slots + 127) / 128, 128, smem, c->side[1]>>>(d_items, n_slots, per_cu, d_in, d_state, d_res, nzr, nzl, meta, elems, 1,
                                                                           split);
        c->launches += 2;
    }
    CK(cudaGetLastError());
    return join_streams(c);
}

static int analyze_pipeline(xb200_ctx *c, xb200_cu_item *items, int64_t n, const xb200_rates *rates, int64_t n_rates, xb200_sbac *states,
                            int64_t n_states, int16_t *coef, int16_t *rec, int64_t elems, const int32_t *order, const int *cnt)
{
    int r;
    xb200_cu_item *d_items;
    xb200_rates   *d_rates;
    xb200_sbac    *d_in, *d_out;
    int32_t       *d_order;
    const int64_t  side_elems = elems / 3 * 2 + 64;
    if((r = to_dev(c, c->b_cu_items, items, (size_t)n, XB200_MEM_HOST, &d_items))) return r;
    if((r = to_dev(c, c->b_cu_rates, rates, (size_t)n_rates, XB200_MEM_HOST, &d_rates))) return r;
    if((r = to_dev(c, c->b_st0, states, (size_t)n_states, XB200_MEM_HOST, &d_in))) return r;
    if((r = to_dev(c, c->b_st1, states, (size_t)n_states, XB200_MEM_HOST, &d_out))) return r;
    if((r = to_dev(c, c->b_cu_order, order, (size_t)4 * n, XB200_MEM_HOST, &d_order))) return r;
    if((r = ensure(c->b_cu_state, (size_t)n * sizeof(CuState)))) return r;
    if((r = ensure(c->b_cu_me, (size_t)n * 8 * sizeof(xb200_me_item)))) return r;
    if((r = ensure(c->b_cu_res, (size_t)n * 3 * sizeof(xb200_residue_item)))) return r;
    if((r = ensure(c->b_cu_mc, (size_t)n * sizeof(xb200_mc_item)))) return r;
    if((r = ensure(c->b_cu_cur, (size_t)n * sizeof(int32_t)))) return r;
    if((r = ensure(c->b_cu_off, (size_t)n * sizeof(int64_t)))) return r;
    if((r = ensure(c->b_cu_side, (size_t)side_elems * 2))) return r;
    if((r = ensure(c->b_scr[0], (size_t)15 * elems * 2 + 64))) return r;
    if((r = ensure(c->b_cu_coef, (size_t)elems * 2 + 64))) return r;
    if(rec && (r = ensure(c->b_cu_rec, (size_t)elems * 2 + 64))) return r;
    CuState            *d_state = static_cast<CuState *>(c->b_cu_state.p);
    xb200_me_item      *d_me = static_cast<xb200_me_item *>(c->b_cu_me.p);
    xb200_residue_item *d_res = static_cast<xb200_residue_item *>(c->b_cu_res.p);
    xb200_mc_item      *d_mc = static_cast<xb200_mc_item *>(c->b_cu_mc.p);
    int32_t            *d_cur = static_cast<int32_t *>(c->b_cu_cur.p);
    int64_t            *d_off = static_cast<int64_t *>(c->b_cu_off.p);
    int16_t            *d_side = static_cast<int16_t *>(c->b_cu_side.p), *d_scr = static_cast<int16_t *>(c->b_scr[0].p);
    int16_t            *d_coef = static_cast<int16_t *>(c->b_cu_coef.p), *d_rec = rec ? static_cast<int16_t *>(c->b_cu_rec.p) : nullptr;
    const int           ni = (int)n;
    bool                any_b = false;
    for(int64_t i = 0; i < n && !any_b; i++) any_b = items[i].slice_type == 0;
    CK(cudaStreamSynchronize(c->stream)); // pageable `order`
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, c->stream));
    // 1. skip candidates
    if((r = fork_streams(c))) return r;
    if((r = launch_skip<6>(c, d_items, d_order + 3 * n, cnt[3], d_in, d_state, d_scr, elems))) return r;
    if((r = launch_skip<5>(c, d_items, d_order + 2 * n, cnt[2], d_in, d_state, d_scr, elems))) return r;
    if((r = launch_skip<4>(c, d_items, d_order + 1 * n, cnt[1], d_in, d_state, d_scr, elems))) return r;
    if((r = launch_skip<3>(c, d_items, d_order + 0 * n, cnt[0], d_in, d_state, d_scr, elems))) return r;
    if((r = join_streams(c))) return r;
    // 2. uni-directional searches
    k_cu_make_me_uni<<<(ni * 8 + 255) / 256, 256, 0, c->stream>>>(d_items, ni, d_state, d_me, c->sq);
    c->launches++;
    if((r = xb200_me(c, d_me, (int64_t)ni * 8, nullptr, 0, XB200_MEM_DEVICE))) return r;
    // 3. best reference / MVP index, DIR | L0 | L1 candidates through the residue operator, cbf decisions
    k_cu_after_uni<<<(ni + PIPE_WARPS - 1) / PIPE_WARPS, PIPE_WARPS * 32, 0, c->stream>>>(d_items, ni, d_in, d_state, d_me, d_res, elems);
    c->launches++;
    if((r = residue_impl(c, d_res, (int64_t)ni * 3, d_rates, n_rates, d_scr, d_scr + elems, 15 * elems, XB200_MEM_DEVICE, d_scr + 2 * elems))) return r;
    if((r = run_decide(c, d_items, ni * 3, 3, d_in, d_state, d_res, d_scr, elems))) return r;
    // 4. analyze_bi
    if(any_b) {
        for(int iter = 0; iter < 4; iter++) {
            k_cu_bi_prep<<<(ni + 127) / 128, 128, 0, c->stream>>>(d_items, ni, d_state, iter, d_mc, d_cur, d_off, d_me, c->sq);
            c->launches++;
            if((r = xb200_bi_org(c, d_mc, n, d_cur, d_off, d_side, side_elems, XB200_MEM_DEVICE))) return r;
            if((r = xb200_me(c, d_me, (int64_t)ni * XB200_MAX_REFP, d_side, side_elems, XB200_MEM_DEVICE))) return r;
            CK(cudaMemsetAsync(c->d_bins, 0, sizeof(int), c->stream));
            k_cu_bi_update<<<(ni + 127) / 128, 128, 0, c->stream>>>(ni, d_state, d_me, c->d_bins);
            c->launches++;
            int active = 0;
            CK(cudaMemcpyAsync(&active, c->d_bins, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            if(!active) break;
        }
        k_cu_bi_emit<<<(ni + 127) / 128, 128, 0, c->stream>>>(d_items, ni, d_state, d_res, elems);
        c->launches++;
        if((r = residue_impl(c, d_res, n, d_rates, n_rates, d_scr, d_scr + elems, 15 * elems, XB200_MEM_DEVICE, d_scr + 2 * elems))) return r;
        if((r = run_decide(c, d_items, ni, 1, d_in, d_state, d_res, d_scr, elems))) return r;
    }
    // 5. winners
    k_cu_final<<<ni, 128, 0, c->stream>>>(d_items, ni, d_state, d_scr, elems, d_out, d_coef, d_rec);
    c->launches++;
    CK(cudaEventRecord(e1, c->stream));
    if((r = to_host(c, items, d_items, (size_t)n, XB200_MEM_HOST))) return r;
    if((r = to_host(c, states, d_out, (size_t)n_states, XB200_MEM_HOST))) return r;
    if((r = to_host(c, coef, d_coef, (size_t)elems, XB200_MEM_HOST))) return r;
    if(rec && (r = to_host(c, rec, d_rec, (size_t)elems, XB200_MEM_HOST))) return r;
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    c->last_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    CK(cudaGetLastError());
    return XB200_OK;
}

} // extern "C"
