// Rate estimation for the inter RDO: the CABAC engine in bit-counting mode and the RDOQ rate tables.
//
//   reference: src_base/xeve_eco.c:455-620 (engine), :674-1260 (syntax elements), src_base/xeve_mode.c:39-373 (counters,
//   xeve_rdoq_bit_est).  In bit-counting mode the encoder's low register never influences the count: after
//   xeve_sbac_bit_reset, xeve_get_bit_number equals the number of renormalisation shifts, so the device engine carries
//   only {range, bits, models}.  Bins of one item are inherently serial (every bin updates a model and the range), so a
//   warp owns an item: lane 0 runs the engine on models held in shared memory, all 32 lanes fetch coefficients in
//   zig-zag order and hand lane 0 only the non-zero ones (ballot + shuffle), which removes the 4096-step serial scan.
#pragma once
#include "xb200_common.cuh"
#include "xb200_cabac.cuh"

XB200_CONST_LINKAGE __device__ uint16_t g_scan[16 + 64 + 256 + 1024 + 4096];   // zig-zag: scan position -> raster, log2 size 2..6
XB200_CONST_LINKAGE __device__ int32_t  g_entropy_bits[1024];                  // xeve_init_bits_est, computed on the host with libm
__device__ __forceinline__ int scan_base(int l2) { return l2 == 2 ? 0 : l2 == 3 ? 16 : l2 == 4 ? 80 : l2 == 5 ? 336 : 1360; }

__device__ __forceinline__ void cb_mvp_idx(Cabac &c, int v)
{
    for(int i = 0; i < 3; i++) {
        cb_bin(c, XB200_CM_MVP_IDX + i, i != v);
        if(i == v) break;
    }
}
__device__ __forceinline__ void cb_mvd(Cabac &c, const int16_t *mvd)
{
    for(int k = 0; k < 2; k++) {
        const uint32_t a = (uint32_t)abs((int)mvd[k]);
        // exp-golomb prefix/suffix of xeve_eco_abs_mvd: len_i = bit length of (a+1)>>1, the first two bins are context coded
        const int      len_i = 32 - __clz((a + 1) >> 1);
        const int      len_c = 2 * len_i + 1;
        const uint32_t code = (1u << len_i) | ((a + 1 - (1u << len_i)) & ((1u << len_i) - 1));
        cb_bin(c, XB200_CM_MVD, (code >> (len_c - 1)) & 1);
        if(len_c > 1) cb_bin(c, XB200_CM_MVD, (code >> (len_c - 2)) & 1);
        cb_ep(c, max(0, len_c - 2) + (a != 0));
    }
}
__device__ __forceinline__ void cb_refi(Cabac &c, int num_refp, int refi)
{
    if(num_refp <= 1) return;
    cb_bin(c, XB200_CM_REFI, refi != 0);
    if(refi == 0) return;
    for(int i = 2; i < num_refp; i++) {
        const int bin = i != refi + 1;
        if(i == 2) cb_bin(c, XB200_CM_REFI + 1, bin); else cb_ep(c);
        if(!bin) break;
    }
}
// xeve_eco_run_length_cc of one transform block; warp-cooperative, engine state valid on lane 0 only
__device__ __forceinline__ void cb_run_length(Cabac &c, const int16_t *coef, int l2, int num_sig, int ch, int lane)
{
    const uint16_t *scan = g_scan + scan_base(l2);
    const int       n = 1 << (2 * l2), t0 = ch == 0 ? 0 : 2;
    uint32_t        run = 0;
    bool            done = false;
    for(int base = 0; base < n && !done; base += 32) {
        const int sp = base + lane;
        const int v = sp < n ? __ldcg(coef + scan[sp]) : 0;   // L2 read: the planes may have been written by this CTA
        uint32_t  nzm = __ballot_sync(0xffffffffu, v != 0);
        int       prev = -1;
        while(nzm && !done) {
            const int b = __ffs(nzm) - 1;
            nzm &= nzm - 1;
            const int cv = __shfl_sync(0xffffffffu, v, b);
            if(lane == 0) {
                run += b - prev - 1;
                cb_unary(c, run, XB200_CM_RUN + t0);
                cb_unary(c, (uint32_t)abs(cv) - 1, XB200_CM_LEVEL + t0);
                cb_ep(c);   // sign
            }
            prev = b;
            run = 0;
            if(base + b == n - 1) { done = true; break; }
            num_sig--;
            if(lane == 0) cb_bin(c, XB200_CM_LAST + (ch != 0), num_sig == 0);
            if(num_sig == 0) done = true;
        }
        run += 31 - prev;
    }
}
// xeve_eco_coef for an inter CU of at most 64x64 (one transform block per plane)
__device__ __forceinline__ void cb_coef(Cabac &c, const int *nnz, int l2w, int l2h, const int16_t *coef, int run_stats, int lane)
{
    const int r0 = run_stats & 1, r1 = (run_stats >> 1) & 1, r2 = (run_stats >> 2) & 1;
    const int f0 = nnz[0] != 0, f1 = nnz[1] != 0, f2 = nnz[2] != 0;
    if(r0 + r1 + r2 == 3) {
        const int all = f0 + f1 + f2;
        if(lane == 0) cb_bin(c, XB200_CM_CBF_ALL, all != 0);
        if(!all) return;
    }
    if(lane == 0) {
        if(r1) cb_bin(c, XB200_CM_CBF_CB, f1);
        if(r2) cb_bin(c, XB200_CM_CBF_CR, f2);
        if(r0 && (f1 + f2 != 0)) cb_bin(c, XB200_CM_CBF_LUMA, f0);
    }
    const int ny = 1 << (l2w + l2h);
    if(f0 && r0) cb_run_length(c, coef, l2w, nnz[0], 0, lane);
    if(f1 && r1) cb_run_length(c, coef + ny, l2w - 1, nnz[1], 1, lane);
    if(f2 && r2) cb_run_length(c, coef + ny + (ny >> 2), l2w - 1, nnz[2], 2, lane);
}
// the four counters of src_base/xeve_mode.c:57-302 on one item (see xb200_bits_item in include/xeve_b200.h)
__device__ __forceinline__ void cb_count_item(Cabac &c, const xb200_bits_item &it, const int16_t *coef, int lane)
{
    const bool B = it.slice_type == 0, inter = it.slice_type != 2;
    if(it.kind == 0) {
        if(lane == 0 && inter) {
            cb_bin(c, XB200_CM_SKIP_FLAG + it.ctx_skip, 1);
            cb_mvp_idx(c, it.mvp_idx[0]);
            if(B) cb_mvp_idx(c, it.mvp_idx[1]);
        }
    }
    else if(it.kind == 1) {
        if(lane == 0 && inter) {
            cb_bin(c, XB200_CM_SKIP_FLAG + it.ctx_skip, 0);
            if(it.all_preds) cb_bin(c, XB200_CM_PRED_MODE + it.ctx_pred_mode, 0);
            cb_bin(c, XB200_CM_DIRECT, it.pidx == 4);
            if(it.pidx != 4) {
                if(it.refi[0] >= 0 && it.refi[1] >= 0) cb_bin(c, XB200_CM_INTER_DIR, 0);
                else {
                    if(B) cb_bin(c, XB200_CM_INTER_DIR, 1);
                    cb_bin(c, XB200_CM_INTER_DIR + 1, it.refi[0] < 0);
                }
                if(it.refi[0] >= 0) { cb_refi(c, it.num_refp[0], it.refi[0]); cb_mvp_idx(c, it.mvp_idx[0]); cb_mvd(c, it.mvd[0]); }
                if(B && it.refi[1] >= 0) { cb_refi(c, it.num_refp[1], it.refi[1]); cb_mvp_idx(c, it.mvp_idx[1]); cb_mvd(c, it.mvd[1]); }
            }
        }
        cb_coef(c, it.nnz, it.log2_cuw, it.log2_cuh, coef + it.coef_off, 7, lane);
    }
    else if(it.kind == 2) {
        if(lane == 0 && it.pidx != 4) {
            if(inter && it.refi[0] >= 0) { cb_mvp_idx(c, it.mvp_idx[0]); cb_mvd(c, it.mvd[0]); }
            if(B && it.refi[1] >= 0) { cb_mvp_idx(c, it.mvp_idx[0]); cb_mvd(c, it.mvd[1]); }
        }
    }
    else cb_coef(c, it.nnz, it.log2_cuw, it.log2_cuh, coef + it.coef_off, 1 << it.ch, lane);
}

#ifndef XB200_CHAIN_TU   // non-template kernels: defined once, in the translation unit of the operators
constexpr int RATE_WARPS = 4;
__global__ void __launch_bounds__(RATE_WARPS * 32) k_rdo_bits(xb200_bits_item *__restrict__ items, long long n,
                                                               const xb200_sbac *__restrict__ st_in, xb200_sbac *__restrict__ st_out,
                                                               const int16_t *__restrict__ coef)
{
    __shared__ uint16_t sm[RATE_WARPS][XB200_CM_COUNT + 4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for(long long i = (long long)blockIdx.x * RATE_WARPS + w; i < n; i += (long long)gridDim.x * RATE_WARPS) {
        const xb200_bits_item it = items[i];
        const xb200_sbac     &s = st_in[it.state_in];
        for(int k = lane; k < XB200_CM_COUNT; k += 32) sm[w][k] = s.m[k];
        __syncwarp();
        Cabac c;
        c.range = s.range; c.bits = 0; c.m = sm[w];
        cb_count_item(c, it, coef, lane);
        __syncwarp();
        if(lane == 0) items[i].bits = c.bits;
        if(it.state_out >= 0) {
            const uint32_t rg = __shfl_sync(0xffffffffu, c.range, 0);
            xb200_sbac    &o = st_out[it.state_out];
            if(lane == 0) o.range = rg;
            for(int k = lane; k < XB200_CM_COUNT; k += 32) o.m[k] = sm[w][k];
        }
        __syncwarp();
    }
}

#endif
// xeve_rdoq_bit_est: one thread per (state, model, bin)
__device__ __forceinline__ int32_t rate_of(uint16_t model, int bin)
{
    const uint32_t mps = model & 1, state = model >> 1;
    return g_entropy_bits[(((uint32_t)bin != mps) ? state : 512 - state) << 1];
}
#ifndef XB200_CHAIN_TU
__global__ void k_rdoq_rates(const xb200_sbac *__restrict__ st, long long n, xb200_rates *__restrict__ out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = t >> 6;
    const int       k = (int)(t & 63);
    if(i >= n) return;
    const uint16_t *m = st[i].m;
    xb200_rates    &o = out[i];
    if(k < 24) {
        for(int b = 0; b < 2; b++) { o.run[k][b] = rate_of(m[XB200_CM_RUN + k], b); o.level[k][b] = rate_of(m[XB200_CM_LEVEL + k], b); }
    }
    else if(k < 26) {
        for(int b = 0; b < 2; b++) o.last[k - 24][b] = rate_of(m[XB200_CM_LAST + k - 24], b);
    }
    else if(k == 26) {
        for(int b = 0; b < 2; b++) {
            o.cbf_all[b] = rate_of(m[XB200_CM_CBF_ALL], b); o.cbf_luma[b] = rate_of(m[XB200_CM_CBF_LUMA], b);
            o.cbf_cb[b] = rate_of(m[XB200_CM_CBF_CB], b); o.cbf_cr[b] = rate_of(m[XB200_CM_CBF_CR], b);
        }
    }
}
#endif
