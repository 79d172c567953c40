// xb200_ctx.h -- host-side context of the library, shared by its translation units (xb200_api.cu: the per-CU operators,
// xb200_frame.cu: the whole-picture operators, xb200_intra.cu: intra analysis, xb200_main.cu: Main-profile operators).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "xb200_common.cuh"

struct DevBuf {
    void  *p   = nullptr;
    size_t cap = 0;
};

struct Pic {
    bool     used = false, padded = false;
    int      pad[3] = {0, 0, 0}, s[3] = {0, 0, 0}, w[3] = {0, 0, 0}, h[3] = {0, 0, 0};
    int16_t *buf[3] = {nullptr, nullptr, nullptr};
};

struct xb200_ctx {
    int              device = 0;
    cudaStream_t     stream = nullptr;
    xb200_seq        seq{};
    SeqDev           sq{};
    std::vector<Pic> pics;
    std::vector<void *> garbage;    // buffers of destroyed pictures whose slot changed kind: freed with the context
    PicDev          *d_pics = nullptr;
    int              d_pics_cap = 0;
    bool             pics_dirty = true;
    int8_t          *d_tm64 = nullptr;
    int             *d_err = nullptr;
    int             *d_bins = nullptr; // 8 counters + 8 max-range
    DevBuf           b_items, b_side, b_aux0, b_aux1, b_aux2, b_order, b_stage;
    DevBuf           b_df; // deblocking: per-SCU edge flags
    DevBuf           b_in_items, b_in_rates, b_in_st0, b_in_st1, b_in_side, b_in_coef, b_in_rec, b_in_order; // intra analysis
    bool             intra_ready = false;
    void            *h_pin = nullptr;   // library-owned pinned staging (compacted coefficient read-back)
    size_t           h_pin_cap = 0;
    DevBuf           b_compact, b_coff;
    DevBuf           b_ats;             // Main profile: the eight 8-bit ATS matrices (xb200_main.cu)
    DevBuf           b_scr[4], b_st0, b_st1; // analyze_cu: mode scratch per size class, coder states in / out
    DevBuf           b_cu_items, b_cu_rates, b_cu_state, b_cu_me, b_cu_res, b_cu_mc, b_cu_cur, b_cu_off, b_cu_side, b_cu_order,
                     b_cu_coef, b_cu_rec, b_cu_nzr, b_cu_nzl, b_cu_meta; // CU pipeline
    cudaEvent_t      ev0 = nullptr, ev1 = nullptr;
    cudaStream_t     side[4] = {nullptr, nullptr, nullptr, nullptr}; // one per CU size: the four size-binned grids overlap
    cudaEvent_t      ev_fork = nullptr, ev_join[4] = {nullptr, nullptr, nullptr, nullptr};
    int              sms = 148;          // cudaDevAttrMultiProcessorCount of `device`
    int              res2_blocks[4][2] = {{0, 0}, {0, 0}, {0, 0}, {0, 0}}; // resident CTAs per SM of k_residue2<L2, TC> on this device
    void            *chain = nullptr;    // decision-pass state (xb200_chain.cu)
    double           last_ms = 0.0;
    int64_t          launches = 0;
};

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if(e_ != cudaSuccess) {                                                                          \
            fprintf(stderr, "xeve_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return e_ == cudaErrorMemoryAllocation ? XB200_ERR_OUT_OF_MEMORY : XB200_ERR_UNEXPECTED;    \
        }                                                                                                \
    } while(0)

// defined in xb200_api.cu
int  xb200_ensure(DevBuf &b, size_t bytes);
int  xb200_finish(xb200_ctx *c);          // record ev1, synchronise, store the kernel time of [ev0, ev1]
int  xb200_pad_planes(xb200_ctx *c, Pic &p, cudaStream_t stream = nullptr); // border replication; default: the context's stream
// both loop-filter passes of picture p on `stream` from device-resident maps and edge flags (xb200_frame.cu)
int  xb200_deblock_dev(xb200_ctx *c, Pic &p, const xb200_df_pic *pp, const uint32_t *d_scu, const int8_t *d_refi, const int16_t *d_mv,
                       const uint8_t *d_flags, cudaStream_t stream);
int  xb200_preload_api_kernels();          // force-load the short kernels that run next to the chain server (lazy module loading)
int  xb200_preload_frame_kernels();
void xb200_chain_drain_end(int device);    // after xb200_chain_free and the context's cudaFree calls: the scheduler may publish again
void xb200_chain_free(xb200_ctx *c);      // releases the decision-pass state of a context (xb200_chain.cu)
int  xb200_sync_pics(xb200_ctx *c);       // refresh the device-side picture table (c->d_pics)

inline int  align_up(int v, int a) { return (v + a - 1) / a * a; }
inline bool pic_ok(const xb200_ctx *c, int h) { return h >= 0 && h < (int)c->pics.size() && c->pics[h].used; }

// bring a caller buffer to the device (or use it in place)
template <typename T> int to_dev(xb200_ctx *c, DevBuf &b, const T *src, size_t count, int mem, T **out)
{
    if(mem == XB200_MEM_DEVICE || src == nullptr) { *out = const_cast<T *>(src); return XB200_OK; }
    int r = xb200_ensure(b, count * sizeof(T) + 64);
    if(r) return r;
    if(count) CK(cudaMemcpyAsync(b.p, src, count * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    *out = static_cast<T *>(b.p);
    return XB200_OK;
}
template <typename T> int to_host(xb200_ctx *c, T *dst, const T *dev, size_t count, int mem)
{
    if(mem == XB200_MEM_DEVICE || (const void *)dst == (const void *)dev) return XB200_OK;
    if(count) CK(cudaMemcpyAsync(dst, dev, count * sizeof(T), cudaMemcpyDeviceToHost, c->stream));
    return XB200_OK;
}
