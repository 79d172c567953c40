// xb200_common.cuh -- shared device-side declarations of the B200 hot-path library.
#pragma once
// Threads of a 64x64 team: 256 in the batched operators; the decision-chain kernel (xb200_chain.cu) runs 128-thread CTAs so that more
// chains fit on an SM and defines XB200_T64 = 128 before including the device code.  Every team routine is generic in T.
#ifndef XB200_T64
#define XB200_T64 256
#endif
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/xeve_b200.h"
#include "xb200_tables.h"

// A device-resident picture: s16 planes, 4:2:0.  p[] point at the top-left sample of the active
// area; the allocation extends `pad` samples on every side for padded (reference) pictures.
// Strides are multiples of 64 samples and the active origin is 16-byte aligned, so any row
// segment that starts at a multiple of 8 samples can be fetched with one bulk async copy.
struct PicDev {
    int16_t *p[3];
    int32_t  s[3];      // stride in samples per plane
    int32_t  w, h;      // luma size
    int32_t  pad_l, pad_c;
    int32_t  valid;
};

struct SeqDev {
    int32_t w, h, bd;
    int32_t me_level, hpel_cnt, qpel_cnt, me_complexity;
    int32_t min_clip[2], max_clip[2];
    int32_t rdoq;
    int32_t merge_num, gop_size; // analyze_cu: skip candidates per list, pi->gop_size
    int32_t tc_dct; // 32/64-point forward DCT on tcgen05 (bit-identical; off by default)
};

// The tables live in the translation unit that uploads them: xb200_api.cu defines them with external linkage, xb200_intra.cu
// keeps its own (static) copies, xb200_frame.cu needs none.
#ifndef XB200_CONST_LINKAGE
#define XB200_CONST_LINKAGE
#endif
#ifndef XB200_NO_CONSTANTS
XB200_CONST_LINKAGE __constant__ int8_t  c_tm64[64 * 64];           // DCT-II matrix, N-point rows at stride 64/N
XB200_CONST_LINKAGE __constant__ int16_t c_mc_l[4][8];              // luma taps by quarter-pel phase
XB200_CONST_LINKAGE __constant__ int16_t c_mc_c[8][4];              // chroma taps by eighth-pel phase
XB200_CONST_LINKAGE __constant__ int32_t c_quant_scale[6];
XB200_CONST_LINKAGE __constant__ int32_t c_dequant_scale[6];
XB200_CONST_LINKAGE __constant__ int64_t c_err_scale[7][6][7];      // [bit depth - 8][qp % 6][log2 size] (host-computed doubles -> s64): every
                                                                    // context finds its own depth, whatever other contexts of the process use
#endif

#define XB_DEV __device__ __forceinline__

XB_DEV int clip3i(int lo, int hi, int v) { return max(lo, min(hi, v)); }
#ifndef XB200_NO_CONSTANTS
XB_DEV int tmN(int log2n, int k, int n) { return c_tm64[(k << (6 - log2n)) * 64 + n]; }
#endif

// Barrier of a TEAM of T threads (the first T threads of the CTA, or a whole warp): a named barrier, so that a team narrower than its
// CTA (the decision-chain kernel runs CUs of every size on one CTA) never waits for the idle threads.  With T == blockDim.x it is
// equivalent to __syncthreads().
template <int T> XB_DEV void team_bar()
{
    if(T == 32) __syncwarp();
    else asm volatile("bar.sync 1, %0;" ::"n"(T) : "memory");
}

// Debug builds (-DXB200_CHAIN_DEBUG): progress words in host-mapped memory, readable by the host while a kernel hangs
#ifdef XB200_CHAIN_DEBUG
XB200_CONST_LINKAGE __device__ volatile int *g_dbg;
#define CH_DBG(slot, val) do { if(g_dbg) g_dbg[(slot)] = (val); } while(0)
#else
#define CH_DBG(slot, val) do { } while(0)
#endif

XB_DEV uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier + bulk async copy (TMA engine, SASS UBLKCP) -----------------------------------
XB_DEV void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
XB_DEV void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
XB_DEV void mbar_wait(uint64_t *bar, uint32_t parity)
{
#ifdef XB200_CHAIN_DEBUG
    {   // bounded wait: a copy that never completes becomes a trap with the barrier address in the debug words
        const long long t0 = clock64();
        uint32_t done = 0;
        while(!done) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
            if(!done && clock64() - t0 > 4000000000ll) { CH_DBG(60, (int)smem_u32(bar)); CH_DBG(61, (int)parity); CH_DBG(62, (int)threadIdx.x); __threadfence_system(); __trap(); }
        }
        return;
    }
#endif
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, bytes multiple of 16, both addresses 16-byte aligned
XB_DEV void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---- packed 16-bit SAD primitives (VIMNMX.S16x2 on sm_100a) -------------------------------------
// |a - b| per 16-bit half for operands whose halves compare correctly as UNSIGNED numbers
// (non-negative samples, or signed samples biased by ^0x8000): max - min never borrows.
XB_DEV uint32_t absdiff_u16x2(uint32_t a, uint32_t b)
{
    uint32_t mx, mn;
    asm("max.u16x2 %0, %1, %2;" : "=r"(mx) : "r"(a), "r"(b));
    asm("min.u16x2 %0, %1, %2;" : "=r"(mn) : "r"(a), "r"(b));
    return mx - mn;
}
XB_DEV uint32_t sum_halves(uint32_t v) { return (v & 0xffffu) + (v >> 16); }

XB_DEV uint64_t shfl_xor_u64(uint64_t v, int m)
{
    uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m), hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
    return ((uint64_t)hi << 32) | lo;
}
XB_DEV int64_t shfl_up_s64(int64_t v, int d)
{
    uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)v, d), hi = __shfl_up_sync(0xffffffffu, (uint32_t)((uint64_t)v >> 32), d);
    return (int64_t)(((uint64_t)hi << 32) | lo);
}
XB_DEV int64_t shfl_s64(int64_t v, int src)
{
    uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src), hi = __shfl_sync(0xffffffffu, (uint32_t)((uint64_t)v >> 32), src);
    return (int64_t)(((uint64_t)hi << 32) | lo);
}
