// xb200_cabac.cuh -- the CABAC engine in bit-counting mode (reference src_base/xeve_eco.c:455-620): after
// xeve_sbac_bit_reset the value of xeve_get_bit_number equals the number of renormalisation shifts, so the engine carries
// only {range, bits, context models}.  Shared by the inter counters (xb200_rate.cuh) and the intra analysis (xb200_intra.cuh).
#pragma once
#include "xb200_common.cuh"

struct Cabac {
    uint32_t  range, bits;
    uint16_t *m;   // XB200_CM_COUNT models (shared memory)
};
__device__ __forceinline__ void cb_bin(Cabac &c, int idx, int bin)
{
    const uint32_t model = c.m[idx];
    uint32_t       mps = model & 1, state = model >> 1;
    uint32_t       lps = (state * c.range) >> 9;
    lps = max(lps, 437u);
    c.range -= lps;
    if((uint32_t)(bin != 0) != mps) {
        if(c.range >= lps) c.range = lps;
        state += (512 - state + 16) >> 5;
        if(state > 256) { mps ^= 1; state = 512 - state; }
    }
    else state -= (state + 16) >> 5;
    c.m[idx] = (uint16_t)((state << 1) | mps);
    const int sh = max(0, __clz(c.range) - 18);   // shifts until range >= 8192 (bit 13 set)
    c.range <<= sh;
    c.bits += sh;
}
__device__ __forceinline__ void cb_ep(Cabac &c, int nbins = 1)   // sbac_encode_bin_ep: range >>= 1, <<= 1 drops the LSB, one shift per bin
{
    if(nbins > 0) { c.range &= ~1u; c.bits += nbins; }
}
// the engine step on a model held in registers (state / mps unpacked)
__device__ __forceinline__ void cb_step(uint32_t &range, uint32_t &bits, uint32_t &state, uint32_t &mps, uint32_t bin)
{
    uint32_t lps = (state * range) >> 9;
    lps = max(lps, 437u);
    range -= lps;
    if(bin != mps) {
        if(range >= lps) range = lps;
        state += (512 - state + 16) >> 5;
        if(state > 256) { mps ^= 1; state = 512 - state; }
    }
    else state -= (state + 16) >> 5;
    const int sh = max(0, __clz(range) - 18);
    range <<= sh;
    bits += sh;
}
// sbac_write_unary_sym with two contexts: bin 0 on model idx, the remaining `sym` bins (sym - 1 ones and a zero) on model
// idx + 1, which stays in registers for the whole run
__device__ __forceinline__ void cb_unary(Cabac &c, uint32_t sym, int idx)
{
    cb_bin(c, idx, sym != 0);
    if(sym == 0) return;
    const uint32_t model = c.m[idx + 1];
    uint32_t       mps = model & 1, state = model >> 1, range = c.range, bits = c.bits;
    for(; sym > 1; sym--) cb_step(range, bits, state, mps, 1);
    cb_step(range, bits, state, mps, 0);
    c.m[idx + 1] = (uint16_t)((state << 1) | mps);
    c.range = range; c.bits = bits;
}
