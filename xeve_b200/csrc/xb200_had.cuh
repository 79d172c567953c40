#pragma once
#include "xb200_common.cuh"

// Hadamard SATD of one TN x TN tile held by one thread (src_base/xeve_sad.c:417-607)
template <int TN> XB_DEV int had_tile_dev(const int16_t *a, int sa, const int16_t *b, int sb)
{
    int m[TN][TN];
#pragma unroll
    for(int y = 0; y < TN; y++)
#pragma unroll
        for(int x = 0; x < TN; x++) m[y][x] = (int)a[(ptrdiff_t)y * sa + x] - (int)b[(ptrdiff_t)y * sb + x];
#pragma unroll
    for(int pass = 0; pass < 2; pass++)
#pragma unroll
        for(int r = 0; r < TN; r++)
#pragma unroll
            for(int len = 1; len < TN; len <<= 1)
#pragma unroll
                for(int i = 0; i < TN; i += len << 1)
#pragma unroll
                    for(int jj = i; jj < i + len; jj++) {
                        int &p = pass ? m[jj][r] : m[r][jj], &q = pass ? m[jj + len][r] : m[r][jj + len];
                        const int u = p + q, v = p - q;
                        p = u; q = v;
                    }
    int s = abs(m[0][0]) >> 2;
#pragma unroll
    for(int y = 0; y < TN; y++)
#pragma unroll
        for(int x = 0; x < TN; x++)
            if(x | y) s += abs(m[y][x]);
    return TN == 8 ? (s + 2) >> 2 : (s + 1) >> 1;
}
