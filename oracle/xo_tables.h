/* xo_tables.h -- TEST INFRASTRUCTURE ONLY: the oracle's OWN codec constants.
 *
 * The checker must not share code with the checked: the library generates its tables in xeve_b200/csrc/xb200_tables.h, the oracle
 * has these -- written separately, formulated differently where a formulation exists (literal tables instead of step functions,
 * a sort instead of a diagonal walk, long-double trigonometry, bit loops instead of clz) -- and tests/test_oracle.py /
 * tests/test_tables_abi.py check BOTH against the reference's own tables (rh_table: xeve_tbl_tm*, xeve_tbl_scan, xeve_tbl_mv_bits,
 * xeve_tbl_refi_bits, xeve_tbl_df_st, xeve_tbl_mpm; src_base/xeve_tbl.c:40-48, 83-257, 286-517, 625-).  Standard-defined constants
 * (filter taps, quantiser scales) are literals here as they are there. */
#ifndef XO_TABLES_H_
#define XO_TABLES_H_
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* DCT-II basis, 64-point: c(k) * cos((2n + 1) k pi / 128) scaled by 64 (row 0) / 64 sqrt 2, rounded half away from zero */
static void xo_gen_tm64(int8_t *tm)
{
    const long double pi = acosl(-1.0L);
    for(int n = 0; n < 64; n++) tm[n] = 64;
    for(int k = 1; k < 64; k++)
        for(int n = 0; n < 64; n++) tm[k * 64 + n] = (int8_t)lroundl(64.0L * sqrtl(2.0L) * cosl(pi * (long double)((2 * n + 1) * k) / 128.0L));
}

/* zig-zag scan: positions ordered by anti-diagonal d = x + y; within an even diagonal by increasing x, within an odd one by increasing y */
typedef struct { int key; uint16_t pos; } xo_scan_ent;
static int xo_scan_cmp(const void *a, const void *b) { return ((const xo_scan_ent *)a)->key - ((const xo_scan_ent *)b)->key; }
static void xo_gen_scan(uint16_t *scan, int log2w, int log2h)
{
    const int    w = 1 << log2w, h = 1 << log2h;
    xo_scan_ent *e = (xo_scan_ent *)malloc(sizeof(xo_scan_ent) * (size_t)w * h);
    for(int y = 0; y < h; y++)
        for(int x = 0; x < w; x++) {
            const int d = x + y;
            e[y * w + x].key = d * (w + h) + ((d & 1) ? y : x);
            e[y * w + x].pos = (uint16_t)(y * w + x);
        }
    qsort(e, (size_t)w * h, sizeof(xo_scan_ent), xo_scan_cmp);
    for(int i = 0; i < w * h; i++) scan[i] = e[i].pos;
    free(e);
}

/* bits of one MVD component as the reference's table holds them for -2048 < v <= 2048 (signed exp-Golomb length; the table is one
 * entry off centre: its first element, v = -2047, holds 22), escape formula of src_base/xeve_pinter.c:74-92 beyond */
static int xo_mvd_bits(int v)
{
    if(v > 2048 || v <= -2048) {
        unsigned a = (unsigned)abs(v);
        int      nn = (int)((a + 1) >> 12), len = 11;
        for(; len < 16 && nn; nn >>= 1) len++;
        return 2 * len + 2;
    }
    if(v == -2047) return 22;   /* the off-centre entry */
    int      bits = 1;
    for(unsigned a = (unsigned)abs(v) + 1; a > 1; a >>= 1) bits += 2;
    return bits + (v != 0);
}
static int xo_refi_bits(int num_refp, int refi) { return num_refp < 2 ? 0 : (refi < num_refp - 1 ? refi + 1 : num_refp - 1); }
static int xo_mv_bits(int dx, int dy, int num_refp, int refi) { return xo_mvd_bits(dx) + xo_mvd_bits(dy) + xo_refi_bits(num_refp, refi); }

static const int16_t xo_mc_l_taps[4][8] = {{0, 0, 0, 64, 0, 0, 0, 0}, {0, 1, -5, 52, 20, -5, 1, 0}, {0, 2, -10, 40, 40, -10, 2, 0}, {0, 1, -5, 20, 52, -5, 1, 0}};
static const int16_t xo_mc_c_taps[8][4] = {{0, 64, 0, 0}, {-2, 58, 10, -2}, {-4, 52, 20, -4}, {-6, 46, 30, -6}, {-8, 40, 40, -8}, {-6, 30, 46, -6},
                                           {-4, 20, 52, -4}, {-2, 10, 58, -2}};
static const int xo_quant_scale[6] = {26214, 23302, 20560, 18396, 16384, 14764};
static const int xo_dequant_scale[6] = {40, 45, 51, 57, 64, 71};

/* RDOQ error scale, src_base/xeve_tq.c:406-423: the reference evaluates it in doubles in this order */
static int64_t xo_err_scale(int qp_rem, int log2_size, int bit_depth)
{
    const int tr_shift = 15 - bit_depth - log2_size;
    double    e = (double)(1 << 15) * pow(2.0, -tr_shift);
    e = e / xo_quant_scale[qp_rem] / (1 << (bit_depth - 8));
    return (int64_t)(e * (double)(1 << 20));
}

/* deblocking strength by [class][qp], EVC Baseline (classes: intra | luma cbf | motion differs | none) */
static const uint8_t xo_df_st[4][52] = {
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 5, 6, 6, 7, 8, 9, 10, 11, 12, 12, 12, 12, 12},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 11, 11, 11, 11, 11},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 4, 4, 5, 6, 7, 8, 9, 10, 10, 10, 10, 10},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
};
static int xo_df_strength(int cls, int qp) { return xo_df_st[cls > 3 ? 3 : cls][qp < 0 ? 0 : (qp > 51 ? 51 : qp)]; }

/* most-probable-mode symbols [left mode + 1][upper mode + 1][mode], index 0 = neighbour not intra / not available */
static const uint8_t xo_mpm_tbl[6][6][5] = {
    {{0, 2, 3, 1, 4}, {0, 2, 1, 3, 4}, {0, 2, 1, 3, 4}, {1, 2, 0, 3, 4}, {0, 2, 1, 3, 4}, {0, 1, 2, 3, 4}},
    {{1, 0, 2, 3, 4}, {0, 1, 2, 3, 4}, {0, 1, 2, 3, 4}, {1, 2, 0, 3, 4}, {0, 1, 3, 2, 4}, {0, 2, 1, 4, 3}},
    {{1, 0, 2, 3, 4}, {1, 0, 2, 3, 4}, {1, 0, 2, 3, 4}, {2, 0, 1, 3, 4}, {1, 0, 3, 2, 4}, {0, 1, 2, 4, 3}},
    {{1, 0, 2, 3, 4}, {0, 2, 1, 3, 4}, {1, 0, 2, 3, 4}, {1, 2, 0, 3, 4}, {0, 1, 2, 3, 4}, {0, 2, 1, 4, 3}},
    {{0, 1, 2, 3, 4}, {0, 3, 2, 1, 4}, {1, 0, 2, 3, 4}, {1, 2, 0, 3, 4}, {1, 2, 3, 0, 4}, {0, 2, 1, 4, 3}},
    {{0, 1, 2, 3, 4}, {0, 1, 2, 4, 3}, {0, 1, 2, 4, 3}, {0, 2, 1, 4, 3}, {0, 1, 2, 3, 4}, {0, 1, 2, 4, 3}}};
#endif
