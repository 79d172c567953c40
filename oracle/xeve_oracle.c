/*
 * xeve_oracle.c -- TEST INFRASTRUCTURE ONLY: a plain scalar C restatement of the reference's
 * inter-search + transform path, used as the checker in tests/, __graft_entry__.smoke() and the
 * cpu_baseline leg of bench.py.  The shipped library never links or calls this file.
 *
 * Parity of this restatement is PINNED: tests/test_oracle.py checks every function here
 * against the compiled reference itself (oracle/_ref, built by oracle/Makefile.ref from the
 * sources under /root/reference) on random blocks, on work lists traced from real encodes, and --
 * for the decision chain at the end of this file -- on whole sequences up to the bitstream the
 * unmodified reference writes from the chain's decisions; tests/golden/ holds fixtures generated
 * the same way for machines without /root/reference.
 * (The reference ships no tests or golden vectors of its own)
 *
 * Each function cites the reference file:line it follows.  The code is written from the
 * arithmetic, not transcribed: transforms are direct matrix products instead of butterflies,
 * the motion search is expressed as "build the candidate list of this round, then take the
 * first strict minimum", RDOQ as a two-pass scan.
 */
#include "xeve_oracle.h"
#include "xo_tables.h"   /* the oracle's own constants: nothing here includes product code */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static int8_t   g_tm64[64 * 64];
static uint16_t g_scan[7][4096]; /* square blocks, index log2 size */
static int      g_init;

static void oracle_init(void)
{
    if(g_init) return;
    xo_gen_tm64(g_tm64);
    for(int l = 1; l <= 6; l++) xo_gen_scan(g_scan[l], l, l);
    g_init = 1;
}
static inline int tm(int log2n, int k, int n) { return g_tm64[(k << (6 - log2n)) * 64 + n]; }
static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : v > hi ? hi : v; }

/* ---------------------------------------------------------------------------------------------
 * distortion kernels
 * ------------------------------------------------------------------------------------------- */
/* src_base/xeve_sad.c:40-61: sum |a-b| then >> (bd-8).  XEVE_ABS16 (src_base/xeve_util.h:55) is a
 * 16-bit sign-fold applied to the int difference; identical to abs() while |d| < 32768. */
int xo_sad(int w, int h, const int16_t *a, int sa, const int16_t *b, int sb, int bd)
{
    int sum = 0;
    for(int y = 0; y < h; y++, a += sa, b += sb)
        for(int x = 0; x < w; x++) {
            int d = (int)a[x] - (int)b[x];
            sum += (d ^ (d >> 15)) - (d >> 15);
        }
    return sum >> (bd - 8);
}

/* src_base/xeve_sad.c:275-297: each squared difference is shifted before summing */
int64_t xo_ssd(int w, int h, const int16_t *a, int sa, const int16_t *b, int sb, int bd)
{
    const int sh  = (bd - 8) << 1;
    int64_t   sum = 0;
    for(int y = 0; y < h; y++, a += sa, b += sb)
        for(int x = 0; x < w; x++) {
            int d = a[x] - b[x];
            sum += (d * d) >> sh;
        }
    return sum;
}

/* src_base/xeve_sad.c:160-178 */
void xo_diff(int w, int h, const int16_t *a, int sa, const int16_t *b, int sb, int16_t *d, int sd)
{
    for(int y = 0; y < h; y++, a += sa, b += sb, d += sd)
        for(int x = 0; x < w; x++) d[x] = (int16_t)(a[x] - b[x]);
}

/* Hadamard SATD for square blocks (what Baseline reaches): 8x8 tiles when both sides are
 * multiples of 8, else 4x4 tiles.  src_base/xeve_sad.c:417-607 (tiles), 1043-1140 (dispatch).
 * Per tile: sum of |H d H^T| with the DC term >> 2, then (s+2)>>2 (8x8) or (s+1)>>1 (4x4). */
static int had_tile(const int16_t *a, int sa, const int16_t *b, int sb, int n)
{
    int m[8][8], t[8][8];
    for(int y = 0; y < n; y++)
        for(int x = 0; x < n; x++) m[y][x] = a[y * sa + x] - b[y * sb + x];
    /* unnormalised Walsh-Hadamard along rows then columns (order of outputs is irrelevant for
     * the sum of magnitudes; only the all-plus (DC) output is treated specially) */
    for(int pass = 0; pass < 2; pass++) {
        for(int r = 0; r < n; r++) {
            int v[8];
            for(int i = 0; i < n; i++) v[i] = pass ? m[i][r] : m[r][i];
            for(int len = 1; len < n; len <<= 1)
                for(int i = 0; i < n; i += len << 1)
                    for(int j = i; j < i + len; j++) {
                        int p = v[j], q = v[j + len];
                        v[j] = p + q; v[j + len] = p - q;
                    }
            for(int i = 0; i < n; i++) if(pass) t[i][r] = v[i]; else t[r][i] = v[i];
        }
        memcpy(m, t, sizeof(m));
    }
    int s = abs(m[0][0]) >> 2;
    for(int y = 0; y < n; y++)
        for(int x = 0; x < n; x++) if(x | y) s += abs(m[y][x]);
    return n == 8 ? (s + 2) >> 2 : (s + 1) >> 1;
}

int xo_satd(int w, int h, const int16_t *a, int sa, const int16_t *b, int sb, int bd)
{
    int n = (w % 8 == 0 && h % 8 == 0) ? 8 : 4, sum = 0;
    for(int y = 0; y < h; y += n)
        for(int x = 0; x < w; x += n) sum += had_tile(a + y * sa + x, sa, b + y * sb + x, sb, n);
    return sum >> (bd - 8);
}

/* ---------------------------------------------------------------------------------------------
 * motion compensation
 * ------------------------------------------------------------------------------------------- */
#define k_luma xo_mc_l_taps
#define k_chroma xo_mc_c_taps

/* Generic separable interpolation, `taps` taps (8 luma / 4 chroma).
 * 1-D cases: sum >> 6 with no rounding offset (src_base/xeve_mc.h:36-48, xeve_mc.c:122-182);
 * 2-D: horizontal pass >> min(4, bd-8) into s16, vertical pass (sum + 2^(s2-1)) >> s2 with
 * s2 = max(8, 20-bd) (src_base/xeve_mc.c:184-254, 337-381); result clipped to [0, 2^bd-1].
 * fx/fy: filter phase selected from the sub-sample bits of the (already clipped) position;
 * want_h/want_v: which passes run -- chosen by the caller from the UNCLIPPED mv (quirk q9). */
static void interp(const int16_t *ref, int sr, int ix, int iy, const int16_t *ch, const int16_t *cv, int taps,
                   int want_h, int want_v, int16_t *pred, int sp, int w, int h, int bd)
{
    const int maxv = (1 << bd) - 1, half = taps / 2 - 1;
    if(!want_h && !want_v) {
        for(int y = 0; y < h; y++) memcpy(pred + y * sp, ref + (iy + y) * sr + ix, sizeof(int16_t) * w);
        return;
    }
    if(want_h && !want_v) {
        for(int y = 0; y < h; y++)
            for(int x = 0; x < w; x++) {
                const int16_t *p = ref + (iy + y) * sr + ix + x - half;
                int acc = 0;
                for(int t = 0; t < taps; t++) acc += ch[t] * p[t];
                pred[y * sp + x] = (int16_t)clip3(0, maxv, acc >> 6);
            }
        return;
    }
    if(!want_h && want_v) {
        for(int y = 0; y < h; y++)
            for(int x = 0; x < w; x++) {
                const int16_t *p = ref + (iy + y - half) * sr + ix + x;
                int acc = 0;
                for(int t = 0; t < taps; t++) acc += cv[t] * p[t * sr];
                pred[y * sp + x] = (int16_t)clip3(0, maxv, acc >> 6);
            }
        return;
    }
    const int s1 = bd - 8 < 4 ? bd - 8 : 4, s2 = 20 - bd > 8 ? 20 - bd : 8;
    int16_t  *tmp = malloc(sizeof(int16_t) * w * (h + taps - 1));
    for(int y = 0; y < h + taps - 1; y++)
        for(int x = 0; x < w; x++) {
            const int16_t *p = ref + (iy + y - half) * sr + ix + x - half;
            int acc = 0;
            for(int t = 0; t < taps; t++) acc += ch[t] * p[t];
            tmp[y * w + x] = (int16_t)(acc >> s1);
        }
    for(int y = 0; y < h; y++)
        for(int x = 0; x < w; x++) {
            int acc = 0;
            for(int t = 0; t < taps; t++) acc += cv[t] * tmp[(y + t) * w + x];
            pred[y * sp + x] = (int16_t)clip3(0, maxv, (acc + (1 << (s2 - 1))) >> s2);
        }
    free(tmp);
}

/* luma block at absolute position (gx, gy) in quarter-pel; sel_x/sel_y = the mv the filter
 * choice is made from (xeve_mc_l macro, src_base/xeve_mc.h:94-97) */
void xo_mc_luma(const int16_t *ref, int sr, int gx, int gy, int sel_x, int sel_y, int16_t *pred, int sp, int w, int h,
                int bd)
{
    interp(ref, sr, gx >> 2, gy >> 2, k_luma[gx & 3], k_luma[gy & 3], 8, (sel_x & 3) != 0, (sel_y & 3) != 0, pred, sp,
           w, h, bd);
}
/* chroma (4:2:0): position in eighth-pel of the chroma grid = the luma quarter-pel value */
void xo_mc_chroma(const int16_t *ref, int sr, int gx, int gy, int sel_x, int sel_y, int16_t *pred, int sp, int w,
                  int h, int bd)
{
    interp(ref, sr, gx >> 3, gy >> 3, k_chroma[gx & 7], k_chroma[gy & 7], 4, (sel_x & 7) != 0, (sel_y & 7) != 0, pred,
           sp, w, h, bd);
}

/* src_base/xeve_mc.c:401-447 (xeve_mv_clip) + 465-610 (xeve_mc): uni/bi prediction of Y,U,V.
 * pred: Y (w*h) | U (w/2*h/2) | V, stride = block width. */
void xo_mc(const xb200_seq *sq, const xo_planes *pl, const xb200_mc_item *it, int16_t *pred)
{
    const int w = it->w, h = it->h, x = it->x, y = it->y, bd = sq->bit_depth;
    const int ny = w * h, nc = ny / 4;
    int16_t  *buf[2] = {pred, NULL};
    int       mvt[2][2], n = 0;
    for(int l = 0; l < 2; l++) {
        mvt[l][0] = it->mv[l][0]; mvt[l][1] = it->mv[l][1];
        if(it->refi[l] < 0) continue;
        const int lo = -(128 << 2), hx = (sq->w - 1 + 128) << 2, hy = (sq->h - 1 + 128) << 2;
        if((x << 2) + it->mv[l][0] < lo) mvt[l][0] = lo - (x << 2);
        if((y << 2) + it->mv[l][1] < lo) mvt[l][1] = lo - (y << 2);
        if((x << 2) + it->mv[l][0] + (w << 2) - 4 > hx) mvt[l][0] = hx - (x << 2) - (w << 2) + 4;
        if((y << 2) + it->mv[l][1] + (h << 2) - 4 > hy) mvt[l][1] = hy - (y << 2) - (h << 2) + 4;
        mvt[l][0] = (int16_t)mvt[l][0]; mvt[l][1] = (int16_t)mvt[l][1];
    }
    for(int l = 0; l < 2; l++) {
        if(it->refi[l] < 0) continue;
        if(l == 1 && it->refi[0] >= 0) {
            /* identical-motion shortcut, src_base/xeve_mc.c:545-551 */
            if(it->ref_poc[0] == it->ref_poc[1] && mvt[0][0] == mvt[1][0] && mvt[0][1] == mvt[1][1]) return;
        }
        const xo_planes *r = &pl[it->ref_pic[l]];
        int16_t *dst = n == 0 ? pred : (buf[1] = malloc(sizeof(int16_t) * (ny + 2 * nc)));
        int gx = (x << 2) + mvt[l][0], gy = (y << 2) + mvt[l][1];
        xo_mc_luma(r->y, r->s_l, gx, gy, it->mv[l][0], it->mv[l][1], dst, w, w, h, bd);
        xo_mc_chroma(r->u, r->s_c, gx, gy, it->mv[l][0], it->mv[l][1], dst + ny, w / 2, w / 2, h / 2, bd);
        xo_mc_chroma(r->v, r->s_c, gx, gy, it->mv[l][0], it->mv[l][1], dst + ny + nc, w / 2, w / 2, h / 2, bd);
        n++;
    }
    if(n == 2) { /* src_base/xeve_mc.c:449-463: (a + b + 1) >> 1, no clip */
        for(int i = 0; i < ny + 2 * nc; i++) pred[i] = (int16_t)((pred[i] + buf[1][i] + 1) >> 1);
        free(buf[1]);
    }
}

/* ---------------------------------------------------------------------------------------------
 * transforms
 * ------------------------------------------------------------------------------------------- */
/* src_base/xeve_tq.c:396-404 + tx_pb*b: stage 0 exact (rows, shift 0), stage 1 (columns) rounded by
 * shift = (log2w - 1 + bd - 8) + (log2h + 6) (src_base/xeve_util.c:34-35), result truncated to s16.
 * The 64-point transform only produces outputs 0..31 (src_base/xeve_tq.c:318-381). */
void xo_fwd_transform(int16_t *blk, int log2w, int log2h, int bd)
{
    oracle_init();
    const int w = 1 << log2w, h = 1 << log2h;
    const int shift = (log2w - 1 + bd - 8) + (log2h + 6);
    const int kw = w == 64 ? 32 : w, kh = h == 64 ? 32 : h;
    int32_t  *t = calloc((size_t)w * h, sizeof(int32_t)); /* t[u][y] */
    for(int y = 0; y < h; y++)
        for(int u = 0; u < kw; u++) {
            int64_t acc = 0;
            for(int x = 0; x < w; x++) acc += (int64_t)tm(log2w, u, x) * blk[y * w + x];
            t[u * h + y] = (int32_t)acc;
        }
    for(int u = 0; u < w; u++)
        for(int v = 0; v < h; v++) {
            int64_t acc = 0;
            if(v < kh) {
                for(int y = 0; y < h; y++) acc += (int64_t)tm(log2h, v, y) * t[u * h + y];
                acc = (acc + ((int64_t)1 << (shift - 1))) >> shift;
            }
            blk[v * w + u] = (int16_t)acc;
        }
    free(t);
}

/* src_base/xeve_itdq.c:435-440 + xeve_itx_pb*b: columns first (shift 0, clip to s32), then rows
 * with shift 7 + 12 - (bd - 8), clip to s16. */
void xo_inv_transform(int16_t *blk, int log2w, int log2h, int bd)
{
    oracle_init();
    const int w = 1 << log2w, h = 1 << log2h, shift = 7 + 12 - (bd - 8);
    int32_t  *t = malloc(sizeof(int32_t) * w * h); /* t[u][y] */
    for(int u = 0; u < w; u++)
        for(int y = 0; y < h; y++) {
            int64_t acc = 0;
            for(int v = 0; v < h; v++) acc += (int64_t)tm(log2h, v, y) * blk[v * w + u];
            t[u * h + y] = acc <= INT32_MIN ? INT32_MIN : acc >= INT32_MAX ? INT32_MAX : (int32_t)acc;
        }
    for(int y = 0; y < h; y++)
        for(int x = 0; x < w; x++) {
            int64_t acc = 0;
            for(int u = 0; u < w; u++) acc += (int64_t)tm(log2w, u, x) * t[u * h + y];
            acc = (acc + ((int64_t)1 << (shift - 1))) >> shift;
            blk[y * w + x] = (int16_t)(acc < -32768 ? -32768 : acc > 32767 ? 32767 : acc);
        }
    free(t);
}

/* ---------------------------------------------------------------------------------------------
 * Main profile, first piece: the "IQT" transforms, two 16-bit stages (sps.tool_iqt = 1).
 * Forward: xeve_trans with iqt_flag (src_main/xevem_tq.c:709-716) over tx_pb2 .. tx_pb64 (:58-334): rows with
 * shift log2w - 1 + bd - 8, then columns with shift log2h + 6 (src_base/xeve_util.c:1348-1351), each stage rounded and
 * stored as s16 (plain truncation); the 64-point stage leaves outputs 32..63 zero.  Inverse: xeve_itrans
 * (src_main/xevem_itdq.c:551-557) over itx_pb2 .. itx_pb64 (:302-549): columns with shift 7, rows with shift 12 - (bd - 8),
 * each stage clipped to s16.  Blocks may be non-square (binary / ternary splits).  The butterflies are exact integer
 * arithmetic, so each stage is a matrix product with the transform matrix the Baseline path uses.
 * ------------------------------------------------------------------------------------------- */
static void iqt_stage_fwd(const int16_t *src, int16_t *dst, int log2n, int shift, int line)   /* tx_pbN: dst[k * line + j] */
{
    const int n = 1 << log2n, kmax = n == 64 ? 32 : n, add = shift == 0 ? 0 : 1 << (shift - 1);
    for(int j = 0; j < line; j++)
        for(int k = 0; k < n; k++) {
            int acc = 0;
            if(k < kmax) {
                for(int x = 0; x < n; x++) acc += tm(log2n, k, x) * src[j * n + x];
                acc = (acc + add) >> shift;
            }
            dst[k * line + j] = (int16_t)acc;
        }
}
static void iqt_stage_inv(const int16_t *src, int16_t *dst, int log2n, int shift, int line)   /* itx_pbN: dst[j * n + x] */
{
    const int n = 1 << log2n, kmax = n == 64 ? 32 : n, add = shift == 0 ? 0 : 1 << (shift - 1);
    for(int j = 0; j < line; j++)
        for(int x = 0; x < n; x++) {
            int acc = 0;
            for(int k = 0; k < kmax; k++) acc += tm(log2n, k, x) * src[k * line + j];
            acc = (acc + add) >> shift;
            dst[j * n + x] = (int16_t)(acc < -32768 ? -32768 : acc > 32767 ? 32767 : acc);
        }
}
void xo_iqt_fwd(int16_t *blk, int log2w, int log2h, int bd)
{
    oracle_init();
    int16_t *t = malloc(sizeof(int16_t) << (log2w + log2h));
    iqt_stage_fwd(blk, t, log2w, log2w - 1 + bd - 8, 1 << log2h);
    iqt_stage_fwd(t, blk, log2h, log2h + 6, 1 << log2w);
    free(t);
}
void xo_iqt_inv(int16_t *blk, int log2w, int log2h, int bd)
{
    oracle_init();
    int16_t *t = malloc(sizeof(int16_t) << (log2w + log2h));
    iqt_stage_inv(blk, t, log2h, 7, 1 << log2w);
    iqt_stage_inv(t, blk, log2w, 12 - (bd - 8), 1 << log2h);
    free(t);
}

/* ATS, the adaptive transform selection of the Main profile: DST-VII / DCT-VIII, 4..32 points, chosen per direction by
 * ats_intra_tridx (horizontal = bit 1, vertical = bit 0; 0 -> DST-VII, 1 -> DCT-VIII, src_main/xevem_tbl.c:419).  Forward
 * xeve_t_MxN_ats_intra (src_main/xevem_tq.c:684-702, stages :336-682), inverse xeve_it_MxN_ats_intra (src_main/xevem_itdq.c:278-300,
 * stages :63-276): the same two 16-bit stages and shifts as the IQT transforms above.  The 8-bit matrices are the rounded basis
 * functions scaled by 64 sqrt(N): c(k, n) = round(64 sqrt(N) sqrt(4 / (2N + 1)) f), f = cos(pi (2k+1)(2n+1) / (4N+2)) for DCT-VIII and
 * sin(pi (2k+1)(n+1) / (2N+1)) for DST-VII (checked against the reference's table in tests/test_oracle.py). */
static int8_t g_ats[2][4][32 * 32];          /* [0 DCT-VIII | 1 DST-VII][log2n - 2][k * n_ + n] */
static int    g_ats_init;
void xo_ats_matrix(int type, int log2n, int8_t *out)
{
    const int    n_ = 1 << log2n;
    const double pi = 3.14159265358979323846, scale = 64.0 * sqrt((double)n_) * sqrt(4.0 / (2 * n_ + 1));
    for(int k = 0; k < n_; k++)
        for(int n = 0; n < n_; n++) {
            const double f = type == 0 ? cos(pi * (2 * k + 1) * (2 * n + 1) / (4 * n_ + 2)) : sin(pi * (2 * k + 1) * (n + 1) / (2 * n_ + 1));
            out[k * n_ + n] = (int8_t)floor(scale * f + 0.5);
        }
}
static const int8_t *ats_tm(int type, int log2n)
{
    if(!g_ats_init) {
        for(int t = 0; t < 2; t++) for(int l = 2; l <= 5; l++) xo_ats_matrix(t, l, g_ats[t][l - 2]);
        g_ats_init = 1;
    }
    return g_ats[type][log2n - 2];
}
static void ats_stage_fwd(const int16_t *src, int16_t *dst, const int8_t *m, int log2n, int shift, int line)
{
    const int n = 1 << log2n, add = 1 << (shift - 1);
    for(int j = 0; j < line; j++)
        for(int k = 0; k < n; k++) {
            int acc = 0;
            for(int x = 0; x < n; x++) acc += m[k * n + x] * src[j * n + x];
            dst[k * line + j] = (int16_t)((acc + add) >> shift);
        }
}
static void ats_stage_inv(const int16_t *src, int16_t *dst, const int8_t *m, int log2n, int shift, int line)
{
    const int n = 1 << log2n, add = 1 << (shift - 1);
    for(int j = 0; j < line; j++)
        for(int x = 0; x < n; x++) {
            int acc = 0;
            for(int k = 0; k < n; k++) acc += m[k * n + x] * src[k * line + j];
            acc = (acc + add) >> shift;
            dst[j * n + x] = (int16_t)(acc < -32768 ? -32768 : acc > 32767 ? 32767 : acc);
        }
}
void xo_ats_fwd(int16_t *blk, int log2w, int log2h, int bd, int tridx)
{
    int16_t *t = malloc(sizeof(int16_t) << (log2w + log2h));
    ats_stage_fwd(blk, t, ats_tm((tridx >> 1) ? 0 : 1, log2w), log2w, log2w - 1 + bd - 8, 1 << log2h);
    ats_stage_fwd(t, blk, ats_tm((tridx & 1) ? 0 : 1, log2h), log2h, log2h + 6, 1 << log2w);
    free(t);
}
void xo_ats_inv(int16_t *blk, int log2w, int log2h, int bd, int tridx)
{
    int16_t *t = malloc(sizeof(int16_t) << (log2w + log2h));
    ats_stage_inv(blk, t, ats_tm((tridx & 1) ? 0 : 1, log2h), log2h, 7, 1 << log2w);
    ats_stage_inv(t, blk, ats_tm((tridx >> 1) ? 0 : 1, log2w), log2w, 12 - (bd - 8), 1 << log2h);
    free(t);
}

/* ---------------------------------------------------------------------------------------------
 * quantisation with RDOQ (square blocks), src_base/xeve_tq.c:425-730
 * ------------------------------------------------------------------------------------------- */
typedef struct { const xb200_rates *r; int ctx; int64_t lambda; } rate_env;

/* get_ic_rate_cost_rl, src_base/xeve_tq.c:425-457 */
static int64_t level_rate(const rate_env *e, uint32_t lev, int run_is_zero)
{
    int32_t        rate;
    const int32_t *rr = e->r->run[e->ctx + (run_is_zero ? 0 : 1)];
    if(lev == 0) rate = rr[1];
    else {
        rate = 32768 + rr[0];
        if(lev == 1) rate += e->r->level[e->ctx][0];
        else rate += e->r->level[e->ctx][1] + e->r->level[e->ctx + 1][1] * (int32_t)(lev - 2) + e->r->level[e->ctx + 1][0];
    }
    return (int64_t)rate * e->lambda;
}

int xo_quant_rdoq(int16_t *coef, int log2n, int qp, double d_lambda, int is_intra, int ch, int slice_type,
                  const xb200_rates *rt, int bd)
{
    oracle_init();
    const int *qs = xo_quant_scale;
    const int        n = 1 << (2 * log2n), q = qs[qp % 6];
    const int        qbits = 14 + (15 - bd - log2n) + qp / 6;
    /* zero-block pre-test, src_base/xeve_tq.c:666-700 (slice_type: 2 == SLICE_I in the reference) */
    {
        int64_t off = (int64_t)(slice_type == 2 ? 201 : 153) << (qbits - 9), thr = ((int64_t)1 << qbits) - off;
        int     coded = 0;
        for(int i = 0; i < n && !coded; i++) coded = (int64_t)abs(coef[i]) * q >= thr;
        if(!coded) { memset(coef, 0, sizeof(int16_t) * n); return 0; }
    }
    const uint16_t *scan = g_scan[log2n];
    const int64_t   es = xo_err_scale(qp % 6, log2n, bd);
    rate_env        e  = {rt, ch == 0 ? 0 : 2, (int64_t)(d_lambda * 32768.0 + 0.5)};
    int64_t *ld   = malloc(sizeof(int64_t) * n);
    int32_t *maxl = malloc(sizeof(int32_t) * n);
    int16_t *out  = calloc(n, sizeof(int16_t));
    int64_t  uncoded_blk = 0;
    long     sum_all = 0;
    for(int i = 0; i < n; i++) { /* position order is irrelevant for this pass */
        int64_t t = (int64_t)abs(coef[i]) * q, cap = (int64_t)INT32_MAX - ((int64_t)1 << (qbits - 1));
        int64_t v = (int32_t)(t < cap ? t : cap);
        ld[i]     = v;
        uint32_t m = (uint32_t)(v >> qbits);
        if(v - ((int64_t)m << qbits) >= ((int64_t)1 << (qbits - 1))) m++;
        maxl[i]  = (int32_t)m;
        int64_t err = (v * es) >> 20;
        uncoded_blk += err * err;
        sum_all += m;
    }
    int nnz = 0;
    if(sum_all) {
        const int32_t *cbf = (!is_intra && ch == 0) ? rt->cbf_all : ch == 0 ? rt->cbf_luma : ch == 1 ? rt->cbf_cb : rt->cbf_cr;
        int64_t best = uncoded_blk + (int64_t)cbf[0] * e.lambda, base = uncoded_blk + (int64_t)cbf[1] * e.lambda;
        const int32_t *last = rt->last[ch == 0 ? 0 : 1];
        int run = 0, best_last = 0;
        for(int sp = 0; sp < n; sp++) {
            int      p = scan[sp];
            /* get_coded_level_rl, src_base/xeve_tq.c:459-489: candidates max, max-1 (>= 1), else 0 */
            int64_t  err0 = (ld[p] * es) >> 20, unc = err0 * err0;
            int64_t  cod = unc + level_rate(&e, 0, run == 0);
            uint32_t lev = 0, mx = (uint32_t)maxl[p], mn = mx > 1 ? mx - 1 : 1;
            for(uint32_t a = mx; a >= mn; a--) {
                int64_t d = ld[p] - ((int64_t)a << qbits), er = (d * es) >> 20;
                int64_t c = er * er + level_rate(&e, a, run == 0);
                if(c < cod) { lev = a; cod = c; }
            }
            out[p] = (int16_t)((coef[p] > 0 ? maxl[p] : -maxl[p]) < 0 ? -(int32_t)lev : (int32_t)lev);
            base += cod - unc;
            if(lev) {
                int64_t as_last = base + (int64_t)last[1] * e.lambda;
                base += (int64_t)last[0] * e.lambda;
                if(as_last < best) { best = as_last; best_last = sp + 1; }
                run = 0;
            }
            else run++;
        }
        for(int sp = 0; sp < n; sp++) {
            int p = scan[sp];
            if(sp < best_last) nnz += out[p] != 0; else out[p] = 0;
        }
    }
    memcpy(coef, out, sizeof(int16_t) * n);
    free(ld); free(maxl); free(out);
    return nnz;
}

/* plain quantiser (rdoq == 0), src_base/xeve_tq.c:704-727 */
int xo_quant_plain(int16_t *coef, int log2n, int qp, int slice_type, int bd)
{
    const int *qs = xo_quant_scale;
    const int n = 1 << (2 * log2n), shift = 14 + (15 - bd - log2n) + qp / 6;
    const int32_t off = (int32_t)(slice_type == 2 ? 171 : 85) << (shift - 9);
    int       nnz = 0;
    for(int i = 0; i < n; i++) {
        int32_t lev = (int16_t)(((int32_t)abs(coef[i]) * qs[qp % 6] + off) >> shift);
        coef[i]     = (int16_t)(coef[i] < 0 ? -lev : lev);
        nnz += coef[i] != 0;
    }
    return nnz;
}

/* xeve_sub_block_tq for CU <= 64 (one transform block per plane), src_base/xeve_tq.c:750-864.
 * planes: Y | U | V contiguous; returns nnz per plane. */
void xo_tq(const xb200_seq *sq, const xb200_tq_item *it, const xb200_rates *rates, int16_t *planes, int nnz[3])
{
    const int ny = 1 << (it->log2_cuw + it->log2_cuh), nc = ny >> 2;
    int16_t  *p[3] = {planes, planes + ny, planes + ny + nc};
    for(int c = 0; c < 3; c++) {
        nnz[c] = 0;
        if(!((it->run_stats >> c) & 1)) continue;
        int l2 = c ? it->log2_cuw - 1 : it->log2_cuw;
        xo_fwd_transform(p[c], l2, l2, sq->bit_depth);
        nnz[c] = sq->rdoq ? xo_quant_rdoq(p[c], l2, it->qp[c], it->lambda[c], it->is_intra, c, it->slice_type,
                                          &rates[it->rate_idx], sq->bit_depth)
                          : xo_quant_plain(p[c], l2, it->qp[c], it->slice_type, sq->bit_depth);
    }
}

/* xeve_itdq for CU <= 64, src_base/xeve_itdq.c:442-580: dequantise (scale << qp/6, rounded shift,
 * clip s16) then inverse transform, for the planes with nnz != 0 */
void xo_itdq(const xb200_seq *sq, const xb200_tq_item *it, int16_t *planes, const int nnz[3])
{
    const int *dq = xo_dequant_scale;
    const int ny = 1 << (it->log2_cuw + it->log2_cuh), nc = ny >> 2;
    int16_t  *p[3] = {planes, planes + ny, planes + ny + nc};
    for(int c = 0; c < 3; c++) {
        if(!nnz[c]) continue;
        int     l2 = c ? it->log2_cuw - 1 : it->log2_cuw, n = 1 << (2 * l2);
        int     shift = 20 - 14 - (15 - sq->bit_depth - l2);
        int64_t scale = (int64_t)dq[it->qp[c] % 6] << (it->qp[c] / 6), off = shift ? 1 << (shift - 1) : 0;
        for(int i = 0; i < n; i++) {
            int64_t v = (p[c][i] * scale + off) >> shift;
            p[c][i]   = (int16_t)(v < -32768 ? -32768 : v > 32767 ? 32767 : v);
        }
        xo_inv_transform(p[c], l2, l2, sq->bit_depth);
    }
}

/* src_base/xeve_recon.c:35-57; the sum is formed in s16 */
void xo_recon(const int16_t *resi, const int16_t *pred, int is_coef, int n, int16_t *rec, int bd)
{
    const int maxv = (1 << bd) - 1;
    for(int i = 0; i < n; i++) {
        int16_t t = is_coef ? (int16_t)(resi[i] + pred[i]) : pred[i];
        rec[i]    = (int16_t)clip3(0, maxv, t);
    }
}

/* ---------------------------------------------------------------------------------------------
 * integer + sub-pel motion search, src_base/xeve_pinter.c:122-140, 363-869
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const xb200_seq *sq;
    const int16_t   *org;  int so;   /* original block or the 2*org - pred block of bi search */
    const int16_t   *ref;  int sr;   /* reference luma, active-area origin */
    int      x, y, w, h, bi, other_bits, num_refp, refi;
    uint32_t lambda_mv;
    int      gmvp[2], lo[2], hi[2];  /* predictor in frame coordinates; current search window */
    int      dyn_range;              /* distance-scaled search range */
    int      static_range;           /* pi->max_search_range */
} search_t;

typedef struct { int x, y; } pt;

static void set_window(search_t *s, int cx, int cy, int bi_mode)
{
    int r = bi_mode ? 5 : s->dyn_range;
    s->lo[0] = clip3(s->sq->min_clip[0], s->sq->max_clip[0], cx - r);
    s->hi[0] = clip3(s->sq->min_clip[0], s->sq->max_clip[0], cx + r);
    s->lo[1] = clip3(s->sq->min_clip[1], s->sq->max_clip[1], cy - r);
    s->hi[1] = clip3(s->sq->min_clip[1], s->sq->max_clip[1], cy + r);
}

static uint32_t mv_cost(const search_t *s, int qx, int qy, int *bits_out)
{
    int bits = xo_mv_bits(qx - s->gmvp[0], qy - s->gmvp[1], s->num_refp, s->refi);
    if(s->bi) bits += s->other_bits;
    *bits_out = bits;
    return (uint32_t)((s->lambda_mv * (uint32_t)bits + (1u << 15)) >> 16);
}

static uint32_t int_cost(const search_t *s, int px, int py, int *bits)
{
    if(px < s->lo[0] || px > s->hi[0] || py < s->lo[1] || py > s->hi[1]) return UINT32_MAX;
    uint32_t c   = mv_cost(s, px << 2, py << 2, bits);
    int      sad = xo_sad(s->w, s->h, s->org, s->so, s->ref + py * s->sr + px, s->sr, s->sq->bit_depth);
    return c + (uint32_t)(s->bi ? sad >> 1 : sad);
}

/* one me_ipel_diamond run; returns best cost, writes best integer position, the step at which
 * the best was found (only when something was found) and the bit count of the winner */
static uint32_t diamond(search_t *s, int sx, int sy, int patience, pt *best_out, int *found_step, int *best_bits)
{
    static const int8_t d8[9][2]   = {{-2, 0}, {-1, 1}, {0, 2}, {1, 1}, {2, 0}, {1, -1}, {0, -2}, {-1, -1}, {0, 0}};
    static const int8_t d16[16][2] = {{-4, 0}, {-3, 1}, {-2, 2}, {-1, 3}, {0, 4}, {1, 3}, {2, 2}, {3, 1},
                                      {4, 0}, {3, -1}, {2, -2}, {1, -3}, {0, -4}, {-1, -3}, {-2, -2}, {-3, -1}};
    pt       c0 = {clip3(s->sq->min_clip[0], s->sq->max_clip[0], sx), clip3(s->sq->min_clip[1], s->sq->max_clip[1], sy)};
    pt       best = c0;
    uint32_t best_cost = UINT32_MAX;
    int      misses = 0, step = 0;
    *best_bits = 0;
    for(;;) {
        pt  cand[128];
        int n = 0, this_step;
        misses++;
        if(step <= 2) { /* dense window around the running best */
            int r  = s->bi == 1 ? 5 : 2;
            int x0 = best.x <= s->lo[0] ? best.x : best.x - r, x1 = best.x >= s->hi[0] ? best.x : best.x + r;
            int y0 = best.y <= s->lo[1] ? best.y : best.y - r, y1 = best.y >= s->hi[1] ? best.y : best.y + r;
            for(int yy = y0; yy <= y1; yy++)
                for(int xx = x0; xx <= x1; xx++) cand[n++] = (pt){xx, yy};
            this_step = 2;
        }
        else if(step <= 8) { /* 4 axis points at step 4, 8-point diamond at step 8, plus the centre */
            for(int i = 0; i < 9; i++) {
                if(step == 4 && (i & 1)) continue;
                cand[n++] = (pt){c0.x + (step >> 1) * d8[i][0], c0.y + (step >> 1) * d8[i][1]};
            }
            this_step = step;
        }
        else {
            for(int i = 0; i < 16; i++) cand[n++] = (pt){c0.x + (step >> 2) * d16[i][0], c0.y + (step >> 2) * d16[i][1]};
            this_step = step;
        }
        for(int i = 0; i < n; i++) {
            int      bits;
            uint32_t c = int_cost(s, (int16_t)cand[i].x, (int16_t)cand[i].y, &bits);
            if(c < best_cost) {
                best_cost = c; best = (pt){(int16_t)cand[i].x, (int16_t)cand[i].y};
                *found_step = this_step; *best_bits = bits; misses = 0;
            }
        }
        if(step <= 2) { /* the window is re-centred on the new best (quirk q6) */
            set_window(s, best.x, best.y, s->bi == 1);
            step += 2;
        }
        if(misses == patience || s->bi == 1) break;
        step <<= 1;
        if(step > s->static_range) break;
    }
    *best_out = best;
    return best_cost;
}

static uint32_t subpel(const search_t *s, const int16_t mvi[2], int16_t mv[2], int *win_bits)
{
    static const int8_t hp[8][2] = {{-2, 0}, {-2, 2}, {0, 2}, {2, 2}, {2, 0}, {2, -2}, {0, -2}, {-2, -2}};
    static const int8_t qp[8][2] = {{-1, 0}, {0, 1}, {1, 0}, {0, -1}, {-1, 1}, {1, 1}, {-1, -1}, {1, -1}};
    int16_t *pred = malloc(sizeof(int16_t) * s->w * s->h);
    uint32_t best = UINT32_MAX;
    *win_bits = 0;
    mv[0] = mvi[0]; mv[1] = mvi[1];
    for(int stage = 0; stage < 2; stage++) {
        if(stage == 1 && s->sq->me_level <= 2) break;
        int cx = (int16_t)(mv[0] + (s->x << 2)), cy = (int16_t)(mv[1] + (s->y << 2));
        int cnt = stage ? s->sq->qpel_cnt : s->sq->hpel_cnt;
        for(int i = 0; i < cnt; i++) {
            int qx = (int16_t)(cx + (stage ? qp[i][0] : hp[i][0])), qy = (int16_t)(cy + (stage ? qp[i][1] : hp[i][1]));
            int      bits;
            uint32_t c = mv_cost(s, qx, qy, &bits);
            xo_mc_luma(s->ref, s->sr, qx, qy, qx, qy, pred, s->w, s->w, s->h, s->sq->bit_depth);
            int sad = xo_sad(s->w, s->h, s->org, s->so, pred, s->w, s->sq->bit_depth);
            c += (uint32_t)(s->bi ? sad >> 1 : sad);
            if(c < best) {
                best = c; mv[0] = (int16_t)(qx - (s->x << 2)); mv[1] = (int16_t)(qy - (s->y << 2));
                if(stage) *win_bits = bits; /* the half-pel loop does not record its bits (quirk q3) */
            }
        }
    }
    free(pred);
    return best;
}

void xo_me(const xb200_seq *sq, const xo_planes *pl, const int16_t *side, xb200_me_item *it)
{
    search_t s;
    memset(&s, 0, sizeof(s));
    const xo_planes *cur = &pl[it->cur_pic], *ref = &pl[it->ref_pic];
    s.sq = sq; s.x = it->x; s.y = it->y; s.w = 1 << it->log2_cuw; s.h = 1 << it->log2_cuh; s.bi = it->bi;
    if(it->bi) { s.org = side + it->org_bi_off; s.so = s.w; }
    else { s.org = cur->y + it->y * cur->s_l + it->x; s.so = cur->s_l; }
    s.ref = ref->y; s.sr = ref->s_l;
    s.lambda_mv = it->lambda_mv; s.num_refp = it->num_refp; s.refi = it->refi;
    s.other_bits = it->mot_bits_in[it->lidx ? 0 : 1];
    s.static_range = it->max_search_range;
    {   /* get_range_ipel, src_base/xeve_pinter.c:122-140 */
        int d = it->poc - it->ref_poc; if(d < 0) d = -d;
        s.dyn_range = clip3(it->max_search_range >> 2, it->max_search_range,
                            (it->max_search_range * d + (it->gop_size >> 1)) / it->gop_size);
    }
    s.gmvp[0] = (int16_t)(it->mvp[0] + (it->x << 2)); s.gmvp[1] = (int16_t)(it->mvp[1] + (it->y << 2));
    int mot_bits[2] = {it->mot_bits_in[0], it->mot_bits_in[1]};
    int16_t mv[2] = {it->mv_in[0], it->mv_in[1]};
    const int16_t *start = it->bi == 1 ? mv : it->mvp;
    int cx = clip3(sq->min_clip[0], sq->max_clip[0], it->x + (start[0] >> 2));
    int cy = clip3(sq->min_clip[1], sq->max_clip[1], it->y + (start[1] >> 2));
    set_window(&s, cx, cy, it->bi == 1);

    uint32_t best = UINT32_MAX, c;
    int      found = 0, beststep = 0, bits;
    pt       b;
    c = diamond(&s, (int16_t)(start[0] + (it->x << 2)) >> 2, (int16_t)(start[1] + (it->y << 2)) >> 2, 3, &b, &found, &bits);
    if(it->bi != 1 && bits > 0) mot_bits[it->lidx] = bits;
    if(c < best) {
        best = c; mv[0] = (int16_t)((b.x - it->x) << 2); mv[1] = (int16_t)((b.y - it->y) << 2);
        beststep = (abs(it->mvp[0] - mv[0]) < 2 && abs(it->mvp[1] - mv[1]) < 2) ? 0 : found;
    }
    /* me_raster (me_algo > 1, placebo only) is not part of the fast/medium presets: not restated */
    while(it->bi != 1 && beststep > 0 && sq->me_complexity > 0) {
        set_window(&s, it->x + (mv[0] >> 2), it->y + (mv[1] >> 2), 0);
        beststep = 0;
        c = diamond(&s, (int16_t)(mv[0] + (it->x << 2)) >> 2, (int16_t)(mv[1] + (it->y << 2)) >> 2, 2, &b, &found, &bits);
        if(bits > 0) mot_bits[it->lidx] = bits;
        if(c < best) {
            best = c; mv[0] = (int16_t)((b.x - it->x) << 2); mv[1] = (int16_t)((b.y - it->y) << 2);
            beststep = (abs(it->mvp[0] - mv[0]) < 2 && abs(it->mvp[1] - mv[1]) < 2) ? 0 : found;
        }
    }
    if(sq->me_level > 1) {
        int16_t t[2];
        c = subpel(&s, mv, t, &bits);
        if(!it->bi && bits > 0) mot_bits[it->lidx] = bits;
        if(c < best) { best = c; mv[0] = t[0]; mv[1] = t[1]; }
    }
    else { /* me_ipel_refinement, src_base/xeve_pinter.c:272-361: 3x3 around the best */
        static const int8_t o9[9][2] = {{0, 0}, {-1, -1}, {-1, 0}, {-1, 1}, {0, -1}, {0, 1}, {1, -1}, {1, 0}, {1, 1}};
        set_window(&s, it->x + (mv[0] >> 2), it->y + (mv[1] >> 2), it->bi == 1);
        int ix = clip3(sq->min_clip[0], sq->max_clip[0], (int16_t)(mv[0] + (it->x << 2)) >> 2);
        int iy = clip3(sq->min_clip[1], sq->max_clip[1], (int16_t)(mv[1] + (it->y << 2)) >> 2);
        uint32_t rb = UINT32_MAX; pt rbp = {ix, iy}; int rbits = 0;
        for(int i = 0; i < 9; i++) {
            c = int_cost(&s, ix + o9[i][0], iy + o9[i][1], &bits);
            if(c < rb) { rb = c; rbp = (pt){ix + o9[i][0], iy + o9[i][1]}; rbits = bits; }
        }
        if(it->bi != 1 && rbits > 0) mot_bits[it->lidx] = rbits;
        if(rb < best) { best = rb; mv[0] = (int16_t)((rbp.x - it->x) << 2); mv[1] = (int16_t)((rbp.y - it->y) << 2); }
    }
    it->mv_out[0] = mv[0]; it->mv_out[1] = mv[1]; it->cost = best;
    it->mot_bits_out[0] = mot_bits[0]; it->mot_bits_out[1] = mot_bits[1];
}

/* ---------------------------------------------------------------------------------------------
 * fused residue item (the distortion/transform body of pinter_residue_rdo,
 * src_base/xeve_pinter.c:961-1056)
 * ------------------------------------------------------------------------------------------- */
void xo_residue(const xb200_seq *sq, const xo_planes *pl, const xb200_rates *rates, xb200_residue_item *it,
                int16_t *coef, int16_t *rec)
{
    const int w = it->mc.w, h = it->mc.h, ny = w * h, nc = ny / 4, bd = sq->bit_depth;
    int       l2 = 0; while((1 << l2) < w) l2++;
    int16_t  *pred = malloc(sizeof(int16_t) * (ny + 2 * nc)), *resi = malloc(sizeof(int16_t) * (ny + 2 * nc));
    xo_mc(sq, pl, &it->mc, pred);
    const xo_planes *o = &pl[it->cur_pic];
    const int16_t   *org[3] = {o->y + it->mc.y * o->s_l + it->mc.x, o->u + (it->mc.y / 2) * o->s_c + it->mc.x / 2,
                               o->v + (it->mc.y / 2) * o->s_c + it->mc.x / 2};
    const int so[3] = {o->s_l, o->s_c, o->s_c}, off[3] = {0, ny, ny + nc}, bw[3] = {w, w / 2, w / 2}, bh[3] = {h, h / 2, h / 2};
    for(int c = 0; c < 3; c++) {
        xo_diff(bw[c], bh[c], org[c], so[c], pred + off[c], bw[c], coef + off[c], bw[c]);
        it->dist_pred[c] = xo_ssd(bw[c], bh[c], pred + off[c], bw[c], org[c], so[c], bd);
    }
    xb200_tq_item tq;
    memset(&tq, 0, sizeof(tq));
    tq.log2_cuw = tq.log2_cuh = (uint8_t)l2; tq.slice_type = it->slice_type; tq.is_intra = 0; tq.run_stats = it->run_stats;
    memcpy(tq.qp, it->qp, 3); tq.rate_idx = it->rate_idx; memcpy(tq.lambda, it->lambda, sizeof(tq.lambda));
    xo_tq(sq, &tq, rates, coef, it->nnz);
    memcpy(resi, coef, sizeof(int16_t) * (ny + 2 * nc));
    xo_itdq(sq, &tq, resi, it->nnz);
    for(int c = 0; c < 3; c++) {
        xo_recon(resi + off[c], pred + off[c], it->nnz[c] != 0, bw[c] * bh[c], rec + off[c], bd);
        it->dist_rec[c] = it->nnz[c] ? xo_ssd(bw[c], bh[c], rec + off[c], bw[c], org[c], so[c], bd) : it->dist_pred[c];
    }
    free(pred); free(resi);
}

/* ---------------------------------------------------------------------------------------------
 * picture preparation: input-depth -> internal depth (src_base/xeve_util.c:1552-1571) and edge
 * replication of the PIC_PAD border (src_base/xeve_util.c:190-248)
 * ------------------------------------------------------------------------------------------- */
void xo_pad_plane(int16_t *buf, int stride, int w, int h, int pad)
{
    int16_t *a = buf + pad * stride + pad;
    for(int y = 0; y < h; y++) {
        for(int x = 1; x <= pad; x++) { a[y * stride - x] = a[y * stride]; a[y * stride + w - 1 + x] = a[y * stride + w - 1]; }
    }
    for(int y = 1; y <= pad; y++) {
        memcpy(a - y * stride - pad, a - pad, sizeof(int16_t) * (w + 2 * pad));
        memcpy(a + (h - 1 + y) * stride - pad, a + (h - 1) * stride - pad, sizeof(int16_t) * (w + 2 * pad));
    }
}

/* ---------------------------------------------------------------------------------------------
 * in-loop deblocking.  Follows src_base/xeve_df.c: boundary-strength class of an edge
 * segment (:34-87), the 4-sample luma / chroma filters (:89-251), the per-CU edge walks with their
 * "already coded in this pass" bookkeeping (:253-471) and the two whole-picture passes of
 * xeve_loop_filter (src_base/xeve_enc.c:2355-2414): vertical edges first, then horizontal edges.
 * ------------------------------------------------------------------------------------------- */
#define XO_MCU_IF(m)   (((m) >> 15) & 1)  /* src_base/xeve_def.h:591 */
#define XO_MCU_QP(m)   (((m) >> 16) & 0x7F)
#define XO_MCU_CBFL(m) (((m) >> 24) & 1)
#define XO_MCU_IBC(m)  (((m) >> 26) & 1)

/* boundary class 0..3 of the edge between SCU p (the one whose QP is used) and SCU q, src_base/xeve_df.c:34-87 */
static int df_class(const uint32_t *scu, const int8_t *refi, const int16_t *mv, int64_t p, int64_t q)
{
    if(XO_MCU_IF(scu[p]) || XO_MCU_IF(scu[q])) return 0;
    if(XO_MCU_CBFL(scu[p]) || XO_MCU_CBFL(scu[q])) return 1;
    if(XO_MCU_IBC(scu[p]) || XO_MCU_IBC(scu[q])) return 2;
    int        m[2][2][2]; /* [scu][list][xy], invalid lists count as zero motion */
    const int8_t *rp = refi + 2 * p, *rq = refi + 2 * q;
    for(int l = 0; l < 2; l++)
        for(int k = 0; k < 2; k++) {
            m[0][l][k] = rp[l] >= 0 ? mv[4 * p + 2 * l + k] : 0;
            m[1][l][k] = rq[l] >= 0 ? mv[4 * q + 2 * l + k] : 0;
        }
    int cross; /* compare list l of p against list l (0) or list 1-l (1) of q */
    if(rp[0] == rq[0] && rp[1] == rq[1]) cross = 0;
    else if(rp[0] == rq[1] && rp[1] == rq[0]) cross = 1;
    else return 2;
    for(int l = 0; l < 2; l++)
        for(int k = 0; k < 2; k++)
            if(abs(m[0][l][k] - m[1][l ^ cross][k]) >= 4) return 2;
    return 3;
}

/* one 4-tap line across an edge: p[-2*step] p[-step] | p[0] p[step]; luma moves all four, chroma the inner two
 * (src_base/xeve_df.c:89-251; s16 intermediates as there, C division truncating toward zero) */
static void df_line(int16_t *p, int step, int st, int luma, int maxv)
{
    int16_t A = p[-2 * step], B = p[-step], Cc = p[0], D = p[step];
    int16_t d    = (int16_t)((A - (B << 2) + (Cc << 2) - D) / 8);
    int16_t ad   = (int16_t)(d < 0 ? -d : d);
    int16_t t16  = (int16_t)((ad - st) << 1); if(t16 < 0) t16 = 0;
    int16_t clip = (int16_t)(ad - t16);       if(clip < 0) clip = 0;
    int16_t d1   = (int16_t)(d < 0 ? -clip : clip);
    B = (int16_t)(B + d1); Cc = (int16_t)(Cc - d1);
    if(luma) {
        clip >>= 1;
        int16_t d2 = (int16_t)((A - D) / 4);
        d2 = d2 < -clip ? (int16_t)-clip : (d2 > clip ? clip : d2);
        A = (int16_t)(A - d2); D = (int16_t)(D + d2);
        p[-2 * step] = (int16_t)(A < 0 ? 0 : (A > maxv ? maxv : A));
        p[step]      = (int16_t)(D < 0 ? 0 : (D > maxv ? maxv : D));
    }
    p[-step] = (int16_t)(B < 0 ? 0 : (B > maxv ? maxv : B));
    p[0]     = (int16_t)(Cc < 0 ? 0 : (Cc > maxv ? maxv : Cc));
}

typedef struct {
    int16_t *pl[3];
    int      s[3], w, h, bd;
    const xb200_df_pic *pp;
    const uint32_t *scu; const int8_t *refi; const int16_t *mv;
    uint8_t *done;      /* MCU_GET_COD of the current pass */
} df_ctx;

/* filter the 4-sample (2 for chroma) segment between SCU p (current side) and q; `hor` = horizontal edge */
static void df_segment(const df_ctx *d, int x_scu, int y_scu, int64_t p, int64_t q, int hor)
{
    const int cls = df_class(d->scu, d->refi, d->mv, p, q), qp = XO_MCU_QP(d->scu[p]);
    const int maxv = (1 << d->bd) - 1, bdo = 6 * (d->bd - 8);
    int st = xo_df_strength(cls, qp) << (d->bd - 8);
    if(st) {
        int16_t *b = d->pl[0] + (int64_t)(y_scu * 4) * d->s[0] + x_scu * 4;
        for(int i = 0; i < 4; i++) df_line(hor ? b + i : b + (int64_t)i * d->s[0], hor ? d->s[0] : 1, st, 1, maxv);
    }
    for(int c = 1; c < 3; c++) {
        int qc = qp + (c == 1 ? d->pp->qp_u_offset : d->pp->qp_v_offset);
        qc = qc < -bdo ? -bdo : (qc > 57 ? 57 : qc);
        st = xo_df_strength(cls, d->pp->chroma_qp[c - 1][qc + bdo]) << (d->bd - 8);
        if(!st) continue;
        int16_t *b = d->pl[c] + (int64_t)(y_scu * 2) * d->s[c] + x_scu * 2;
        for(int i = 0; i < 2; i++) df_line(hor ? b + i : b + (int64_t)i * d->s[c], hor ? d->s[c] : 1, st, 0, maxv);
    }
}

void xo_deblock(int16_t *y, int16_t *u, int16_t *v, int s_l, int s_c, int w, int h, const xb200_df_cu *cus, int64_t n,
                const xb200_df_pic *pp, const uint32_t *map_scu, const int8_t *map_refi, const int16_t *map_mv, int bit_depth)
{
    const int ws = pp->w_scu;
    df_ctx d = {{y, u, v}, {s_l, s_c, s_c}, w, h, bit_depth, pp, map_scu, map_refi, map_mv, NULL};
    d.done = malloc((size_t)ws * pp->h_scu);
    for(int hor = 0; hor <= 1; hor++) {
        memset(d.done, 0, (size_t)ws * pp->h_scu);
        for(int64_t i = 0; i < n; i++) {
            const int xs = cus[i].x >> 2, ys = cus[i].y >> 2, cw = (1 << cus[i].log2_cuw) >> 2, ch = (1 << cus[i].log2_cuh) >> 2;
            const int64_t t = xs + (int64_t)ys * ws;
            if(hor) { /* top edge, src_base/xeve_df.c:296-324 */
                if(cus[i].y > 0)
                    for(int k = 0; k < cw; k++) df_segment(&d, xs + k, ys, t + k, t + k - ws, 1);
            }
            else {    /* left edge if the left neighbour was visited in this pass (:386-417), then the right edge if the
                         right neighbour already was (:419-462; cannot happen in z-scan order, kept for fidelity) */
                if(cus[i].x > 0 && d.done[t - 1])
                    for(int k = 0; k < ch; k++) df_segment(&d, xs, ys + k, t + (int64_t)k * ws, t + (int64_t)k * ws - 1, 0);
                if(cus[i].x + (cw << 2) < w && d.done[t + cw])
                    for(int k = 0; k < ch; k++)
                        df_segment(&d, xs + cw, ys + k, t + (int64_t)k * ws + cw, t + (int64_t)k * ws + cw - 1, 0);
            }
            for(int r = 0; r < ch; r++) memset(d.done + t + (int64_t)r * ws, 1, cw);
        }
    }
    free(d.done);
}

/* get_org_bi, src_base/xeve_pinter.c:143-156, applied to the luma prediction of one fn_mc call */
void xo_bi_org(const xb200_seq *sq, const xo_planes *pl, const xb200_mc_item *it, int cur_pic, int16_t *org_bi)
{
    const int w = it->w, h = it->h;
    int16_t  *pred = malloc(sizeof(int16_t) * w * h * 3 / 2);
    xo_mc(sq, pl, it, pred);
    const xo_planes *o = &pl[cur_pic];
    for(int y = 0; y < h; y++)
        for(int x = 0; x < w; x++)
            org_bi[y * w + x] = (int16_t)(((int16_t)o->y[(it->y + y) * o->s_l + it->x + x] << 1) - pred[y * w + x]);
    free(pred);
}
void xo_bi_org_batch(const xb200_seq *sq, const xo_planes *pl, const xb200_mc_item *items, int64_t n, const int32_t *cur_pic,
                     const int64_t *off, int16_t *side)
{
    for(int64_t i = 0; i < n; i++) xo_bi_org(sq, pl, &items[i], cur_pic[i], side + off[i]);
}

/* ---------------------------------------------------------------------------------------------
 * CABAC bit counting for inter RDO (src_base/xeve_eco.c:455-620 engine, :674-1260 syntax;
 * src_base/xeve_mode.c:39-302 counters).  After xeve_sbac_bit_reset the value of xeve_get_bit_number
 * is the number of renormalisation shifts: every carry_propagate moves exactly one byte into
 * {bitcounter, stacked_zero, stacked_ff, pending}, so 8*bytes + 11 - code_bits == shifts.
 * ------------------------------------------------------------------------------------------- */
typedef struct { xb200_sbac s; uint32_t bits; uint16_t ipm[2]; /* ctx.intra_dir: model indices XO_CM_IPM, +1 */ } cabac_t;
#define XO_CM_IPM XB200_CM_COUNT

static void cb_bin(cabac_t *c, int m, int bin) /* xeve_sbac_encode_bin */
{
    uint16_t *pm = m < XB200_CM_COUNT ? &c->s.m[m] : &c->ipm[m - XB200_CM_COUNT];
    uint16_t model = *pm, mps = model & 1, state = model >> 1;
    uint32_t lps = (state * c->s.range) >> 9;
    if(lps < 437) lps = 437;
    c->s.range -= lps;
    if((uint16_t)(bin != 0) != mps) {
        if(c->s.range >= lps) c->s.range = lps;
        state = state + ((512 - state + 16) >> 5);
        if(state > 256) { mps = 1 - mps; state = 512 - state; }
    }
    else state = state - ((state + 16) >> 5);
    *pm = (uint16_t)((state << 1) + mps);
    while(c->s.range < 8192) { c->s.range <<= 1; c->bits++; }
}
static void cb_ep(cabac_t *c) { c->s.range &= ~1u; c->bits++; } /* sbac_encode_bin_ep: range >>= 1, <<= 1 drops the LSB */
static void cb_unary(cabac_t *c, uint32_t sym, int m) /* sbac_write_unary_sym with num_ctx == 2 */
{
    cb_bin(c, m, sym ? 1 : 0);
    if(sym == 0) return;
    while(sym--) cb_bin(c, m + 1, sym ? 1 : 0);
}
static void cb_mvp_idx(cabac_t *c, int idx) /* sbac_write_truncate_unary_sym(idx, 3, 4) */
{
    for(int i = 0; i < 3; i++) {
        int symbol = (i == idx) ? 0 : 1;
        cb_bin(c, XB200_CM_MVP_IDX + i, symbol);
        if(!symbol) break;
    }
}
static void cb_abs_mvd(cabac_t *c, uint32_t val) /* xeve_eco_abs_mvd: exp-golomb, first two bins context coded */
{
    int len_i = 0, nn = (int)((val + 1) >> 1);
    while(len_i < 16 && nn != 0) { nn >>= 1; len_i++; }
    uint32_t info = val + 1 - (1u << len_i), code = (1u << len_i) | (info & ((1u << len_i) - 1));
    int      len_c = (len_i << 1) + 1;
    for(int i = 0; i < len_c; i++) {
        int bin = (code >> (len_c - 1 - i)) & 1;
        if(i <= 1) cb_bin(c, XB200_CM_MVD, bin); else cb_ep(c);
    }
}
static void cb_mvd(cabac_t *c, const int16_t mvd[2])
{
    for(int k = 0; k < 2; k++) {
        int mv = mvd[k] < 0 ? -mvd[k] : mvd[k];
        cb_abs_mvd(c, (uint32_t)mv);
        if(mv) cb_ep(c);
    }
}
static void cb_refi(cabac_t *c, int num_refp, int refi) /* xeve_eco_refi */
{
    if(num_refp <= 1) return;
    if(refi == 0) { cb_bin(c, XB200_CM_REFI, 0); return; }
    cb_bin(c, XB200_CM_REFI, 1);
    for(int i = 2; i < num_refp; i++) {
        int bin = (i == refi + 1) ? 0 : 1;
        if(i == 2) cb_bin(c, XB200_CM_REFI + 1, bin); else cb_ep(c);
        if(!bin) break;
    }
}
static void cb_run_length(cabac_t *c, const int16_t *coef, int log2n, int num_sig, int ch) /* xeve_eco_run_length_cc */
{
    oracle_init();
    const uint16_t *scan = g_scan[log2n];
    const int       n = 1 << (2 * log2n), t0 = ch == 0 ? 0 : 2;
    uint32_t        run = 0;
    for(int sp = 0; sp < n; sp++) {
        int v = coef[scan[sp]];
        if(!v) { run++; continue; }
        uint32_t level = (uint32_t)(v < 0 ? -v : v);
        cb_unary(c, run, XB200_CM_RUN + t0);
        cb_unary(c, level - 1, XB200_CM_LEVEL + t0);
        cb_ep(c); /* sign */
        if(sp == n - 1) break;
        run = 0;
        num_sig--;
        cb_bin(c, XB200_CM_LAST + (ch == 0 ? 0 : 1), num_sig == 0);
        if(num_sig == 0) break;
    }
}
/* xeve_eco_coef for an inter CU <= 64 (one transform block per plane), b_no_cbf = 0 */
static void cb_coef(cabac_t *c, const xb200_bits_item *it, const int16_t *coef, int run_stats)
{
    const int run[3] = {run_stats & 1, (run_stats >> 1) & 1, (run_stats >> 2) & 1};
    const int cbf[3] = {it->nnz[0] != 0, it->nnz[1] != 0, it->nnz[2] != 0};
    int       cbf_all = 0;
    for(int k = 0; k < 3; k++) if(run[k]) cbf_all += cbf[k];
    if(run[0] + run[1] + run[2] == 3) {
        cb_bin(c, XB200_CM_CBF_ALL, cbf_all != 0);
        if(!cbf_all) return;
    }
    if(run[1]) cb_bin(c, XB200_CM_CBF_CB, cbf[1]);
    if(run[2]) cb_bin(c, XB200_CM_CBF_CR, cbf[2]);
    if(run[0] && (cbf[1] + cbf[2] != 0)) cb_bin(c, XB200_CM_CBF_LUMA, cbf[0]);
    const int ny = 1 << (it->log2_cuw + it->log2_cuh), off[3] = {0, ny, ny + (ny >> 2)};
    for(int k = 0; k < 3; k++)
        if(it->nnz[k] && run[k]) cb_run_length(c, coef + off[k], k ? it->log2_cuw - 1 : it->log2_cuw, it->nnz[k], k);
}
void xo_rdo_bits(xb200_bits_item *it, xb200_sbac *states, const int16_t *coef)
{
    cabac_t c;
    c.s = states[it->state_in];
    c.bits = 0;
    const int B = it->slice_type == 0, inter = it->slice_type != 2; /* SLICE_B 0, SLICE_P 1, SLICE_I 2 */
    if(it->kind == 0) {
        if(inter) {
            cb_bin(&c, XB200_CM_SKIP_FLAG + it->ctx_skip, 1);
            cb_mvp_idx(&c, it->mvp_idx[0]);
            if(B) cb_mvp_idx(&c, it->mvp_idx[1]);
        }
    }
    else if(it->kind == 1) {
        if(inter) {
            cb_bin(&c, XB200_CM_SKIP_FLAG + it->ctx_skip, 0);
            if(it->all_preds) cb_bin(&c, XB200_CM_PRED_MODE + it->ctx_pred_mode, 0);
            cb_bin(&c, XB200_CM_DIRECT, it->pidx == 4);
            if(it->pidx != 4) {
                if(it->refi[0] >= 0 && it->refi[1] >= 0) cb_bin(&c, XB200_CM_INTER_DIR, 0);
                else {
                    if(B) cb_bin(&c, XB200_CM_INTER_DIR, 1);
                    cb_bin(&c, XB200_CM_INTER_DIR + 1, it->refi[0] >= 0 ? 0 : 1);
                }
                if(it->refi[0] >= 0) { cb_refi(&c, it->num_refp[0], it->refi[0]); cb_mvp_idx(&c, it->mvp_idx[0]); cb_mvd(&c, it->mvd[0]); }
                if(B && it->refi[1] >= 0) { cb_refi(&c, it->num_refp[1], it->refi[1]); cb_mvp_idx(&c, it->mvp_idx[1]); cb_mvd(&c, it->mvd[1]); }
            }
        }
        cb_coef(&c, it, coef + it->coef_off, 7);
    }
    else if(it->kind == 2) {
        if(it->pidx != 4) {
            if(inter && it->refi[0] >= 0) { cb_mvp_idx(&c, it->mvp_idx[0]); cb_mvd(&c, it->mvd[0]); }
            if(B && it->refi[1] >= 0) { cb_mvp_idx(&c, it->mvp_idx[0]); cb_mvd(&c, it->mvd[1]); }
        }
    }
    else cb_coef(&c, it, coef + it->coef_off, 1 << it->ch);
    it->bits = c.bits;
    if(it->state_out >= 0) states[it->state_out] = c.s;
}
void xo_rdo_bits_batch(xb200_bits_item *items, int64_t n, xb200_sbac *states, const int16_t *coef)
{
    for(int64_t i = 0; i < n; i++) xo_rdo_bits(&items[i], states, coef);
}
/* xeve_rdoq_bit_est + biari_no_bits + xeve_init_bits_est, src_base/xeve_mode.c:304-373 */
void xo_rdoq_rates(const xb200_sbac *st, int64_t n, xb200_rates *out)
{
    static int32_t ebits[1024];
    static int     init;
    if(!init) {
        for(int i = 0; i < 1024; i++) { double p = (512 * (i + 0.5)) / 1024; ebits[i] = (int32_t)(-32768 * (log(p) / log(2.0) - 9)); }
        init = 1;
    }
#define NB(sym, model) ebits[(((uint16_t)((sym) != 0) != ((model) & 1)) ? ((model) >> 1) : (512 - ((model) >> 1))) << 1]
    for(int64_t i = 0; i < n; i++) {
        const uint16_t *m = st[i].m;
        memset(&out[i], 0, sizeof(out[i]));
        for(int b = 0; b < 2; b++) {
            out[i].cbf_all[b] = NB(b, m[XB200_CM_CBF_ALL]); out[i].cbf_luma[b] = NB(b, m[XB200_CM_CBF_LUMA]);
            out[i].cbf_cb[b] = NB(b, m[XB200_CM_CBF_CB]); out[i].cbf_cr[b] = NB(b, m[XB200_CM_CBF_CR]);
            for(int k = 0; k < 24; k++) { out[i].run[k][b] = NB(b, m[XB200_CM_RUN + k]); out[i].level[k][b] = NB(b, m[XB200_CM_LEVEL + k]); }
            for(int k = 0; k < 2; k++) out[i].last[k][b] = NB(b, m[XB200_CM_LAST + k]);
        }
    }
#undef NB
}

/* ---------------------------------------------------------------------------------------------
 * xeve_pinter_analyze_cu (src_base/xeve_pinter.c:1839-2056) and its parts: xeve_analyze_skip (:1337-1530),
 * analyze_t_direct (:1532-1565), check_best_mvp (:1772-1837), analyze_bi (:1567-1683), pinter_residue_rdo with the
 * cbf decisions (:906-1335).  rdo_dbk_switch = 0, cu_qp_delta off (presets fast / medium).
 * Costs are doubles evaluated in the reference's association order; this file is compiled without FMA contraction.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const xb200_seq *sq; const xo_planes *pl; const xb200_rates *rates; const xb200_cu_item *cu;
    int w, ny, nc, n;                 /* cu width, luma / chroma plane sizes, 3/2 * ny */
    xb200_sbac st0;                   /* core->s_curr_best[..] */
} cu_env;
typedef struct {
    int8_t refi[2]; int16_t mv[2][2], mvd[2][2]; uint8_t mvp_idx[2]; int nnz[3];
    int16_t *coef, *pred; xb200_sbac st; double cost;
} cu_mode;

static uint32_t cu_bits(const cu_env *e, int kind, int pidx, const int8_t refi[2], int16_t mvd[2][2], const uint8_t mvp_idx[2],
                        const int nnz[3], int ch, const int16_t *coef, const xb200_sbac *in, xb200_sbac *out)
{
    xb200_bits_item it;
    xb200_sbac      st[2];
    memset(&it, 0, sizeof(it));
    it.kind = (uint8_t)kind; it.slice_type = e->cu->slice_type; it.log2_cuw = e->cu->log2_cuw; it.log2_cuh = e->cu->log2_cuh;
    it.pidx = (uint8_t)pidx; it.ch = (uint8_t)ch; it.ctx_skip = e->cu->ctx_skip; it.ctx_pred_mode = e->cu->ctx_pred_mode;
    it.all_preds = e->cu->all_preds;
    it.num_refp[0] = e->cu->num_refp[0]; it.num_refp[1] = e->cu->num_refp[1];
    if(refi) { it.refi[0] = refi[0]; it.refi[1] = refi[1]; }
    if(mvp_idx) { it.mvp_idx[0] = mvp_idx[0]; it.mvp_idx[1] = mvp_idx[1]; }
    if(mvd) memcpy(it.mvd, mvd, sizeof(it.mvd));
    if(nnz) { it.nnz[0] = nnz[0]; it.nnz[1] = nnz[1]; it.nnz[2] = nnz[2]; }
    it.state_in = 0; it.state_out = 1; it.coef_off = 0;
    st[0] = *in;
    xo_rdo_bits(&it, st, coef);
    if(out) *out = st[1];
    return it.bits;
}
static void cu_mc_item(const cu_env *e, const int8_t refi[2], int16_t mv[2][2], xb200_mc_item *m)
{
    memset(m, 0, sizeof(*m));
    m->poc = e->cu->poc; m->x = e->cu->x; m->y = e->cu->y; m->w = m->h = (int16_t)e->w;
    for(int l = 0; l < 2; l++) {
        m->refi[l] = refi[l]; m->mv[l][0] = mv[l][0]; m->mv[l][1] = mv[l][1];
        m->ref_pic[l] = refi[l] >= 0 ? e->cu->ref_pic[l][refi[l]] : -1;
        m->ref_poc[l] = refi[l] >= 0 ? e->cu->ref_poc[l][refi[l]] : -1;
    }
}
/* pinter_residue_rdo: m->refi/mv/mvd set; fills m->coef, m->pred, m->nnz, m->st (s_temp_best), returns the cost */
static double cu_residue_rdo(const cu_env *e, int pidx, cu_mode *m, const uint8_t mvp_idx[2])
{
    const xb200_cu_item *cu = e->cu;
    const double         w0 = cu->dist_chroma_weight[0], w1 = cu->dist_chroma_weight[1];
    xb200_residue_item   it;
    int16_t             *rec = malloc(sizeof(int16_t) * e->n);
    memset(&it, 0, sizeof(it));
    cu_mc_item(e, m->refi, m->mv, &it.mc);
    it.cur_pic = cu->cur_pic; it.slice_type = cu->slice_type; it.run_stats = 7; memcpy(it.qp, cu->qp, 3);
    it.rate_idx = cu->rate_idx; memcpy(it.lambda, cu->lambda, sizeof(it.lambda));
    xo_residue(e->sq, e->pl, e->rates, &it, m->coef, rec);
    xo_mc(e->sq, e->pl, &it.mc, m->pred);
    free(rec);
    const int64_t *d0 = it.dist_pred, *d1 = it.dist_rec;
    const int      store[3] = {it.nnz[0], it.nnz[1], it.nnz[2]}, zero[3] = {0, 0, 0};
    double         cost, best = 1.7e+308;
    int            cbf[3] = {0, 0, 0};
    xb200_sbac     run;
    if(store[0] + store[1] + store[2]) {
#define DSEL(c, on) ((on) ? d1[c] : d0[c])
#define TRY(n0, n1, n2)                                                                                                  \
    do {                                                                                                                 \
        const int nn_[3] = {(n0) ? store[0] : 0, (n1) ? store[1] : 0, (n2) ? store[2] : 0};                              \
        cost = (double)DSEL(0, n0) + (((double)DSEL(1, n1) * w0) + ((double)DSEL(2, n2) * w1));                          \
        cost += (double)cu_bits(e, 1, pidx, m->refi, m->mvd, mvp_idx, nn_, 0, m->coef, &e->st0, &run) * cu->lambda[0];   \
        if(cost < best) { best = cost; cbf[0] = (n0); cbf[1] = (n1); cbf[2] = (n2); m->st = run; }                       \
    } while(0)
        if(pidx != 4) TRY(0, 0, 0);                                       /* forced all-zero */
        TRY(store[0] > 0, store[1] > 0, store[2] > 0);                    /* as it is */
        int        idx_best[3] = {0, 0, 0}, cur[3] = {store[0], store[1], store[2]};
        xb200_sbac prev_best = e->st0;
        for(int i = 0; i < 3; i++) {                                      /* per-component cbf test */
            if(store[i] <= 0) continue;
            double           comp_best = 1.7e+308;
            const xb200_sbac prev_run = prev_best;
            for(int j = 0; j < 2; j++) {
                cost = i == 0 ? (double)(DSEL(i, j) * 1) : (double)DSEL(i, j) * cu->dist_chroma_weight[i - 1];
                cur[i] = j ? store[i] : 0;
                cost += (double)cu_bits(e, 3, pidx, m->refi, m->mvd, mvp_idx, cur, i, m->coef, &prev_run, &run) * cu->lambda[i];
                if(cost < comp_best) { comp_best = cost; idx_best[i] = j; prev_best = run; }
            }
        }
        if(idx_best[0] || idx_best[1] || idx_best[2]) {
            const int differs = (idx_best[0] ? store[0] : 0) != store[0] || (idx_best[1] ? store[1] : 0) != store[1] ||
                                (idx_best[2] ? store[2] : 0) != store[2];
            if(differs) TRY(idx_best[0], idx_best[1], idx_best[2]);
        }
#undef TRY
#undef DSEL
        const int off[3] = {0, e->ny, e->ny + e->nc}, sz[3] = {e->ny, e->nc, e->nc};
        for(int i = 0; i < 3; i++) {
            m->nnz[i] = cbf[i] ? store[i] : 0;
            if(m->nnz[i] == 0 && store[i] != 0) memset(m->coef + off[i], 0, sizeof(int16_t) * sz[i]);
        }
    }
    else {
        best = (double)d0[0] + (w0 * (double)d0[1]) + (w1 * (double)d0[2]);
        best += (double)cu_bits(e, 1, pidx, m->refi, m->mvd, mvp_idx, zero, 0, m->coef, &e->st0, &m->st) * cu->lambda[0];
        m->nnz[0] = m->nnz[1] = m->nnz[2] = 0;
    }
    return best;
}

static int16_t *g_pred_y_sink; /* xo_chain_picture: where the winner's luma prediction (mi->pred_y_best) goes */
void xo_analyze_cu(const xb200_seq *sq, const xo_planes *pl, const xb200_rates *rates, xb200_cu_item *cu, xb200_sbac *states,
                   int16_t *coef_out, int16_t *rec_out)
{
    cu_env e;
    e.sq = sq; e.pl = pl; e.rates = rates; e.cu = cu;
    e.w = 1 << cu->log2_cuw; e.ny = e.w * e.w; e.nc = e.ny >> 2; e.n = e.ny + 2 * e.nc;
    e.st0 = states[cu->state_in];
    const int    B = cu->slice_type == 0, bd = sq->bit_depth;
    const double w0 = cu->dist_chroma_weight[0], w1 = cu->dist_chroma_weight[1];
    cu_mode      md[5];
    int16_t     *buf = calloc((size_t)e.n * 11, sizeof(int16_t)), *tmp_pred = buf + 10 * e.n;
    double       cost_inter[5], cost_best = 1.7e+308;
    int          best_idx = 3;
    for(int i = 0; i < 5; i++) {
        memset(&md[i], 0, sizeof(md[i]));
        md[i].coef = buf + (2 * i) * e.n; md[i].pred = buf + (2 * i + 1) * e.n;
        cost_inter[i] = 1.7e+308;
    }
    const xo_planes *o = &pl[cu->cur_pic];
    const int16_t   *org[3] = {o->y + cu->y * o->s_l + cu->x, o->u + (cu->y / 2) * o->s_c + cu->x / 2,
                               o->v + (cu->y / 2) * o->s_c + cu->x / 2};

    /* ---- xeve_analyze_skip ---- */
    int64_t best_ssd = (int64_t)1 << (2 * cu->log2_cuw + 16);
    {
        cu_mode *m = &md[3];
        double   sb = 1.7e+308;
        for(int idx0 = 0; idx0 < sq->merge_num; idx0++) {
            int dup = 0;
            for(int t = idx0 - 1; t >= 0; t--) dup |= cu->mvp[0][t][0] == cu->mvp[0][idx0][0] && cu->mvp[0][t][1] == cu->mvp[0][idx0][1];
            if(dup) continue;
            const int cnt = B ? sq->merge_num : 1;
            for(int idx1 = 0; idx1 < cnt; idx1++) {
                dup = 0;
                for(int t = idx1 - 1; t >= 0; t--) dup |= cu->mvp[1][t][0] == cu->mvp[1][idx1][0] && cu->mvp[1][t][1] == cu->mvp[1][idx1][1];
                if(dup) continue;
                int8_t  refi[2] = {cu->refi_pred[0][idx0], B ? cu->refi_pred[1][idx1] : -1};
                int16_t mv[2][2] = {{cu->mvp[0][idx0][0], cu->mvp[0][idx0][1]}, {cu->mvp[1][idx1][0], cu->mvp[1][idx1][1]}};
                if(refi[0] < 0 && refi[1] < 0) continue;
                xb200_mc_item mc;
                cu_mc_item(&e, refi, mv, &mc);
                xo_mc(sq, pl, &mc, tmp_pred);
                const int64_t cy = xo_ssd(e.w, e.w, tmp_pred, e.w, org[0], o->s_l, bd);
                const int64_t cb = xo_ssd(e.w / 2, e.w / 2, tmp_pred + e.ny, e.w / 2, org[1], o->s_c, bd);
                const int64_t cr = xo_ssd(e.w / 2, e.w / 2, tmp_pred + e.ny + e.nc, e.w / 2, org[2], o->s_c, bd);
                double        cost = (double)cy + (w0 * (double)cb) + (w1 * (double)cr);
                const uint8_t mi[2] = {(uint8_t)idx0, (uint8_t)idx1};
                xb200_sbac    run;
                cost += (double)cu_bits(&e, 0, 3, NULL, NULL, mi, NULL, 0, NULL, &e.st0, &run) * cu->lambda[0];
                if(cost < sb) {
                    sb = cost;
                    m->mvp_idx[0] = mi[0]; m->mvp_idx[1] = mi[1];
                    memcpy(m->mv, mv, sizeof(mv)); memset(m->mvd, 0, sizeof(m->mvd));
                    m->refi[0] = refi[0]; m->refi[1] = refi[1];
                    best_ssd = cy + cb + cr;
                    memcpy(m->pred, tmp_pred, sizeof(int16_t) * e.n);
                    m->st = run;
                }
            }
        }
        cost_inter[3] = sb;
        if(sb < cost_best) { best_idx = 3; cost_best = sb; }
    }
    if(best_idx == 3 && cost_best < 1.7e+308 && (double)best_ssd > 0.0) {
        if(B) {   /* ---- analyze_t_direct ---- */
            cu_mode *m = &md[4];
            memcpy(m->mv, cu->mv_dir, sizeof(m->mv)); memset(m->mvd, 0, sizeof(m->mvd));
            m->refi[0] = m->refi[1] = 0;
            const double c = cost_inter[4] = cu_residue_rdo(&e, 4, m, m->mvp_idx);
            if(c < cost_best) { best_idx = 4; cost_best = c; }
        }
        /* ---- uni-directional motion search ---- */
        int32_t mot_bits[2] = {0, 0};
        int16_t mv_scale[2][XB200_MAX_REFP][2];
        uint8_t mvp_idx[2] = {0, 0};
        int     num_refp_cur = 0;
        memset(mv_scale, 0, sizeof(mv_scale));
        for(int lidx = 0; lidx <= (B ? 1 : 0); lidx++) {
            cu_mode *m = &md[lidx];
            uint32_t best_me = 0xFFFFFFFFu;
            int      refi_t = 0;
            num_refp_cur = cu->num_refp[lidx];
            mvp_idx[lidx] = md[3].mvp_idx[lidx];
            for(int r = 0; r < num_refp_cur; r++) {
                xb200_me_item me;
                memset(&me, 0, sizeof(me));
                me.poc = cu->poc; me.cur_pic = cu->cur_pic; me.ref_pic = cu->ref_pic[lidx][r]; me.ref_poc = cu->ref_poc[lidx][r];
                me.x = cu->x; me.y = cu->y; me.log2_cuw = cu->log2_cuw; me.log2_cuh = cu->log2_cuh; me.lidx = (uint8_t)lidx; me.bi = 0;
                me.refi = (int8_t)r; me.num_refp = (uint8_t)num_refp_cur;
                me.mvp[0] = cu->mvp[lidx][mvp_idx[lidx]][0]; me.mvp[1] = cu->mvp[lidx][mvp_idx[lidx]][1];
                me.lambda_mv = cu->lambda_mv; me.mot_bits_in[0] = mot_bits[0]; me.mot_bits_in[1] = mot_bits[1];
                me.max_search_range = cu->max_search_range; me.gop_size = sq->gop_size; me.org_bi_off = -1;
                xo_me(sq, pl, NULL, &me);
                mot_bits[0] = me.mot_bits_out[0]; mot_bits[1] = me.mot_bits_out[1];
                mv_scale[lidx][r][0] = me.mv_out[0]; mv_scale[lidx][r][1] = me.mv_out[1];
                if(me.cost < best_me) { best_me = me.cost; refi_t = r; }
            }
            m->mv[lidx][0] = mv_scale[lidx][refi_t][0]; m->mv[lidx][1] = mv_scale[lidx][refi_t][1];
            m->refi[lidx] = (int8_t)refi_t; m->refi[1 - lidx] = -1;
            /* check_best_mvp: the loop never updates its reference cost (quirk q1) */
            {
                int16_t  mvd[2][2] = {{0, 0}, {0, 0}};
                uint8_t  mi[2] = {mvp_idx[lidx], 0};
                const int16_t (*cand)[2] = cu->mvp[lidx];
                mvd[lidx][0] = (int16_t)(m->mv[lidx][0] - cand[mi[0]][0]); mvd[lidx][1] = (int16_t)(m->mv[lidx][1] - cand[mi[0]][1]);
                const double ref_cost = (double)cu_bits(&e, 2, lidx, m->refi, mvd, mi, NULL, 0, NULL, &e.st0, NULL) * cu->lambda[0];
                int          best = mi[0];
                for(int idx = 0; idx < 4; idx++) {
                    int dup = 0;
                    for(int t = idx - 1; t >= 0; t--) dup |= cand[idx][0] == cand[t][0] && cand[idx][1] == cand[t][1];
                    if(dup) continue;
                    mvd[lidx][0] = (int16_t)(m->mv[lidx][0] - cand[idx][0]); mvd[lidx][1] = (int16_t)(m->mv[lidx][1] - cand[idx][1]);
                    mi[0] = (uint8_t)idx;
                    const double c = (double)cu_bits(&e, 2, lidx, m->refi, mvd, mi, NULL, 0, NULL, &e.st0, NULL) * cu->lambda[0];
                    if(c < ref_cost) best = idx;
                }
                mvp_idx[lidx] = (uint8_t)best;
                m->mvd[lidx][0] = (int16_t)(m->mv[lidx][0] - cand[best][0]); m->mvd[lidx][1] = (int16_t)(m->mv[lidx][1] - cand[best][1]);
            }
            m->mvp_idx[lidx] = mvp_idx[lidx];
            const double c = cost_inter[lidx] = cu_residue_rdo(&e, lidx, m, mvp_idx);
            if(c < cost_best) { best_idx = lidx; cost_best = c; }
        }
        if(B) {   /* ---- analyze_bi ---- */
            cu_mode *m = &md[2];
            int      lidx_ref = cost_inter[0] <= cost_inter[1] ? 0 : 1, lidx_cnd = 1 - lidx_ref;
            int8_t   refi[2] = {-1, -1};
            uint32_t best_me = 0xFFFFFFFFu;
            int      refi_best = 0;
            int16_t *org_bi = malloc(sizeof(int16_t) * e.ny);
            m->mvp_idx[0] = md[0].mvp_idx[0]; m->mvp_idx[1] = md[1].mvp_idx[1];
            m->refi[0] = md[0].refi[0]; m->refi[1] = md[1].refi[1];
            for(int l = 0; l < 2; l++) { m->mv[l][0] = md[l].mv[l][0]; m->mv[l][1] = md[l].mv[l][1]; }
            refi[lidx_ref] = m->refi[lidx_ref];
            for(int it = 0; it < 4; it++) {   /* BI_ITER */
                xb200_mc_item mc;
                cu_mc_item(&e, refi, m->mv, &mc);
                xo_bi_org(sq, pl, &mc, cu->cur_pic, org_bi);
                { int8_t t = refi[lidx_ref]; refi[lidx_ref] = refi[lidx_cnd]; refi[lidx_cnd] = t; }
                { int t = lidx_ref; lidx_ref = lidx_cnd; lidx_cnd = t; }
                const int mi = m->mvp_idx[lidx_ref];
                int       changed = 0;
                for(int r = 0; r < num_refp_cur; r++) {
                    xb200_me_item me;
                    memset(&me, 0, sizeof(me));
                    refi[lidx_ref] = (int8_t)r;
                    me.poc = cu->poc; me.cur_pic = cu->cur_pic; me.ref_pic = cu->ref_pic[lidx_ref][r]; me.ref_poc = cu->ref_poc[lidx_ref][r];
                    me.x = cu->x; me.y = cu->y; me.log2_cuw = cu->log2_cuw; me.log2_cuh = cu->log2_cuh; me.lidx = (uint8_t)lidx_ref; me.bi = 1;
                    me.refi = (int8_t)r; me.num_refp = (uint8_t)num_refp_cur;
                    me.mvp[0] = cu->mvp[lidx_ref][mi][0]; me.mvp[1] = cu->mvp[lidx_ref][mi][1];
                    me.mv_in[0] = mv_scale[lidx_ref][r][0]; me.mv_in[1] = mv_scale[lidx_ref][r][1];
                    me.lambda_mv = cu->lambda_mv; me.mot_bits_in[0] = mot_bits[0]; me.mot_bits_in[1] = mot_bits[1];
                    me.max_search_range = cu->max_search_range; me.gop_size = sq->gop_size; me.org_bi_off = 0;
                    xo_me(sq, pl, org_bi, &me);
                    mot_bits[0] = me.mot_bits_out[0]; mot_bits[1] = me.mot_bits_out[1];
                    mv_scale[lidx_ref][r][0] = me.mv_out[0]; mv_scale[lidx_ref][r][1] = me.mv_out[1];
                    if(me.cost < best_me) {
                        refi_best = r; best_me = me.cost; changed = 1;
                        m->refi[lidx_ref] = (int8_t)refi_best;   /* the other list keeps pi->refi[pidx][lidx_cnd] */
                        m->mv[lidx_ref][0] = me.mv_out[0]; m->mv[lidx_ref][1] = me.mv_out[1];
                    }
                }
                refi[lidx_ref] = (int8_t)refi_best; refi[lidx_cnd] = -1;
                if(!changed) break;
            }
            free(org_bi);
            for(int l = 0; l < 2; l++) {
                m->mvd[l][0] = (int16_t)(m->mv[l][0] - cu->mvp[l][m->mvp_idx[l]][0]);
                m->mvd[l][1] = (int16_t)(m->mv[l][1] - cu->mvp[l][m->mvp_idx[l]][1]);
            }
            const double c = cost_inter[2] = cu_residue_rdo(&e, 2, m, m->mvp_idx);
            if(c < cost_best) { best_idx = 2; cost_best = c; }
        }
    }
    /* ---- reconstruct and report the winner ---- */
    cu_mode *m = &md[best_idx];
    if(best_idx == 3) { memset(m->coef, 0, sizeof(int16_t) * e.n); m->nnz[0] = m->nnz[1] = m->nnz[2] = 0; }
    memcpy(coef_out, m->coef, sizeof(int16_t) * e.n);
    {
        xb200_tq_item tq;
        int16_t      *resi = malloc(sizeof(int16_t) * e.n);
        memset(&tq, 0, sizeof(tq));
        tq.log2_cuw = tq.log2_cuh = cu->log2_cuw; tq.slice_type = cu->slice_type; tq.run_stats = 7; memcpy(tq.qp, cu->qp, 3);
        memcpy(resi, m->coef, sizeof(int16_t) * e.n);
        xo_itdq(sq, &tq, resi, m->nnz);
        const int off[3] = {0, e.ny, e.ny + e.nc}, sz[3] = {e.ny, e.nc, e.nc};
        for(int c = 0; c < 3; c++) xo_recon(resi + off[c], m->pred + off[c], m->nnz[c] != 0, sz[c], rec_out + off[c], bd);
        free(resi);
    }
    cu->cost = cost_inter[best_idx]; cu->best_idx = (uint8_t)best_idx;
    for(int l = 0; l < 2; l++) {
        cu->refi[l] = m->refi[l]; cu->mvp_idx[l] = m->mvp_idx[l];
        cu->mv[l][0] = m->mv[l][0]; cu->mv[l][1] = m->mv[l][1]; cu->mvd[l][0] = m->mvd[l][0]; cu->mvd[l][1] = m->mvd[l][1];
    }
    cu->nnz[0] = m->nnz[0]; cu->nnz[1] = m->nnz[1]; cu->nnz[2] = m->nnz[2];
    if(cu->state_out >= 0) states[cu->state_out] = m->st;
    if(g_pred_y_sink) memcpy(g_pred_y_sink, m->pred, sizeof(int16_t) * e.ny); /* src_base/xeve_pinter.c:2040 */
    free(buf);
}
void xo_analyze_cu_batch(const xb200_seq *sq, const xo_planes *pl, const xb200_rates *rates, xb200_cu_item *items, int64_t n,
                         xb200_sbac *states, int16_t *coef, int16_t *rec)
{
    for(int64_t i = 0; i < n; i++) xo_analyze_cu(sq, pl, rates, &items[i], states, coef + items[i].out_off, rec + items[i].out_off);
}

/* ---------------------------------------------------------------------------------------------
 * intra analysis: pintra_analyze_cu, make_ipred_list, pintra_residue_rdo
 * (src_base/xeve_pintra.c:69-374, 544-698), the Baseline predictors (src_base/xeve_ipred.c:99-228) and the intra bit
 * counters (src_base/xeve_mode.c:81-171, src_base/xeve_eco.c:793-905, 1104-1121).
 * ------------------------------------------------------------------------------------------- */
/* src_base/xeve_ipred.c:99-228; le / up point at sample 0 of the left column / upper row (index -1 = corner) */
static void ipred_b(const int16_t *le, const int16_t *up, int16_t *dst, int ipm, int w, int h)
{
    int l2 = 0, dc = 0;
    while((1 << l2) < w) l2++;
    if(ipm == 0) {
        for(int i = 0; i < h; i++) dc += le[i];
        for(int j = 0; j < w; j++) dc += up[j];
        dc = (dc + w) >> (l2 + 1);
    }
    for(int i = 0; i < h; i++)
        for(int j = 0; j < w; j++) {
            int v;
            switch(ipm) {
            case 0: v = dc; break;
            case 1: v = le[i]; break;
            case 2: v = up[j]; break;
            case 3: v = i > j ? le[i - j - 1] : (i == j ? up[-1] : up[j - i - 1]); break;
            default: v = (up[i + j + 1] + le[i + j + 1]) >> 1; break;
            }
            dst[i * w + j] = (int16_t)v;
        }
}

/* xeve_eco_coef for an intra CU (pred_mode == MODE_INTRA branch of xeve_eco_cbf): cbf_cb, cbf_cr, cbf_luma for the planes
 * named by run_stats, then the run-length coded planes */
static void cb_coef_intra(cabac_t *c, const int nnz[3], int log2_cuw, const int16_t *coef, int run_stats)
{
    const int ny = 1 << (2 * log2_cuw), off[3] = {0, ny, ny + (ny >> 2)};
    if(run_stats & 2) cb_bin(c, XB200_CM_CBF_CB, nnz[1] != 0);
    if(run_stats & 4) cb_bin(c, XB200_CM_CBF_CR, nnz[2] != 0);
    if(run_stats & 1) cb_bin(c, XB200_CM_CBF_LUMA, nnz[0] != 0);
    for(int k = 0; k < 3; k++)
        if(nnz[k] && ((run_stats >> k) & 1)) cb_run_length(c, coef + off[k], k ? log2_cuw - 1 : log2_cuw, nnz[k], k);
}

void xo_analyze_intra(const xb200_seq *sq, const xo_planes *pl, const xb200_rates *rates, xb200_intra_item *it, xb200_sbac *states,
                      const int16_t *side, int16_t *coef_out, int16_t *rec_out)
{
    const int l2 = it->log2_cuw, w = 1 << l2, ny = w * w, nc = ny >> 2, wc = w >> 1, bd = sq->bit_depth;
    const xo_planes *o = &pl[it->cur_pic];
    const int16_t *org[3] = {o->y + it->y * o->s_l + it->x, o->u + (it->y >> 1) * o->s_c + (it->x >> 1),
                             o->v + (it->y >> 1) * o->s_c + (it->x >> 1)};
    /* neighbour samples: per plane left[-1..2n-1] | up[-1..2n-1] */
    const int16_t *nb = side + it->nb_off, *le[3], *up[3];
    for(int c = 0; c < 3; c++) {
        const int n = c ? wc : w;
        le[c] = nb + 1; up[c] = nb + (2 * n + 1) + 1;
        nb += 2 * (2 * n + 1);
    }
    cabac_t base, run;
    base.s = states[it->state_in]; base.ipm[0] = it->cm_ipm_in[0]; base.ipm[1] = it->cm_ipm_in[1]; base.bits = 0;

    /* ---- make_ipred_list: every mode predicted once (pred_cache), ranked by SATD + sqrt(lambda) * mode bits ---- */
    int16_t *pred_cache = malloc(sizeof(int16_t) * 5 * ny);
    int      list[5];
    double   cand_cost[5];
    uint32_t cand_satd[5];
    for(int i = 0; i < 5; i++) { list[i] = 0; cand_cost[i] = 1.7e+308; cand_satd[i] = UINT32_MAX; }
    for(int i = 0; i < 5; i++) {
        ipred_b(le[0], up[0], pred_cache + i * ny, i, w, w);
        const uint32_t satd = (uint32_t)xo_satd(w, w, org[0], o->s_l, pred_cache + i * ny, w, bd);
        run = base;
        cb_unary(&run, it->mpm[i], XO_CM_IPM);
        double cost = (double)satd;
        cost += (double)run.bits * it->sqrt_lambda0;
        int shift = 0;
        while(shift < 5 && cost < cand_cost[4 - shift]) shift++;
        if(shift) {
            for(int j = 1; j < shift; j++) {
                list[5 - j] = list[4 - j]; cand_cost[5 - j] = cand_cost[4 - j]; cand_satd[5 - j] = cand_satd[4 - j];
            }
            list[5 - shift] = i; cand_cost[5 - shift] = cost; cand_satd[5 - shift] = satd;
        }
    }
    int pred_cnt = 5;
    for(int i = 4; i >= 1; i--) {
        if(cand_satd[i] > it->inter_satd * (1.2)) pred_cnt--;
        else break;
    }

    /* ---- luma RDO over the surviving modes (pintra_residue_rdo, mode 0) ------------------------------------------- */
    xb200_tq_item tq;
    memset(&tq, 0, sizeof(tq));
    tq.log2_cuw = tq.log2_cuh = (uint8_t)l2; tq.slice_type = it->slice_type; tq.is_intra = 1;
    tq.qp[0] = it->qp[0]; tq.qp[1] = it->qp[1]; tq.qp[2] = it->qp[2]; tq.rate_idx = it->rate_idx;
    tq.lambda[0] = it->lambda[0]; tq.lambda[1] = it->lambda[1]; tq.lambda[2] = it->lambda[2];
    int16_t *tmp = calloc((size_t)ny + 2 * nc, sizeof(int16_t)), *rec = calloc((size_t)ny + 2 * nc, sizeof(int16_t));
    int16_t *pred_c = malloc(sizeof(int16_t) * 2 * nc);
    double   cost = 1.7e+308;
    int      best_ipd = -1, nnz[3] = {0, 0, 0}, nnz_best[3] = {0, 0, 0};
    int32_t  best_dist_y = 0, best_dist_c = 0;
    run = base;
    for(int j = 0; j < pred_cnt; j++) {
        const int      ipm = list[j];
        const int16_t *pred = pred_cache + ipm * ny;
        xo_diff(w, w, org[0], o->s_l, pred, w, tmp, w);
        tq.run_stats = 1;
        xo_tq(sq, &tq, rates, tmp, nnz);
        int16_t *cy = malloc(sizeof(int16_t) * ny);
        memcpy(cy, tmp, sizeof(int16_t) * ny);
        run = base; run.bits = 0;
        if(it->slice_type != 2 && it->all_preds) {
            cb_bin(&run, XB200_CM_SKIP_FLAG + it->ctx_skip, 0);
            cb_bin(&run, XB200_CM_PRED_MODE + it->ctx_pred_mode, 1);
        }
        cb_unary(&run, it->mpm[ipm], XO_CM_IPM);
        cb_coef_intra(&run, nnz, l2, tmp, 1);
        xo_itdq(sq, &tq, tmp, nnz);
        xo_recon(tmp, pred, nnz[0], ny, rec, bd);
        double cost_t = 0;
        cost_t += (double)xo_ssd(w, w, rec, w, org[0], o->s_l, bd);
        const int32_t dist_t = (int32_t)cost_t;
        cost_t += (double)run.bits * it->lambda[0];
        if(cost_t < cost) {
            cost = cost_t; best_dist_y = dist_t; best_ipd = ipm;
            memcpy(coef_out, cy, sizeof(int16_t) * ny);
            memcpy(rec_out, rec, sizeof(int16_t) * ny);
            nnz_best[0] = nnz[0];
        }
        free(cy);
    }
    if(pred_cnt == 0) { /* cannot happen (the loop above keeps at least one mode) but mirrors :600-602 */
        it->cost = 1.7e+308;
        goto done;
    }

    /* ---- chroma with the winning luma mode (pintra_residue_rdo, mode 1); its bit count does not reach the result ---- */
    for(int c = 1; c < 3; c++) {
        ipred_b(le[c], up[c], pred_c + (c - 1) * nc, best_ipd, wc, wc);
        xo_diff(wc, wc, org[c], o->s_c, pred_c + (c - 1) * nc, wc, tmp + ny + (c - 1) * nc, wc);
    }
    tq.run_stats = 6;
    xo_tq(sq, &tq, rates, tmp, nnz);
    memcpy(coef_out + ny, tmp + ny, sizeof(int16_t) * 2 * nc);
    nnz[0] = 0;
    xo_itdq(sq, &tq, tmp, nnz);
    {
        double cc = 0;
        for(int c = 1; c < 3; c++) {
            xo_recon(tmp + ny + (c - 1) * nc, pred_c + (c - 1) * nc, nnz[c], nc, rec_out + ny + (c - 1) * nc, bd);
            cc += it->dist_chroma_weight[c - 1] * (double)xo_ssd(wc, wc, rec_out + ny + (c - 1) * nc, wc, org[c], o->s_c, bd);
        }
        best_dist_c = (int32_t)cc;
    }
    nnz_best[1] = nnz[1]; nnz_best[2] = nnz[2];

    /* ---- final bit count of the whole CU from the input state (xeve_rdo_bit_cnt_cu_intra) -------------------------- */
    run = base; run.bits = 0;
    if(it->slice_type != 2) {
        cb_bin(&run, XB200_CM_SKIP_FLAG + it->ctx_skip, 0);
        cb_bin(&run, XB200_CM_PRED_MODE + it->ctx_pred_mode, 1);
    }
    cb_unary(&run, it->mpm[best_ipd], XO_CM_IPM);
    cb_coef_intra(&run, nnz_best, l2, coef_out, 7);
    cost = (double)run.bits * it->lambda[0];
    cost += best_dist_y;
    cost += best_dist_c;
    it->cost = cost;
    it->dist_cu = best_dist_y + best_dist_c;
    it->ipm[0] = it->ipm[1] = (int8_t)best_ipd;
    it->nnz[0] = nnz_best[0]; it->nnz[1] = nnz_best[1]; it->nnz[2] = nnz_best[2];
    it->cm_ipm_out[0] = run.ipm[0]; it->cm_ipm_out[1] = run.ipm[1];
    states[it->state_out] = run.s;
done:
    free(pred_cache); free(tmp); free(rec); free(pred_c);
}
void xo_analyze_intra_batch(const xb200_seq *sq, const xo_planes *pl, const xb200_rates *rates, xb200_intra_item *items, int64_t n,
                            xb200_sbac *states, const int16_t *side, int16_t *coef, int16_t *rec)
{
    for(int64_t i = 0; i < n; i++) xo_analyze_intra(sq, pl, rates, &items[i], states, side, coef + items[i].out_off, rec + items[i].out_off);
}

/* ---------------------------------------------------------------------------------------------
 * MV-predictor inputs of one CU: xeve_get_avail_inter (src_base/xeve_util.c:652-715, one tile),
 * xeve_get_motion (:526-573: left / up / up-right spatial candidates or (1,1), colocated MV as the fourth, refi always 0) and
 * the temporal-direct MVs of xeve_get_mv_dir (:619-650: POC-scaled colocated MV of list 1, C integer division).
 * ------------------------------------------------------------------------------------------- */
void xo_mvp(xb200_mvp_item *it, const xb200_mvp_pic *pp, const uint32_t *map_scu, const int16_t *map_mv, const int16_t *col_mv0,
            const int16_t *col_mv1)
{
    const int x = it->x_scu, y = it->y_scu, w = pp->w_scu, h = pp->h_scu;
    const int scuw = (1 << it->log2_cuw) >> 2, scuh = (1 << it->log2_cuh) >> 2, scup = x + y * w, l = it->lidx;
#define M_COD(p) ((map_scu[p] >> 31) & 1)
#define M_IF(p)  ((map_scu[p] >> 15) & 1)
#define M_IBC(p) ((map_scu[p] >> 26) & 1)
#define M_INTER(p) (!M_IF(p) && !M_IBC(p))
    unsigned av = 0;
    if(x > 0 && M_INTER(scup - 1) && M_COD(scup - 1)) {
        av |= 1u << 1;                                                                       /* AVAIL_LE */
        if(y + scuh < h && M_COD(scup + scuh * w - 1) && M_INTER(scup + scuh * w - 1)) av |= 1u << 7;  /* LO_LE */
    }
    if(y > 0) {
        if(M_INTER(scup - w)) av |= 1u << 0;                                                 /* AVAIL_UP: no COD test */
        if(M_INTER(scup - w + scuw - 1)) av |= 1u << 9;                                      /* AVAIL_RI_UP */
        if(x > 0 && M_INTER(scup - w - 1) && M_COD(scup - w - 1)) av |= 1u << 5;             /* UP_LE */
        /* MCU_IS_COD_NIF: coded and not intra (the IBC flag is not part of this test) */
        if(x + scuw < w && M_COD(scup - w + scuw) && !M_IF(scup - w + scuw)) av |= 1u << 6;  /* UP_RI */
    }
    if(x + scuw < w && M_INTER(scup + scuw) && M_COD(scup + scuw)) {
        av |= 1u << 3;                                                                       /* AVAIL_RI */
        if(y + scuh < h && M_COD(scup + scuh * w + scuw) && M_INTER(scup + scuh * w + scuw)) av |= 1u << 8;  /* LO_RI */
    }
    it->avail = (uint16_t)av;
    const int nb[3] = {scup - 1, scup - w, scup - w + scuw}, need[3] = {1, 0, 6};
    for(int k = 0; k < 3; k++) {
        it->refi[k] = 0;
        if((av >> need[k]) & 1) { it->mvp[k][0] = map_mv[(nb[k] * 2 + l) * 2]; it->mvp[k][1] = map_mv[(nb[k] * 2 + l) * 2 + 1]; }
        else it->mvp[k][0] = it->mvp[k][1] = 1;
    }
    const int16_t *col = l ? col_mv1 : col_mv0;                  /* refp[0][lidx].map_mv[scup][0] */
    it->refi[3] = 0;
    it->mvp[3][0] = col[(scup * 2 + 0) * 2]; it->mvp[3][1] = col[(scup * 2 + 0) * 2 + 1];
    /* temporal direct: colocated MV (list 0 entry) of the list-1 reference at the CU's bottom-right SCU */
    const int br = scup + (scuw - 1) + (scuh - 1) * w;
    const int mvc[2] = {col_mv1[(br * 2 + 0) * 2], col_mv1[(br * 2 + 0) * 2 + 1]};
    const int dpoc_co = pp->ref_poc[1] - pp->col_list_poc0, d0 = pp->poc - pp->ref_poc[0], d1 = pp->ref_poc[1] - pp->poc;
    for(int k = 0; k < 2; k++) {
        it->mv_dir[0][k] = (int16_t)(dpoc_co ? d0 * mvc[k] / dpoc_co : 0);
        it->mv_dir[1][k] = (int16_t)(dpoc_co ? -d1 * mvc[k] / dpoc_co : 0);
    }
#undef M_COD
#undef M_IF
#undef M_IBC
#undef M_INTER
}
void xo_mvp_batch(xb200_mvp_item *items, int64_t n, const xb200_mvp_pic *pp, const uint32_t *map_scu, const int16_t *map_mv,
                  const int16_t *col_mv0, const int16_t *col_mv1)
{
    for(int64_t i = 0; i < n; i++) xo_mvp(&items[i], pp, map_scu, map_mv, col_mv0, col_mv1);
}

/* xeve_get_avail_intra (src_base/xeve_util.c:717-772), xeve_get_nbr (src_base/xeve_ipred.c:33-97) for the three planes and
 * xeve_get_mpm (:230-252), single tile.  y/u/v: active-area origins of the picture reconstructed so far. */
void xo_intra_nbr(const int16_t *y, const int16_t *u, const int16_t *v, int s_l, int s_c, xb200_nbr_item *it, const uint32_t *map_scu,
                  const int8_t *map_ipm, int w_scu, int h_scu, int cip, int bd, int16_t *side)
{
    const uint8_t (*mpm_tbl)[6][5] = xo_mpm_tbl;
    const int xs = it->x >> 2, ys = it->y >> 2, scuw = (1 << it->log2_cuw) >> 2, scuh = (1 << it->log2_cuh) >> 2;
    const int scup = xs + ys * w_scu, half = 1 << (bd - 1);
#define COD(p) ((map_scu[p] >> 31) & 1)
#define IFL(p) ((map_scu[p] >> 15) & 1)
    unsigned av = 0;
    if(xs > 0 && COD(scup - 1)) {
        av |= 1u << 1;
        if(ys + scuh + scuw - 1 < h_scu && COD(scup + w_scu * (scuw + scuh) - w_scu - 1)) av |= 1u << 7;
    }
    if(ys > 0) {
        av |= (1u << 0) | (1u << 9);
        if(xs > 0 && COD(scup - w_scu - 1)) av |= 1u << 5;
        if(xs + scuw < w_scu && COD(scup - w_scu + scuw)) av |= 1u << 6;
    }
    if(xs + scuw < w_scu && COD(scup + scuw)) {
        av |= 1u << 3;
        if(ys + scuh + scuw - 1 < h_scu && COD(scup + w_scu * (scuw + scuh - 1) + scuw)) av |= 1u << 8;
    }
    it->avail = (uint16_t)av;
    int ipm_l = 0, ipm_u = 0;
    if(xs > 0 && IFL(scup - 1) && COD(scup - 1)) ipm_l = map_ipm[scup - 1] + 1;
    if(ys > 0 && IFL(scup - w_scu) && COD(scup - w_scu)) ipm_u = map_ipm[scup - w_scu] + 1;
    memcpy(it->mpm, mpm_tbl[ipm_l][ipm_u], 5);
    int16_t *out = side + it->nb_off;
    for(int c = 0; c < 3; c++) {
        const int      n = c ? (1 << it->log2_cuw) >> 1 : 1 << it->log2_cuw, unit = c ? 2 : 4, s = c ? s_c : s_l;
        const int16_t *src = (c == 0 ? y : c == 1 ? u : v) + (c ? (it->y >> 1) * s + (it->x >> 1) : it->y * s + it->x);
        int16_t       *left = out + 1, *up = out + (2 * n + 1) + 1;
        up[-1] = (int16_t)(((av >> 5) & 1) && (!cip || IFL(scup - w_scu - 1)) ? src[-s - 1] : half);
        for(int i = 0; i < scuw + scuh; i++) {
            const int ok = ys > 0 && xs + i < w_scu && COD(scup - w_scu + i) && (!cip || IFL(scup - w_scu + i));
            for(int k = 0; k < unit; k++) up[i * unit + k] = (int16_t)(ok ? src[-s + i * unit + k] : half);
        }
        for(int i = 0; i < scuh + scuw; i++) {
            const int ok = xs > 0 && ys + i < h_scu && COD(scup - 1 + i * w_scu) && (!cip || IFL(scup - 1 + i * w_scu));
            for(int k = 0; k < unit; k++) left[i * unit + k] = (int16_t)(ok ? src[(i * unit + k) * s - 1] : half);
        }
        left[-1] = up[-1];
        out += 2 * (2 * n + 1);
    }
#undef COD
#undef IFL
}
void xo_intra_nbr_batch(const int16_t *y, const int16_t *u, const int16_t *v, int s_l, int s_c, xb200_nbr_item *items, int64_t n,
                        const uint32_t *map_scu, const int8_t *map_ipm, int w_scu, int h_scu, int cip, int bd, int16_t *side)
{
    for(int64_t i = 0; i < n; i++) xo_intra_nbr(y, u, v, s_l, s_c, &items[i], map_scu, map_ipm, w_scu, h_scu, cip, bd, side);
}


/* ---------------------------------------------------------------------------------------------
 * CU decision chain of one picture (SURVEY 8a' q15 / q16, the caller of rows a1 and f-3): mode_analyze_lcu ->
 * mode_coding_tree -> mode_coding_unit -> mode_check_inter / mode_check_intra (src_base/xeve_mode.c:1170-1348, 2007-2374,
 * 2521-2608) with the bookkeeping around them: init_cu_data / copy_cu_data / copy_to_cu_data (:375-632, 868-1034),
 * update_map_scu / clear_map_scu (:1036-1155), mode_cpy_rec_to_ref (:797-866), the split_cu_flag count of
 * xeve_eco_split_mode (src_base/xeve_eco.c:1377-1429) and the CTU loop of xeve_ctu_mt_core (src_base/xeve_enc.c:103-175,
 * one thread, one tile, no delta QP).  Baseline: quad split only, the context-index flags are all 0, and the coder state
 * a CTU starts from equals the state the previous CTU's decision pass ended with (checked against the reference's own
 * per-CTU states in tests/test_oracle.py), so a picture needs no input besides its original, its references and the
 * colocated MV maps.
 * ------------------------------------------------------------------------------------------- */
#define XO_MAX_COST 1.7e+308
enum { XO_MODE_SKIP = 0, XO_MODE_DIR = 1, XO_MODE_INTER = 2, XO_MODE_INTRA = 3 };
typedef struct {              /* XEVE_CU_DATA of one quad-tree level, SCU-granular, row stride = block width in SCUs */
    uint8_t  mode[256], log2[256];
    int8_t   ipm[256], refi[256][2];
    int16_t  mv[256][2][2];
    int32_t  nnz[256][3];
    uint8_t  mvp_idx[256][2];
    int16_t  mvd[256][2][2];
    uint32_t scu[256];
    int16_t *rec[3], *coef[3]; /* reconstruction and quantised coefficients, CU rectangles at their place, stride = block width */
} cud_t;
typedef struct {
    const xb200_seq *sq; const xo_planes *pl; const xo_ctu_rec *pp;
    const int16_t *col[2];
    int       w, h, w_scu, h_scu;
    int16_t  *rec[3]; int s_l, s_c;
    uint32_t *map_scu; int8_t *map_ipm, *map_refi; int16_t *map_mv;
    xo_state  curr[5], next[5], before[5];
    cud_t     best[5], temp[5];
    int       cu_mode, dist_cu_best;
    int16_t  *coef, *rec_cu, *pred_y, *side;
    xb200_cu_item *cu_log; xb200_intra_item *intra_log; int64_t cu_cap, intra_cap, n_cu, n_intra;
} chain_t;

/* Optional stand-ins for the two per-CU analyses (tests plug the CUDA library in here, so that the operators under test run inside
 * the real decision chain): same records as xo_analyze_cu / xo_analyze_intra with one item, states[0] in / states[1] out. */
static xo_chain_cu_fn    g_chain_cu_fn;
static xo_chain_intra_fn g_chain_intra_fn;
static xo_chain_mvp_fn   g_chain_mvp_fn;    /* ... and for their inputs: xo_mvp / xo_intra_nbr of the CU under analysis */
static xo_chain_nbr_fn   g_chain_nbr_fn;
void xo_chain_set_callbacks(xo_chain_cu_fn cu_fn, xo_chain_intra_fn intra_fn) { g_chain_cu_fn = cu_fn; g_chain_intra_fn = intra_fn; }
void xo_chain_set_input_callbacks(xo_chain_mvp_fn mvp_fn, xo_chain_nbr_fn nbr_fn) { g_chain_mvp_fn = mvp_fn; g_chain_nbr_fn = nbr_fn; }

static void cud_init(cud_t *d, int L)       /* init_cu_data: everything the chain reads later */
{
    const int n = 1 << (2 * L);
    memset(d->mode, 0, n); memset(d->log2, 0, n); memset(d->ipm, 0, n); memset(d->refi, 0, 2 * n);
    memset(d->mv, 0, 8 * n); memset(d->nnz, 0, 12 * n); memset(d->scu, 0, 4 * n); memset(d->mvp_idx, 0, 2 * n); memset(d->mvd, 0, 8 * n);
}
static void cud_copy(cud_t *dst, int Ld, const cud_t *src, int Ls, int xoff, int yoff) /* copy_cu_data */
{
    const int ns = 1 << Ls, nd = 1 << Ld, ws = 4 << Ls, wd = 4 << Ld;
    for(int j = 0; j < ns; j++) {
        const int d = ((yoff >> 2) + j) * nd + (xoff >> 2), s0 = j * ns;
        memcpy(dst->mode + d, src->mode + s0, ns); memcpy(dst->log2 + d, src->log2 + s0, ns); memcpy(dst->ipm + d, src->ipm + s0, ns);
        memcpy(dst->refi + d, src->refi + s0, 2 * ns); memcpy(dst->mv + d, src->mv + s0, 8 * ns);
        memcpy(dst->nnz + d, src->nnz + s0, 12 * ns); memcpy(dst->scu + d, src->scu + s0, 4 * ns);
        memcpy(dst->mvp_idx + d, src->mvp_idx + s0, 2 * ns); memcpy(dst->mvd + d, src->mvd + s0, 8 * ns);
    }
    for(int j = 0; j < ws; j++) {
        memcpy(dst->rec[0] + (yoff + j) * wd + xoff, src->rec[0] + j * ws, 2 * ws);
        memcpy(dst->coef[0] + (yoff + j) * wd + xoff, src->coef[0] + j * ws, 2 * ws);
    }
    for(int c = 1; c < 3; c++)
        for(int j = 0; j < ws / 2; j++) {
            memcpy(dst->rec[c] + ((yoff >> 1) + j) * (wd / 2) + (xoff >> 1), src->rec[c] + j * (ws / 2), ws);
            memcpy(dst->coef[c] + ((yoff >> 1) + j) * (wd / 2) + (xoff >> 1), src->coef[c] + j * (ws / 2), ws);
        }
}
static void chain_clear_map(chain_t *k, int x, int y, int cuw)                          /* clear_map_scu */
{
    const int w = ((x + cuw > k->w ? k->w - x : cuw) >> 2), h = ((y + cuw > k->h ? k->h - y : cuw) >> 2);
    for(int j = 0; j < h; j++) memset(k->map_scu + ((y >> 2) + j) * k->w_scu + (x >> 2), 0, 4 * w);
}
static void chain_update_map(chain_t *k, int x, int y, int L)                           /* update_map_scu from cu_data_best[L] */
{
    const cud_t *b = &k->best[L];
    const int cuw = 4 << L, n = 1 << L, w = ((x + cuw > k->w ? k->w - x : cuw) >> 2), h = ((y + cuw > k->h ? k->h - y : cuw) >> 2);
    for(int j = 0; j < h; j++) {
        const int64_t p = (int64_t)((y >> 2) + j) * k->w_scu + (x >> 2);
        memcpy(k->map_scu + p, b->scu + j * n, 4 * w); memcpy(k->map_ipm + p, b->ipm + j * n, w);
        memcpy(k->map_mv + p * 4, b->mv + j * n, 8 * w); memcpy(k->map_refi + p * 2, b->refi + j * n, 2 * w);
    }
}
static void chain_rec_to_pic(chain_t *k, int x, int y, int L)                           /* mode_cpy_rec_to_ref from cu_data_best[L] */
{
    const cud_t *b = &k->best[L];
    const int cuw = 4 << L, w = x + cuw > k->w ? k->w - x : cuw, h = y + cuw > k->h ? k->h - y : cuw;
    for(int j = 0; j < h; j++) memcpy(k->rec[0] + (int64_t)(y + j) * k->s_l + x, b->rec[0] + j * cuw, 2 * w);
    for(int c = 1; c < 3; c++)
        for(int j = 0; j < h / 2; j++) memcpy(k->rec[c] + (int64_t)(y / 2 + j) * k->s_c + x / 2, b->rec[c] + j * (cuw / 2), w);
}
/* copy_to_cu_data into cu_data_temp[L] */
static void chain_store_cu(chain_t *k, int L, int mode, int ipm, const int8_t refi[2], int16_t mv[2][2], const int32_t nnz[3],
                           const uint8_t mvp_idx[2], int16_t mvd[2][2], const int16_t *coef)
{
    cud_t *t = &k->temp[L];
    const int n = 1 << (2 * L), ny = 16 << (2 * L);
    const uint32_t word = (uint32_t)(((uint32_t)k->pp->tile_qp << 16) | ((uint32_t)(mode == XO_MODE_INTRA) << 15) | (1u << 31) |
                                     ((uint32_t)(mode == XO_MODE_SKIP) << 23));
    for(int i = 0; i < n; i++) {
        t->mode[i] = (uint8_t)mode; t->log2[i] = (uint8_t)(L + 2); memcpy(t->nnz[i], nnz, 12); t->scu[i] = word;
        if(mode == XO_MODE_INTRA) { t->ipm[i] = (int8_t)ipm; memset(t->mv[i], 0, 8); t->refi[i][0] = t->refi[i][1] = -1; }
        else {
            t->refi[i][0] = refi[0]; t->refi[i][1] = refi[1]; memcpy(t->mv[i], mv, 8);
            memcpy(t->mvp_idx[i], mvp_idx, 2); memcpy(t->mvd[i], mvd, 8);
        }
    }
    memcpy(t->rec[0], k->rec_cu, 2 * ny); memcpy(t->rec[1], k->rec_cu + ny, ny / 2); memcpy(t->rec[2], k->rec_cu + ny + ny / 4, ny / 2);
    memcpy(t->coef[0], coef, 2 * ny); memcpy(t->coef[1], coef + ny, ny / 2); memcpy(t->coef[2], coef + ny + ny / 4, ny / 2);
}
static uint32_t chain_split_flag(xo_state *st, int cuw, int split)   /* xeve_eco_split_mode in bit-count mode; returns the bits */
{
    if(cuw < 8) return 0;
    cabac_t c;
    c.s = st->s; c.bits = 0; c.ipm[0] = st->split;       /* borrow the first spare model slot for ctx.split_cu_flag */
    cb_bin(&c, XO_CM_IPM, split);
    st->s = c.s; st->split = c.ipm[0];
    return c.bits;
}
static double chain_unit(chain_t *k, int x, int y, int L, int cud)   /* mode_coding_unit */
{
    const xo_ctu_rec *pp = k->pp;
    const int log2 = L + 2, cuw = 1 << log2, ny = cuw * cuw, B = pp->slice_type == 0;
    xb200_rates rates;
    xb200_sbac  st[2];
    double      cost_best = XO_MAX_COST;
    int         nnz[3] = {0, 0, 0};
    (void)cud;
    xo_rdoq_rates(&k->curr[L].s, 1, &rates);                          /* mode_cu_init -> xeve_rdoq_bit_est */
    k->cu_mode = XO_MODE_INTRA;
    if(pp->slice_type != 2) {                                         /* mode_check_inter */
        xb200_cu_item cu;
        memset(&cu, 0, sizeof(cu));
        cu.poc = pp->poc; cu.cur_pic = pp->cur_pic; cu.x = (int16_t)x; cu.y = (int16_t)y; cu.log2_cuw = cu.log2_cuh = (uint8_t)log2;
        cu.slice_type = (uint8_t)pp->slice_type; cu.all_preds = 1;
        cu.max_search_range = pp->max_search_range; cu.lambda_mv = pp->lambda_mv;
        for(int l = 0; l < 2; l++) {
            cu.num_refp[l] = (uint8_t)pp->num_refp[l];
            for(int r = 0; r < 4; r++) { cu.ref_pic[l][r] = pp->ref_pic[l][r]; cu.ref_poc[l][r] = pp->ref_poc[l][r]; }
        }
        for(int i = 0; i < 3; i++) { cu.qp[i] = (uint8_t)pp->qp[i]; cu.lambda[i] = pp->lambda[i]; }
        cu.dist_chroma_weight[0] = pp->dist_chroma_weight[0]; cu.dist_chroma_weight[1] = pp->dist_chroma_weight[1];
        xb200_mvp_pic mp = {k->w_scu, k->h_scu, pp->poc, {pp->ref_poc[0][0], pp->ref_poc[1][0]}, pp->col_list_poc0};
        for(int l = 0; l < (B ? 2 : 1); l++) {
            xb200_mvp_item mi;
            memset(&mi, 0, sizeof(mi));
            mi.x_scu = (int16_t)(x >> 2); mi.y_scu = (int16_t)(y >> 2); mi.log2_cuw = mi.log2_cuh = (uint8_t)log2; mi.lidx = (uint8_t)l;
            if(g_chain_mvp_fn) g_chain_mvp_fn(&mi, &mp);
            else xo_mvp(&mi, &mp, k->map_scu, k->map_mv, k->col[0], k->col[1] ? k->col[1] : k->col[0]);
            memcpy(cu.mvp[l], mi.mvp, sizeof(mi.mvp)); memcpy(cu.refi_pred[l], mi.refi, 4);
            if(B) memcpy(cu.mv_dir, mi.mv_dir, sizeof(mi.mv_dir));
        }
        cu.rate_idx = 0; cu.state_in = 0; cu.state_out = 1; cu.out_off = 0;
        st[0] = k->curr[L].s;
        if(g_chain_cu_fn) g_chain_cu_fn(&cu, st, &rates, k->coef, k->rec_cu, k->pred_y);
        else {
            g_pred_y_sink = k->pred_y;
            xo_analyze_cu(k->sq, k->pl, &rates, &cu, st, k->coef, k->rec_cu);
            g_pred_y_sink = NULL;
        }
        if(k->n_cu < k->cu_cap) k->cu_log[k->n_cu] = cu;
        k->n_cu++;
        k->cu_mode = cu.best_idx == 3 ? XO_MODE_SKIP : cu.best_idx == 4 ? XO_MODE_DIR : XO_MODE_INTER;
        nnz[0] = cu.nnz[0]; nnz[1] = cu.nnz[1]; nnz[2] = cu.nnz[2];
        k->next[L] = k->curr[L]; k->next[L].s = st[1];                /* SBAC_STORE(s_next_best, s_temp_best), src_base/xeve_pinter.c */
        if(cu.cost < cost_best) {
            cost_best = cu.cost;
            chain_store_cu(k, L, k->cu_mode, 0, cu.refi, cu.mv, cu.nnz, cu.mvp_idx, cu.mvd, k->coef);
        }
    }
    if(pp->slice_type == 2 || nnz[0] || nnz[1] || nnz[2] || cost_best == XO_MAX_COST) {   /* mode_check_intra */
        const xo_planes *o = &k->pl[pp->cur_pic];
        xb200_intra_item it;
        xb200_nbr_item   nb;
        memset(&it, 0, sizeof(it)); memset(&nb, 0, sizeof(nb));
        k->dist_cu_best = 0x7fffffff;
        it.inter_satd = cost_best != XO_MAX_COST
            ? (uint32_t)xo_satd(cuw, cuw, o->y + (int64_t)y * o->s_l + x, o->s_l, k->pred_y, cuw, k->sq->bit_depth) : 0xffffffffu;
        nb.x = (int16_t)x; nb.y = (int16_t)y; nb.log2_cuw = nb.log2_cuh = (uint8_t)log2; nb.nb_off = 0;
        if(g_chain_nbr_fn) g_chain_nbr_fn(&nb, k->side);
        else xo_intra_nbr(k->rec[0], k->rec[1], k->rec[2], k->s_l, k->s_c, &nb, k->map_scu, k->map_ipm, k->w_scu, k->h_scu, pp->cip,
                          k->sq->bit_depth, k->side);
        it.poc = pp->poc; it.cur_pic = pp->cur_pic; it.x = (int16_t)x; it.y = (int16_t)y; it.log2_cuw = it.log2_cuh = (uint8_t)log2;
        it.slice_type = (uint8_t)pp->slice_type; it.all_preds = 1;
        for(int i = 0; i < 3; i++) { it.qp[i] = (uint8_t)pp->qp[i]; it.lambda[i] = pp->lambda[i]; }
        memcpy(it.mpm, nb.mpm, 5);
        it.rate_idx = 0; it.state_in = 0; it.state_out = 1;
        it.cm_ipm_in[0] = k->curr[L].ipm[0]; it.cm_ipm_in[1] = k->curr[L].ipm[1];
        it.sqrt_lambda0 = pp->sqrt_lambda0;
        it.dist_chroma_weight[0] = pp->dist_chroma_weight[0]; it.dist_chroma_weight[1] = pp->dist_chroma_weight[1];
        it.nb_off = 0; it.out_off = 0;
        st[0] = k->curr[L].s; st[1] = st[0];
        /* the inter winner's reconstruction already sits in cu_data_temp; the intra trial works in its own buffers */
        int16_t *coef_i = k->coef + 3 * ny / 2 + 64, *rec_i = k->rec_cu + 3 * ny / 2 + 64;
        if(g_chain_intra_fn) g_chain_intra_fn(&it, st, &rates, k->side, coef_i, rec_i);
        else xo_analyze_intra(k->sq, k->pl, &rates, &it, st, k->side, coef_i, rec_i);
        if(k->n_intra < k->intra_cap) k->intra_log[k->n_intra] = it;
        k->n_intra++;
        if(it.cost < cost_best) {
            cost_best = it.cost;
            k->cu_mode = XO_MODE_INTRA;
            k->next[L] = k->curr[L]; k->next[L].s = st[1]; k->next[L].ipm[0] = it.cm_ipm_out[0]; k->next[L].ipm[1] = it.cm_ipm_out[1];
            k->dist_cu_best = it.dist_cu;
            memcpy(k->rec_cu, rec_i, sizeof(int16_t) * 3 * ny / 2);
            int8_t  none[2] = {-1, -1};
            int16_t zero[2][2] = {{0, 0}, {0, 0}};
            chain_store_cu(k, L, XO_MODE_INTRA, it.ipm[0], none, zero, it.nnz, NULL, zero, coef_i);
        }
    }
    return cost_best;
}
static double chain_tree(chain_t *k, int x0, int y0, int L, int cud, int next_split)     /* mode_coding_tree */
{
    const xo_ctu_rec *pp = k->pp;
    const int    cuw = 4 << L, log2 = L + 2, intra_slice = pp->slice_type == 2;
    const int    boundary = !(x0 + cuw <= k->w && y0 + cuw <= k->h);
    const int    check_max = intra_slice ? pp->max_cu_intra : pp->max_cu_inter, check_min = intra_slice ? pp->min_cu_intra : pp->min_cu_inter;
    const double lambda0 = pp->lambda[0];
    double       cost_best = XO_MAX_COST, cost_temp;
    xo_state     s_depth;
    memset(&s_depth, 0, sizeof(s_depth));
    k->before[L] = k->curr[L];
    if(!boundary && cuw <= check_max) {
        cost_temp = 0.0;
        if(cuw > 4) {
            const uint32_t bits = chain_split_flag(&k->curr[L], cuw, 0);
            cost_temp += (double)bits * lambda0;
        }
        cud_init(&k->temp[L], L);
        chain_clear_map(k, x0, y0, cuw);
        double cost_dqp = cost_temp;
        cost_dqp += chain_unit(k, x0, y0, L, cud);
        int cu_mode_dqp = 0, dist_dqp = 0;
        if(cost_best > cost_dqp) {
            cu_mode_dqp = k->cu_mode; dist_dqp = k->dist_cu_best;
            cud_copy(&k->best[L], L, &k->temp[L], L, 0, 0);
            cost_best = cost_dqp;
            s_depth = k->next[L];
            chain_rec_to_pic(k, x0, y0, L);
        }
        k->cu_mode = cu_mode_dqp; k->dist_cu_best = dist_dqp;
    }
    if(cost_best != XO_MAX_COST && cud >= ((pp->poc % 2) ? 2 : 4) && k->cu_mode == XO_MODE_SKIP) next_split = 0;   /* ENC_ECU_ADAPTIVE */
    if(cost_best != XO_MAX_COST && intra_slice) {
        const int dist_cu = k->dist_cu_best, dist_th = 1 << (2 * log2 + 7);
        if(dist_cu < dist_th) {
            const uint8_t inc = (uint8_t)((2 * log2 >= 6 ? 2 : 0) + 8);
            if(dist_cu < lambda0 * inc) next_split = 0;
        }
    }
    if(cuw > 4 && next_split && cuw > check_min) {
        cud_init(&k->temp[L], L);
        chain_clear_map(k, x0, y0, cuw);
        cost_temp = 0.0;
        k->curr[L] = k->before[L];
        cost_temp += (double)chain_split_flag(&k->curr[L], cuw, 1) * lambda0;
        const int half = cuw >> 1;
        for(int part = 0; part < 4; part++) {
            const int xp = x0 + (part & 1) * half, yp = y0 + (part >> 1) * half;
            if(xp < k->w && yp < k->h) {
                k->curr[L - 1] = part == 0 ? k->curr[L] : k->next[L - 1];
                cost_temp += chain_tree(k, xp, yp, L - 1, cud + 2, 1);      /* xeve_split_get_part_structure: a quad split is two levels */
                cud_copy(&k->temp[L], L, &k->best[L - 1], L - 1, xp - x0, yp - y0);
                chain_update_map(k, xp, yp, L - 1);
            }
        }
        if(cost_best - 0.0001 > cost_temp) {
            cud_copy(&k->best[L], L, &k->temp[L], L, 0, 0);
            cost_best = cost_temp;
            s_depth = k->next[L - 1];
        }
    }
    chain_rec_to_pic(k, x0, y0, L);
    k->next[L] = s_depth;
    return cost_best > XO_MAX_COST ? XO_MAX_COST : cost_best;
}
static void chain_leaves(const chain_t *k, int xc, int yc, int x, int y, int L, xb200_df_cu *cus, int64_t cap, int64_t *n)
{
    if(x >= k->w || y >= k->h) return;
    const int idx = ((y - yc) >> 2) * 16 + ((x - xc) >> 2);
    if(k->best[4].log2[idx] == L + 2 || L == 0) {
        if(*n < cap) { cus[*n].x = (int16_t)x; cus[*n].y = (int16_t)y; cus[*n].log2_cuw = cus[*n].log2_cuh = (uint8_t)(L + 2); }
        (*n)++;
        return;
    }
    const int half = 2 << L;
    for(int part = 0; part < 4; part++) chain_leaves(k, xc, yc, x + (part & 1) * half, y + (part >> 1) * half, L - 1, cus, cap, n);
}
/* xeve_eco_tree + xeve_eco_unit (src_base/xeve_enc.c:35-101, src_base/xeve_eco.c:1431-1603) over the decided CTU, models and
 * range only: the state the next CTU's decision pass is loaded from (src_base/xeve_enc.c:139).  In B and I slices it equals the
 * decision pass's own end state; in P slices it does not (the RDO counter codes a direct_mode_flag the bitstream lacks). */
static void chain_eco(const chain_t *k, xo_state *st, int xc, int yc, int x, int y, int L)
{
    if(x >= k->w || y >= k->h) return;
    const cud_t *b = &k->best[4];
    const int    idx = ((y - yc) >> 2) * 16 + ((x - xc) >> 2), cuw = 4 << L;
    if(b->log2[idx] != L + 2 && L > 0) {
        chain_split_flag(st, cuw, 1);
        for(int part = 0; part < 4; part++) chain_eco(k, st, xc, yc, x + (part & 1) * (cuw >> 1), y + (part >> 1) * (cuw >> 1), L - 1);
        return;
    }
    if(cuw > 4) chain_split_flag(st, cuw, 0);
    const int slice = k->pp->slice_type, B = slice == 0, mode = b->mode[idx], ny = cuw * cuw;
    cabac_t c;
    c.s = st->s; c.bits = 0; c.ipm[0] = st->ipm[0]; c.ipm[1] = st->ipm[1];
    int16_t *coef = malloc(sizeof(int16_t) * 3 * ny / 2);                            /* coef_rect_to_series */
    for(int j = 0; j < cuw; j++) memcpy(coef + j * cuw, b->coef[0] + (y - yc + j) * 64 + (x - xc), 2 * cuw);
    for(int p = 1; p < 3; p++)
        for(int j = 0; j < cuw / 2; j++)
            memcpy(coef + ny + (p - 1) * (ny / 4) + j * (cuw / 2), b->coef[p] + ((y - yc) / 2 + j) * 32 + (x - xc) / 2, cuw);
    if(slice != 2) {
        cb_bin(&c, XB200_CM_SKIP_FLAG, mode == XO_MODE_SKIP);
        if(mode == XO_MODE_SKIP) {
            cb_mvp_idx(&c, b->mvp_idx[idx][0]);
            if(B) cb_mvp_idx(&c, b->mvp_idx[idx][1]);
        }
        else {
            cb_bin(&c, XB200_CM_PRED_MODE, mode == XO_MODE_INTRA);
            if(mode != XO_MODE_INTRA) {
                if(B) cb_bin(&c, XB200_CM_DIRECT, mode == XO_MODE_DIR);
                if(mode != XO_MODE_DIR) {
                    const int r0 = b->refi[idx][0], r1 = b->refi[idx][1];
                    if(B) {                                                            /* xeve_eco_inter_pred_idc */
                        if(r0 >= 0 && r1 >= 0) cb_bin(&c, XB200_CM_INTER_DIR, 0);
                        else { cb_bin(&c, XB200_CM_INTER_DIR, 1); cb_bin(&c, XB200_CM_INTER_DIR + 1, r0 >= 0 ? 0 : 1); }
                    }
                    if(r0 >= 0) { cb_refi(&c, k->pp->num_refp[0], r0); cb_mvp_idx(&c, b->mvp_idx[idx][0]); cb_mvd(&c, b->mvd[idx][0]); }
                    if(B && r1 >= 0) { cb_refi(&c, k->pp->num_refp[1], r1); cb_mvp_idx(&c, b->mvp_idx[idx][1]); cb_mvd(&c, b->mvd[idx][1]); }
                }
            }
        }
    }
    if(mode == XO_MODE_INTRA) {
        xb200_nbr_item nb;
        memset(&nb, 0, sizeof(nb));
        nb.x = (int16_t)x; nb.y = (int16_t)y; nb.log2_cuw = nb.log2_cuh = (uint8_t)(L + 2);
        xo_intra_nbr(k->rec[0], k->rec[1], k->rec[2], k->s_l, k->s_c, &nb, k->map_scu, k->map_ipm, k->w_scu, k->h_scu, k->pp->cip,
                     k->sq->bit_depth, k->side);                                        /* xeve_get_mpm */
        cb_unary(&c, nb.mpm[b->ipm[idx]], XO_CM_IPM);
        cb_coef_intra(&c, b->nnz[idx], L + 2, coef, 7);
    }
    else if(mode != XO_MODE_SKIP) {
        xb200_bits_item it;
        memset(&it, 0, sizeof(it));
        it.log2_cuw = it.log2_cuh = (uint8_t)(L + 2); memcpy(it.nnz, b->nnz[idx], 12);
        cb_coef(&c, &it, coef, 7);
    }
    free(coef);
    st->s = c.s; st->ipm[0] = c.ipm[0]; st->ipm[1] = c.ipm[1];
}
int xo_sizeof_chain(int what) { return what == 0 ? (int)sizeof(xo_ctu_rec) : what == 1 ? (int)sizeof(xo_state) : (int)sizeof(xo_scu_rec); }
/* One picture.  pp: picture-level inputs (the per-CTU fields are ignored); col_mv0/1: refp[0][REFP_0/1].map_mv; out[]: one record
 * per CTU with lcu_num / x_pel / y_pel / state_in / state_out filled; rec_*: the reconstruction before deblocking (active area);
 * maps: frame maps as they stand when the loop filter starts (COD, intra, QP, skip and luma-cbf bits of map_scu); cus: leaf CUs
 * in coding order; cu_log / intra_log: every inter / intra CU analysis in call order.  n_out: {leaf CUs, inter calls, intra calls}.
 * ctu_limit > 0 stops after that many CTUs.  scu_out / coef_out (optional): per CTU the 256 SCU records and the coefficient planes
 * (Y 64x64 | U 32x32 | V 32x32, CU rectangles in place) -- the contents of the reference's ctx->map_cu_data[lcu]. */
void xo_chain_picture(const xb200_seq *sq, const xo_planes *pl, const xo_ctu_rec *pp, const int16_t *col_mv0, const int16_t *col_mv1,
                      xo_ctu_rec *out, double *ctu_cost, int16_t *rec_y, int16_t *rec_u, int16_t *rec_v, int s_l, int s_c,
                      uint32_t *map_scu, int8_t *map_ipm, int8_t *map_refi, int16_t *map_mv, xb200_df_cu *cus, int64_t cus_cap,
                      xb200_cu_item *cu_log, int64_t cu_cap, xb200_intra_item *intra_log, int64_t intra_cap, int64_t *n_out,
                      int ctu_limit, xo_scu_rec *scu_out, int16_t *coef_out)
{
    chain_t *k = calloc(1, sizeof(chain_t));
    k->sq = sq; k->pl = pl; k->pp = pp; k->col[0] = col_mv0; k->col[1] = col_mv1;
    k->w = sq->w; k->h = sq->h; k->w_scu = (sq->w + 3) >> 2; k->h_scu = (sq->h + 3) >> 2;
    k->rec[0] = rec_y; k->rec[1] = rec_u; k->rec[2] = rec_v; k->s_l = s_l; k->s_c = s_c;
    k->map_scu = map_scu; k->map_ipm = map_ipm; k->map_refi = map_refi; k->map_mv = map_mv;
    k->cu_log = cu_log; k->cu_cap = cu_cap; k->intra_log = intra_log; k->intra_cap = intra_cap;
    for(int L = 0; L < 5; L++) {
        const int ny = 16 << (2 * L);
        for(int t = 0; t < 2; t++) {
            cud_t *d = t ? &k->temp[L] : &k->best[L];
            d->rec[0] = calloc((size_t)ny * 3, sizeof(int16_t)); d->rec[1] = d->rec[0] + ny; d->rec[2] = d->rec[1] + ny / 4;
            d->coef[0] = d->rec[2] + ny / 4; d->coef[1] = d->coef[0] + ny; d->coef[2] = d->coef[1] + ny / 4;
        }
    }
    k->coef = calloc(2 * (64 * 64 * 3 / 2 + 64), sizeof(int16_t)); k->rec_cu = calloc(2 * (64 * 64 * 3 / 2 + 64), sizeof(int16_t));
    k->pred_y = calloc(64 * 64, sizeof(int16_t)); k->side = calloc(8 * 64 + 6 + 16, sizeof(int16_t));
    const int64_t f = (int64_t)k->w_scu * k->h_scu;
    memset(map_scu, 0, 4 * f); memset(map_ipm, 0, f); memset(map_refi, 0, 2 * f); memset(map_mv, 0, 8 * f);
    /* xeve_sbac_reset with cm_init off: every model PROB_INIT, range 16384.  With threads = n the reference runs CTU rows y, y + n,
     * y + 2n, ... as one chain per thread, each reset at its first row (src_base/xeve_enc.c:103-175, 322-358; a CTU waits for its
     * upper-right neighbour only), so raster order with one state per chain reproduces it. */
    const int w_lcu = (k->w + 63) >> 6, h_lcu = (k->h + 63) >> 6;
    const int n_chain = pp->parallel_rows > 1 ? (pp->parallel_rows > h_lcu ? h_lcu : pp->parallel_rows) : 1;
    xo_state *chain_st = calloc((size_t)n_chain, sizeof(xo_state));
    for(int t = 0; t < n_chain; t++) {
        chain_st[t].s.range = 16384;
        for(int i = 0; i < XB200_CM_COUNT; i++) chain_st[t].s.m[i] = 512;
        chain_st[t].ipm[0] = chain_st[t].ipm[1] = chain_st[t].split = 512;
    }
    int64_t   n_leaf = 0;
    for(int lcu = 0; lcu < w_lcu * h_lcu && (ctu_limit <= 0 || lcu < ctu_limit); lcu++) {
        const int x = (lcu % w_lcu) << 6, y = (lcu / w_lcu) << 6;
        xo_state  st = chain_st[(lcu / w_lcu) % n_chain];
        out[lcu] = *pp;
        out[lcu].lcu_num = lcu; out[lcu].x_pel = x; out[lcu].y_pel = y; out[lcu].state_in = st;
        cud_init(&k->best[4], 4); cud_init(&k->temp[4], 4);                               /* mode_init_lcu */
        k->curr[4] = st;
        const double c = chain_tree(k, x, y, 4, 0, 1);
        if(ctu_cost) ctu_cost[lcu] = c;
        out[lcu].state_out = k->next[4];
        /* mode_analyze_lcu: update_to_ctx_map; then the bitstream pass marks every CU coded and sets the luma cbf bit
         * (xeve_eco_unit, src_base/xeve_eco.c:1575-1603) */
        chain_update_map(k, x, y, 4);
        const int ws = ((x + 64 > k->w ? k->w - x : 64) >> 2), hs = ((y + 64 > k->h ? k->h - y : 64) >> 2);
        for(int j = 0; j < hs; j++)
            for(int i = 0; i < ws; i++)
                if(k->best[4].nnz[j * 16 + i][0] > 0) k->map_scu[(int64_t)((y >> 2) + j) * k->w_scu + (x >> 2) + i] |= 1u << 24;
        chain_leaves(k, x, y, x, y, 4, cus, cus_cap, &n_leaf);
        chain_eco(k, &st, x, y, x, y, 4);
        chain_st[(lcu / w_lcu) % n_chain] = st;
        if(scu_out) {                                 /* what crosses the boundary to the host's entropy coder: XEVE_CU_DATA of the CTU */
            const cud_t *b = &k->best[4];
            for(int i = 0; i < 256; i++) {
                xo_scu_rec *o = &scu_out[(int64_t)lcu * 256 + i];
                memset(o, 0, sizeof(*o));
                o->mode = b->mode[i]; o->log2 = b->log2[i]; o->ipm = b->ipm[i]; memcpy(o->refi, b->refi[i], 2);
                memcpy(o->mvp_idx, b->mvp_idx[i], 2); memcpy(o->mv, b->mv[i], 8); memcpy(o->mvd, b->mvd[i], 8); memcpy(o->nnz, b->nnz[i], 12);
            }
        }
        if(coef_out) memcpy(coef_out + (int64_t)lcu * 6144, k->best[4].coef[0], sizeof(int16_t) * 6144);
    }
    n_out[0] = n_leaf; n_out[1] = k->n_cu; n_out[2] = k->n_intra;
    for(int L = 0; L < 5; L++) { free(k->best[L].rec[0]); free(k->temp[L].rec[0]); }
    free(k->coef); free(k->rec_cu); free(k->pred_y); free(k->side); free(k); free(chain_st);
}

/* FNV-1a over per-item output slots (the hash the harness records for in-situ results) */
void xo_hash_slots(const int16_t *buf, const int64_t *off, const int64_t *elems, int64_t n, uint64_t *out)
{
    for(int64_t i = 0; i < n; i++) {
        uint64_t       h = 0xcbf29ce484222325ULL;
        const uint8_t *p = (const uint8_t *)(buf + off[i]);
        for(int64_t k = 0; k < elems[i] * 2; k++) { h ^= p[k]; h *= 0x100000001b3ULL; }
        out[i] = h;
    }
}

/* batch drivers --------------------------------------------------------------------------------- */
void xo_me_batch(const xb200_seq *sq, const xo_planes *pl, const int16_t *side, xb200_me_item *items, int64_t n)
{
    for(int64_t i = 0; i < n; i++) xo_me(sq, pl, side, &items[i]);
}
void xo_mc_batch(const xb200_seq *sq, const xo_planes *pl, const xb200_mc_item *items, int64_t n, const int64_t *off,
                 int16_t *pred)
{
    for(int64_t i = 0; i < n; i++) xo_mc(sq, pl, &items[i], pred + off[i]);
}
void xo_tq_batch(const xb200_seq *sq, xb200_tq_item *items, int64_t n, const xb200_rates *rates, int16_t *coef,
                 int16_t *resi)
{
    for(int64_t i = 0; i < n; i++) {
        xo_tq(sq, &items[i], rates, coef + items[i].in_off, items[i].nnz);
        if(resi) {
            int sz = (3 << (items[i].log2_cuw + items[i].log2_cuh)) >> 1;
            memcpy(resi + items[i].in_off, coef + items[i].in_off, sizeof(int16_t) * sz);
            xo_itdq(sq, &items[i], resi + items[i].in_off, items[i].nnz);
        }
    }
}
void xo_residue_batch(const xb200_seq *sq, const xo_planes *pl, const xb200_rates *rates, xb200_residue_item *items,
                      int64_t n, int16_t *coef, int16_t *rec)
{
    for(int64_t i = 0; i < n; i++) xo_residue(sq, pl, rates, &items[i], coef + items[i].out_off, rec + items[i].out_off);
}
