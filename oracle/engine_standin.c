/* engine_standin.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A CPU table for include/xeve_b200_engine.h built from the oracle (xo_chain_picture + xo_deblock + xo_pad_plane), plus a small
 * driver of the xeve C API (push / encode loop over frames held in memory).  tests/test_dropin.py installs the table through
 * xeve_b200_set_engine() to pin the HOST plumbing of integration/xeve_b200_dropin.c -- shadow-context picture plans, enqueue
 * order, tail re-planning, record hand-over to the reference's entropy coder -- on machines without a GPU: the bitstream of the
 * drop-in library must equal the unmodified reference's.  The GPU tests run the same driver with the default (CUDA) table.
 * Only tests load this file (oracle/_ref/libengine_standin.so). */
#define _GNU_SOURCE
#include "xeve.h"
#include "../include/xeve_b200_engine.h"
#include "xeve_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define SI_MAXPIC 1024
typedef struct {
    int      used, padded;
    int16_t *buf[3];        /* allocation incl. padding */
    int      s[3], pad[3];
    int16_t *map_mv;        /* colocated motion of a decided picture */
    xo_scu_rec *scu;        /* results of the last decision into this picture */
    int16_t    *coef;
    int64_t     n_inter, n_intra;
    int         polls;
} SiPic;
struct xb200_ctx {          /* the stand-in's own context behind the opaque handle */
    xb200_seq seq;
    SiPic     pic[SI_MAXPIC];
    xo_planes planes[SI_MAXPIC];
    int       n_lcu, f_scu;
};

static int si_create(xb200_ctx **out, int device, const xb200_seq *seq)
{
    (void)device;
    xb200_ctx *c = calloc(1, sizeof(*c));
    if(!c) return XB200_ERR_OUT_OF_MEMORY;
    c->seq = *seq;
    c->n_lcu = ((seq->w + 63) >> 6) * ((seq->h + 63) >> 6);
    c->f_scu = ((seq->w + 3) >> 2) * ((seq->h + 3) >> 2);
    *out = c;
    return XB200_OK;
}
static void si_pic_free(SiPic *p)
{
    for(int q = 0; q < 3; q++) free(p->buf[q]);
    free(p->map_mv); free(p->scu); free(p->coef);
    memset(p, 0, sizeof(*p));
}
static void si_destroy(xb200_ctx *c)
{
    if(!c) return;
    for(int i = 0; i < SI_MAXPIC; i++) if(c->pic[i].used) si_pic_free(&c->pic[i]);
    free(c);
}
static int si_pic_create(xb200_ctx *c, int padded, int32_t *handle)
{
    for(int i = 0; i < SI_MAXPIC; i++) {
        SiPic *p = &c->pic[i];
        if(p->used) continue;
        p->used = 1; p->padded = padded;
        for(int q = 0; q < 3; q++) {
            const int w = q ? c->seq.w >> 1 : c->seq.w, h = q ? c->seq.h >> 1 : c->seq.h;
            p->pad[q] = padded ? (q ? XB200_PAD_C : XB200_PAD_L) : 0;
            p->s[q] = w + 2 * p->pad[q];
            p->buf[q] = calloc((size_t)p->s[q] * (h + 2 * p->pad[q]), sizeof(int16_t));
        }
        xo_planes *pl = &c->planes[i];
        pl->y = p->buf[0] + (size_t)p->pad[0] * p->s[0] + p->pad[0];
        pl->u = p->buf[1] + (size_t)p->pad[1] * p->s[1] + p->pad[1];
        pl->v = p->buf[2] + (size_t)p->pad[2] * p->s[2] + p->pad[2];
        pl->s_l = p->s[0]; pl->s_c = p->s[1]; pl->w_l = c->seq.w; pl->h_l = c->seq.h; pl->poc = 0;
        *handle = i;
        return XB200_OK;
    }
    return XB200_ERR_OUT_OF_MEMORY;
}
static int si_pic_destroy(xb200_ctx *c, int32_t h)
{
    if(h < 0 || h >= SI_MAXPIC || !c->pic[h].used) return XB200_ERR_INVALID_ARGUMENT;
    si_pic_free(&c->pic[h]);
    return XB200_OK;
}
static int16_t *si_plane(xb200_ctx *c, int h, int q) { return q == 0 ? c->planes[h].y : q == 1 ? c->planes[h].u : c->planes[h].v; }
static void si_pad(xb200_ctx *c, int h)
{
    SiPic *p = &c->pic[h];
    for(int q = 0; q < 3; q++)
        if(p->pad[q]) xo_pad_plane(p->buf[q], p->s[q], q ? c->seq.w >> 1 : c->seq.w, q ? c->seq.h >> 1 : c->seq.h, p->pad[q]);
}
static int si_pic_upload(xb200_ctx *c, int32_t h, const void *const planes[3], const int32_t stride_bytes[3], int in_bit_depth, int mem)
{
    (void)mem;
    if(h < 0 || h >= SI_MAXPIC || !c->pic[h].used) return XB200_ERR_INVALID_ARGUMENT;
    const int sh = c->seq.bit_depth - in_bit_depth;
    for(int q = 0; q < 3; q++) {
        const int w = q ? c->seq.w >> 1 : c->seq.w, hh = q ? c->seq.h >> 1 : c->seq.h;
        int16_t  *d = si_plane(c, h, q);
        for(int y = 0; y < hh; y++)
            for(int x = 0; x < w; x++) {
                const uint8_t *row = (const uint8_t *)planes[q] + (size_t)y * stride_bytes[q];
                const int      v = in_bit_depth <= 8 ? row[x] : ((const uint16_t *)row)[x];
                d[(size_t)y * c->pic[h].s[q] + x] = (int16_t)(sh >= 0 ? v << sh : v >> -sh);
            }
    }
    si_pad(c, h);
    return XB200_OK;
}
static int si_pic_download(xb200_ctx *c, int32_t h, int with_padding, int16_t *const planes[3], const int32_t stride_elems[3])
{
    if(h < 0 || h >= SI_MAXPIC || !c->pic[h].used || with_padding) return XB200_ERR_INVALID_ARGUMENT;
    for(int q = 0; q < 3; q++) {
        const int w = q ? c->seq.w >> 1 : c->seq.w, hh = q ? c->seq.h >> 1 : c->seq.h;
        for(int y = 0; y < hh; y++) memcpy(planes[q] + (size_t)y * stride_elems[q], si_plane(c, h, q) + (size_t)y * c->pic[h].s[q], (size_t)w * 2);
    }
    return XB200_OK;
}
static int si_analyze_picture(xb200_ctx *c, const xb200_picture *pp)
{
    const int h = pp->rec_pic;
    if(h < 0 || h >= SI_MAXPIC || !c->pic[h].used || !c->pic[h].padded) return XB200_ERR_INVALID_ARGUMENT;
    if(pp->slice_type == 1) return XB200_ERR_UNSUPPORTED;
    SiPic     *p = &c->pic[h];
    xo_ctu_rec r;
    memset(&r, 0, sizeof(r));
    r.poc = pp->poc; r.slice_type = pp->slice_type; r.tile_qp = pp->tile_qp; r.cur_pic = pp->cur_pic;
    const int16_t *col[2] = {NULL, NULL};
    for(int l = 0; l < 2; l++) {
        r.num_refp[l] = pp->num_refp[l];
        for(int k = 0; k < 4; k++) {
            r.ref_pic[l][k] = pp->ref_pic[l][k]; r.ref_poc[l][k] = pp->ref_poc[l][k];
            if(k == 0 && pp->ref_pic[l][k] >= 0) col[l] = c->pic[pp->ref_pic[l][k]].map_mv;
        }
    }
    r.col_list_poc0 = pp->col_list_poc0;
    r.max_cu_inter = pp->max_cu_inter; r.min_cu_inter = pp->min_cu_inter; r.max_cu_intra = pp->max_cu_intra; r.min_cu_intra = pp->min_cu_intra;
    r.cip = pp->cip;
    for(int i = 0; i < 3; i++) { r.qp[i] = pp->qp[i]; r.lambda[i] = pp->lambda[i]; }
    r.lambda_mv = pp->lambda_mv; r.max_search_range = pp->max_search_range; r.parallel_rows = pp->parallel_rows;
    r.sqrt_lambda0 = pp->sqrt_lambda0; r.dist_chroma_weight[0] = pp->dist_chroma_weight[0]; r.dist_chroma_weight[1] = pp->dist_chroma_weight[1];
    r.col_off[0] = r.col_off[1] = -1;
    const size_t f = (size_t)c->f_scu, n = (size_t)c->n_lcu;
    xo_ctu_rec  *out = calloc(n, sizeof(xo_ctu_rec));
    double      *cost = calloc(n, sizeof(double));
    uint32_t    *map_scu = calloc(f, 4);
    int8_t      *map_ipm = calloc(f, 1), *map_refi = calloc(f, 2);
    xb200_df_cu *cus = calloc(f, sizeof(xb200_df_cu));
    int64_t      n_out[3] = {0, 0, 0};
    if(!p->map_mv) p->map_mv = calloc(f, 8);
    if(!p->scu) p->scu = calloc(n * 256, sizeof(xo_scu_rec));
    if(!p->coef) p->coef = calloc(n * 6144, sizeof(int16_t));
    memset(p->map_mv, 0, f * 8);
    xo_chain_picture(&c->seq, c->planes, &r, col[0], col[1] ? col[1] : col[0], out, cost, c->planes[h].y, c->planes[h].u, c->planes[h].v,
                     p->s[0], p->s[1], map_scu, map_ipm, map_refi, p->map_mv, cus, (int64_t)f, NULL, 0, NULL, 0, n_out, 0, p->scu, p->coef);
    p->n_inter = n_out[1]; p->n_intra = n_out[2]; p->polls = 0;
    if(pp->deblock) {
        xb200_df_pic df = pp->df;
        df.w_scu = (c->seq.w + 3) >> 2; df.h_scu = (c->seq.h + 3) >> 2;
        xo_deblock(c->planes[h].y, c->planes[h].u, c->planes[h].v, p->s[0], p->s[1], c->seq.w, c->seq.h, cus, n_out[0], &df, map_scu, map_refi,
                   p->map_mv, c->seq.bit_depth);
    }
    si_pad(c, h);
    free(out); free(cost); free(map_scu); free(map_ipm); free(map_refi); free(cus);
    return XB200_OK;
}
static int si_picture_fetch(xb200_ctx *c, int32_t h, xb200_scu_rec *scu, int16_t *coef, xb200_state *st, double *cost, xb200_picture_stat *stat)
{
    (void)st; (void)cost;
    if(h < 0 || h >= SI_MAXPIC || !c->pic[h].used || !c->pic[h].scu) return XB200_ERR_INVALID_ARGUMENT;
    if(scu) memcpy(scu, c->pic[h].scu, (size_t)c->n_lcu * 256 * sizeof(xb200_scu_rec));
    if(coef) memcpy(coef, c->pic[h].coef, (size_t)c->n_lcu * 6144 * sizeof(int16_t));
    if(stat) { memset(stat, 0, sizeof(*stat)); stat->n_inter = c->pic[h].n_inter; stat->n_intra = c->pic[h].n_intra; }
    return XB200_OK;
}
/* XO_STANDIN_POLLS=n (test knob): a picture reports "not ready" to its first n polls, so the drop-in's non-blocking xeve_encode path
 * (XEVE_OK_OUT_NOT_AVAILABLE while the device works) is exercised although this engine decides synchronously */
static int si_picture_ready(xb200_ctx *c, int32_t h)
{
    if(h < 0 || h >= SI_MAXPIC || !c->pic[h].used || !c->pic[h].scu) return XB200_ERR_INVALID_ARGUMENT;
    const char *e = getenv("XO_STANDIN_POLLS");
    const int   n = e ? atoi(e) : 0;
    if(c->pic[h].polls < n) { c->pic[h].polls++; return 0; }
    return 1;
}
static const xb200_engine g_standin = {si_create, si_destroy, si_pic_create, si_pic_destroy, si_pic_upload, si_pic_download, si_analyze_picture,
                                       si_picture_fetch, si_picture_ready};
XO_API const xb200_engine *xo_engine(void) { return &g_standin; }

/* ---- driver of the public API (inc/xeve.h only): frames in memory -> bitstream, like the reference app's main loop ---------------- */
static int img_addref(XEVE_IMGB *i) { return ++i->refcnt; }
static int img_getref(XEVE_IMGB *i) { return i->refcnt; }
static int img_release(XEVE_IMGB *i) { return --i->refcnt; }
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
/* returns seconds inside xeve_encode (app/xeve_app.c:1236-1246), < 0 on error; stats (optional) = xeve_b200_get_stats before delete */
XO_API double xo_api_encode_clip(const void *yuv, int nframes, int w, int h, int in_depth, int preset, int qp, int threads, int bframes,
                                 const char *extra, uint8_t *bs_out, int64_t bs_cap, int64_t *bs_len, xeve_b200_stats *stats)
{
    XEVE_CDSC cdsc;
    memset(&cdsc, 0, sizeof(cdsc));
    XEVE_PARAM *p = &cdsc.param;
    xeve_param_default(p);
    xeve_param_ppt(p, XEVE_PROFILE_BASELINE, preset, XEVE_TUNE_NONE);
    p->w = w; p->h = h; p->fps.num = 30; p->fps.den = 1;
    p->threads = threads;
    if(qp >= 0) p->qp = qp;
    if(bframes >= 0) p->bframes = bframes;
    p->cs = XEVE_CS_SET(XEVE_CF_YCBCR420, p->codec_bit_depth, 0);
    if(extra && *extra) {
        char *dup = strdup(extra), *save = NULL;
        for(char *tok = strtok_r(dup, ";", &save); tok; tok = strtok_r(NULL, ";", &save)) {
            char *eq = strchr(tok, '=');
            if(!eq) continue;
            *eq = 0;
            if(xeve_param_parse(p, tok, eq + 1) != XEVE_OK) fprintf(stderr, "xo_api_encode_clip: bad param %s\n", tok);
        }
        free(dup);
    }
    cdsc.max_bs_buf_size = 16 * 1024 * 1024;
    if(xeve_param_check(p) != XEVE_OK) return -2.0;
    int  err = 0;
    XEVE id = xeve_create(&cdsc, &err);
    if(!id) return -1.0;
    const int bps = in_depth > 8 ? 2 : 1;
    const size_t fsz = (size_t)w * h * 3 / 2 * bps;
    XEVE_IMGB img;
    uint8_t  *bs = malloc(16 * 1024 * 1024);
    XEVE_BITB bitb;
    XEVE_STAT stat;
    memset(&bitb, 0, sizeof(bitb));
    bitb.addr = bs; bitb.bsize = 16 * 1024 * 1024;
    double  t_enc = 0;
    int64_t total = 0;
    int     pushed = 0, bumping = 0, ret;
    for(;;) {
        if(!bumping) {
            if(pushed < nframes) {
                const uint8_t *f = (const uint8_t *)yuv + fsz * pushed;
                memset(&img, 0, sizeof(img));
                img.cs = XEVE_CS_SET(XEVE_CF_YCBCR420, in_depth, 0);
                img.np = 3;
                for(int c = 0; c < 3; c++) {
                    const int cw = c ? w / 2 : w, ch = c ? h / 2 : h;
                    img.w[c] = img.aw[c] = cw; img.h[c] = img.ah[c] = ch; img.s[c] = cw * bps; img.e[c] = ch;
                }
                img.a[0] = (void *)f; img.a[1] = (void *)(f + (size_t)w * h * bps); img.a[2] = (void *)(f + (size_t)w * h * bps * 5 / 4);
                img.addref = img_addref; img.getref = img_getref; img.release = img_release; img.refcnt = 1;
                img.ts[XEVE_TS_PTS] = pushed;
                ret = xeve_push(id, &img);
                if(XEVE_FAILED(ret)) { t_enc = -3; break; }
                pushed++;
            }
            else {
                int val = 1, size = sizeof(int);
                xeve_config(id, XEVE_CFG_SET_FORCE_OUT, &val, &size);
                bumping = 1;
            }
        }
        const double t0 = now_s();
        ret = xeve_encode(id, &bitb, &stat);
        t_enc += now_s() - t0;
        if(XEVE_FAILED(ret)) { t_enc = -4; break; }
        if(ret == XEVE_OK_NO_MORE_FRM) break;
        if(ret == XEVE_OK && stat.write > 0) {
            if(bs_out && total + stat.write <= bs_cap) memcpy(bs_out + total, bs, stat.write);
            total += stat.write;
        }
    }
    if(bs_len) *bs_len = total;
    if(stats) xeve_b200_get_stats(id, stats);
    free(bs);
    xeve_delete(id);
    return t_enc;
}
XO_API void xo_api_use_standin(int on) { xeve_b200_set_engine(on ? &g_standin : NULL); }
