/* xeve_b200_dropin.c -- the reference-side binding of the device decision pass: the xeve_create / xeve_push / xeve_encode /
 * xeve_config / xeve_param_* C API of inc/xeve.h:600-680 with the reference's own host code around libxeve_b200.so.
 *
 * north_star: "keeping the xeve_create/xeve_encode/xeve_push/xeve_pull C API and XEVE_CDSC config surface so it is a drop-in for that
 * path: host code stays C calling CUDA through a thin C-ABI ... entropy coding (xeve_eco) and the bitstream writer stay on the host".
 * This file is that host code's hook file.  It is compiled against the reference's headers where they lie (never copied) and linked
 * with the reference's own objects by integration/Makefile; the only change to the reference build is that src_base/xeve.c is
 * compiled with -Dxeve_create=xeve_create_host, so that the xeve_create exported here can install hooks on the context the
 * reference's xeve_create returns.  A maintainer of the reference would instead add the few lines of `install()` below to
 * xeve_platform_init (src_base/xeve_enc.c:757-833), see INTEGRATION.md.
 *
 * What the hooks do (all are function pointers the reference already calls, src_base/xeve_type.h:931-963):
 *   ctx->fn_push               + the pushed picture goes to the device (xb200_pic_upload) and a SHADOW context -- the reference's own
 *                                control plane run dry on a 64x64 picture -- yields the picture-level plan (slice type, POC, QPs,
 *                                lambdas, reference lists: with constant QP none depends on a decision) of every picture that can be
 *                                coded from the frames pushed so far; those pictures are enqueued at once (xb200_analyze_picture),
 *                                so the pictures of a GOP run concurrently on the device (picture DAG);
 *   ctx->fn_mode_analyze_frame   blocks until the picture the reference is about to code has been decided, fetches its records;
 *   ctx->fn_mode_analyze_lcu     copies the CTU's records to where mode_analyze_lcu leaves its own results (core->cu_data_best ->
 *                                update_to_ctx_map -> ctx->map_cu_data[lcu], src_base/xeve_mode.c:2521-2608): no decision on the host;
 *   ctx->fn_loop_filter          nothing (the device filtered and border-expanded the picture, it stays there as a reference), or a
 *                                download of the deblocked picture when the caller wants the reconstruction;
 *   ctx->fn_enc                  at the end of the stream (XEVE_CFG_SET_FORCE_OUT) the tail of the plan is recomputed exactly (the
 *                                reference restructures the last, partial GOP) and pictures enqueued on a guess are redone;
 *   ctx->fn_flush                releases the device context.
 * The reference then entropy-codes the records with its own xeve_eco_tree and writes the bitstream: byte-identical to the reference
 * run with the same `threads` (the decision pass runs `threads` coder-state chains like the reference's worker threads).
 *
 * Configurations outside the path (rate control, AQ / cu-tree / look-ahead, P slices, tiles, rdo-deblk-switch, me-algo > 1) keep
 * the reference's host code, with a notice on stderr.  Without an sm_100 device xeve_create FAILS (no CPU fallback). */
#define _GNU_SOURCE
#include "xeve_type.h"
#include "xeve_b200_engine.h"
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

XEVE xeve_create_host(XEVE_CDSC *cdsc, int *err);   /* the reference's xeve_create (src_base/xeve.c:35) under its build-time alias */

#define DI_LEAD   16     /* frames the shadow context runs ahead of the pushes (one default GOP) */
#define DI_MAXPIC 256    /* device pictures per pool */
#define DI_MAX_LAG 32    /* frames the encoder may fall behind its normal pace while the device works (the reference's input ring holds
                            XEVE_MAX_INBUF_CNT = 70 frames, its normal delay is bframes + 1 = 16) */

/* ---- engine table ------------------------------------------------------------------------------------------------------------------ */
static const xb200_engine g_cuda_engine = {xb200_create, xb200_destroy, xb200_pic_create, xb200_pic_destroy, xb200_pic_upload,
                                           xb200_pic_download, xb200_analyze_picture, xb200_picture_fetch, xb200_picture_ready};
static const xb200_engine *g_engine = &g_cuda_engine;
XB200_API void xeve_b200_set_engine(const xb200_engine *e) { g_engine = e ? e : &g_cuda_engine; }

/* ---- per-encoder state (ctx->pf) ---------------------------------------------------------------------------------------------------- */
typedef struct {
    xb200_picture pp;       /* ref_pic / cur_pic / rec_pic are resolved when the picture is enqueued */
    int           input;    /* index of the input frame (XEVE_PICO.pic_icnt) */
    int           rec, org; /* device handles once enqueued */
    int           state;    /* 0 planned, 1 enqueued, 2 fetched */
} Plan;
typedef struct { int poc, handle, idx; XEVE_PIC *rpic; } RecEnt;
typedef struct { int input, handle; } OrgEnt;

typedef struct DropIn {
    XEVE_CTX           *real, *shadow;
    XEVE_CDSC           cdsc;
    const xb200_engine *E;
    xb200_ctx          *dev;
    int  (*real_push)(XEVE_CTX *, XEVE_IMGB *);
    void (*real_flush)(XEVE_CTX *);
    int  (*real_enc)(XEVE_CTX *, XEVE_BITB *, XEVE_STAT *);
    int  (*real_frame)(XEVE_CTX *);
    int  (*real_header)(XEVE_CTX *);
    int     label_threads;   /* the caller's `threads`: coder-state chains per picture on the device, and what the parameter SEI says */
    Plan   *plan;
    int     n_plan, cap_plan;
    int     next_enq, next_fetch, pushed, shadow_pushed;
    int     tail_done, sync_mode, failed, want_recon, lag_ok, catchup;
    int     n_lcu, parallel_rows;
    RecEnt  rec[DI_MAXPIC];
    int     n_rec;
    OrgEnt  org[DI_MAXPIC];
    int     n_org;
    int     free_rec[DI_MAXPIC], n_free_rec, free_org[DI_MAXPIC], n_free_org;
    xb200_scu_rec *scu;     /* records of the picture being coded */
    int16_t       *coef;
    int            cur_rec; /* its device picture */
    xeve_b200_stats stats;
    XEVE_MTIME    *ts;      /* presentation time stamps of the frames pushed so far (closed GOPs: the reference's slice-type decision reads them) */
    int            n_ts, cap_ts;
    int            need_sync;   /* sync mode: the picture is planned, enqueued and fetched at its first CTU */
    uint8_t       *sh_bs;   /* bitstream buffer of the shadow context */
    void          *dummy_buf;
    struct Shadow *s1;      /* sink of the incremental shadow context */
} DropIn;

static double now_ms(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return 1e3 * ts.tv_sec + 1e-6 * ts.tv_nsec;
}
#define DI_FAIL(d, ...) do { fprintf(stderr, "xeve_b200 drop-in: " __VA_ARGS__); fputc('\n', stderr); (d)->failed = 1; } while(0)

/* ---- the shadow context: the reference's control plane run dry ------------------------------------------------------------------------ */
typedef struct Shadow {       /* what the dry-run hooks of a shadow context append to */
    DropIn        *owner;
    Plan          *out;       /* NULL: owner->plan (grown on demand) */
    int           *n_out, cap_out;
} Shadow;
static xb200_scu_rec g_dummy_scu[2][256];   /* [inter: 8x8 SKIP, zero motion | intra: 8x8 DC], no residual */
static int16_t       g_dummy_coef[6144];
static pthread_once_t g_dummy_once = PTHREAD_ONCE_INIT;
static void dummy_init(void)
{
    for(int k = 0; k < 2; k++)
        for(int i = 0; i < 256; i++) {
            xb200_scu_rec *u = &g_dummy_scu[k][i];
            memset(u, 0, sizeof(*u));
            u->log2 = 3;
            if(k) { u->mode = 3; u->refi[0] = u->refi[1] = -1; }
        }
}

/* picture-level inputs of xb200_analyze_picture from a context that is inside xeve_pic, at its first CTU (core = ctx->core[0]) */
static void plan_from_ctx(XEVE_CTX *ctx, XEVE_CORE *core, xb200_picture *p)
{
    XEVE_PINTER *pi = &ctx->pinter[core->thread_cnt];
    memset(p, 0, sizeof(*p));
    p->poc = (int)ctx->poc.poc_val; p->slice_type = ctx->slice_type;
    p->cur_pic = p->rec_pic = -1; p->unfiltered_pic = -1;
    p->tile_qp = ctx->tile[core->tile_idx].qp;
    for(int l = 0; l < 2; l++) {
        p->num_refp[l] = ctx->rpm.num_refp[l];
        for(int k = 0; k < XB200_MAX_REFP; k++) {
            p->ref_pic[l][k] = -1; p->ref_poc[l][k] = -1;
            if(ctx->slice_type != SLICE_I && k < p->num_refp[l] && (l == 0 || ctx->slice_type == SLICE_B))
                p->ref_poc[l][k] = (int)ctx->refp[k][l].poc;
        }
    }
    if(ctx->slice_type == SLICE_B && ctx->refp[0][REFP_1].list_poc) p->col_list_poc0 = (int)ctx->refp[0][REFP_1].list_poc[0];
    p->max_cu_inter = ctx->param.max_cu_inter; p->min_cu_inter = ctx->param.min_cu_inter;
    p->max_cu_intra = ctx->param.max_cu_intra; p->min_cu_intra = ctx->param.min_cu_intra;
    p->cip = ctx->pps.constrained_intra_pred_flag;
    {   /* mode_cu_init, src_base/xeve_mode.c:776-783, with core->qp == tile qp (no delta QP) */
        const int q = p->tile_qp, bdc = ctx->sps.bit_depth_chroma_minus8;
        p->qp[0] = GET_LUMA_QP(q, ctx->sps.bit_depth_luma_minus8);
        p->qp[1] = ctx->qp_chroma_dynamic[0][XEVE_CLIP3(-6 * bdc, 57, q + ctx->sh->qp_u_offset)] + 6 * bdc;
        p->qp[2] = ctx->qp_chroma_dynamic[1][XEVE_CLIP3(-6 * bdc, 57, q + ctx->sh->qp_v_offset)] + 6 * bdc;
    }
    p->lambda_mv = pi->lambda_mv; p->max_search_range = pi->max_search_range;
    for(int i = 0; i < 3; i++) p->lambda[i] = core->lambda[i];
    p->sqrt_lambda0 = core->sqrt_lambda[0];
    p->dist_chroma_weight[0] = core->dist_chroma_weight[0]; p->dist_chroma_weight[1] = core->dist_chroma_weight[1];
    p->deblock = ctx->sh->deblocking_filter_on;
    {   /* loop-filter inputs (xeve_deblock: src_base/xeve_df.c:522-573) */
        const int bdo = 6 * (ctx->param.codec_bit_depth - 8);
        p->df.qp_u_offset = ctx->sh->qp_u_offset; p->df.qp_v_offset = ctx->sh->qp_v_offset;
        for(int c = 0; c < 2; c++)
            for(int q = -bdo; q <= 57; q++) p->df.chroma_qp[c][q + bdo] = ctx->qp_chroma_dynamic[c][q];
    }
}
static int same_plan(const xb200_picture *a, const xb200_picture *b)
{
    xb200_picture x = *a, y = *b;   /* everything but the handles and the sizes the owner fills in */
    x.cur_pic = y.cur_pic = x.rec_pic = y.rec_pic = x.unfiltered_pic = y.unfiltered_pic = 0;
    x.parallel_rows = y.parallel_rows = 0;
    x.df.w_scu = y.df.w_scu = x.df.h_scu = y.df.h_scu = 0;
    memset(x.ref_pic, 0, sizeof(x.ref_pic)); memset(y.ref_pic, 0, sizeof(y.ref_pic));
    return memcmp(&x, &y, sizeof(x)) == 0;
}

/* the CTU's records -> core->cu_data_best[CTU] -> ctx->map_cu_data[lcu] (what mode_analyze_lcu leaves behind) */
static void inject_split(XEVE_CTX *ctx, XEVE_CU_DATA *cd, const xb200_scu_rec *scu, int x0, int y0, int x, int y, int log2, int cud, int cup)
{
    if(x >= ctx->w || y >= ctx->h) return;
    const int cuw = 1 << log2, leaf = scu[((y - y0) >> 2) * 16 + ((x - x0) >> 2)].log2 == log2 || log2 == 2;
    xeve_set_split_mode(leaf ? NO_SPLIT : SPLIT_QUAD, cud, cup, cuw, cuw, ctx->max_cuwh, cd->split_mode);
    if(leaf) return;
    XEVE_SPLIT_STRUCT ss;
    xeve_split_get_part_structure(SPLIT_QUAD, x, y, cuw, cuw, cup, cud, ctx->log2_culine, &ss);
    for(int i = 0; i < ss.part_count; i++)
        inject_split(ctx, cd, scu, x0, y0, ss.x_pos[i], ss.y_pos[i], ss.log_cuw[i], ss.cud[i], ss.cup[i]);
}
static int inject_lcu(XEVE_CTX *ctx, XEVE_CORE *core, const xb200_scu_rec *scu, const int16_t *coef)
{
    const int L = ctx->log2_max_cuwh - 2, x0 = core->x_pel, y0 = core->y_pel, q = ctx->tile[core->tile_idx].qp;
    const int bdc = ctx->sps.bit_depth_chroma_minus8;
    const int qp_y = GET_LUMA_QP(q, ctx->sps.bit_depth_luma_minus8);
    const int qp_u = ctx->qp_chroma_dynamic[0][XEVE_CLIP3(-6 * bdc, 57, q + ctx->sh->qp_u_offset)] + 6 * bdc;
    const int qp_v = ctx->qp_chroma_dynamic[1][XEVE_CLIP3(-6 * bdc, 57, q + ctx->sh->qp_v_offset)] + 6 * bdc;
    XEVE_CU_DATA *cd = &core->cu_data_best[L][L];
    init_cu_data(cd, ctx->log2_max_cuwh, ctx->log2_max_cuwh, ctx->qp, ctx->qp, ctx->qp);
    inject_split(ctx, cd, scu, x0, y0, x0, y0, ctx->log2_max_cuwh, 0, 0);
    static const u8 mode_of[4] = {MODE_SKIP, MODE_DIR, MODE_INTER, MODE_INTRA};
    for(int i = 0; i < 256; i++) {
        const xb200_scu_rec *r = &scu[i];
        if(x0 + (i & 15) * 4 >= ctx->w || y0 + (i >> 4) * 4 >= ctx->h) continue;
        const int mode = mode_of[r->mode & 3];
        cd->pred_mode[i] = cd->pred_mode_chroma[i] = mode;
        cd->skip_flag[i] = mode == MODE_SKIP; cd->mmvd_flag[i] = 0; cd->affine_flag[i] = 0; cd->ibc_flag[i] = 0;
        for(int c = 0; c < 3; c++) {
            cd->nnz[c][i] = r->nnz[c];
            for(int sb = 0; sb < MAX_SUB_TB_NUM; sb++) cd->nnz_sub[c][sb][i] = sb == 0 ? r->nnz[c] : 0;
        }
        cd->qp_y[i] = qp_y; cd->qp_u[i] = qp_u; cd->qp_v[i] = qp_v;
        cd->map_scu[i] = 0;
        MCU_SET_IF_COD_SN_QP(cd->map_scu[i], mode == MODE_INTRA, ctx->slice_num, q);
        if(mode == MODE_SKIP) MCU_SET_SF(cd->map_scu[i]);
        cd->depth[i] = 2 * (ctx->log2_max_cuwh - r->log2);
        cd->map_cu_mode[i] = 0;
        MCU_SET_LOGW(cd->map_cu_mode[i], r->log2); MCU_SET_LOGH(cd->map_cu_mode[i], r->log2);
        cd->ipm[0][i] = cd->ipm[1][i] = mode == MODE_INTRA ? r->ipm : 0;
        for(int l = 0; l < 2; l++) {
            cd->refi[i][l] = r->refi[l]; cd->mvp_idx[i][l] = r->mvp_idx[l];
            cd->mv[i][l][0] = r->mv[l][0]; cd->mv[i][l][1] = r->mv[l][1]; cd->mvd[i][l][0] = r->mvd[l][0]; cd->mvd[i][l][1] = r->mvd[l][1];
        }
        cd->mvr_idx[i] = 0; cd->bi_idx[i] = 0; cd->mmvd_idx[i] = 0; cd->dmvr_flag[i] = 0;
    }
    memcpy(cd->coef[Y_C], coef, 4096 * 2); memcpy(cd->coef[U_C], coef + 4096, 1024 * 2); memcpy(cd->coef[V_C], coef + 5120, 1024 * 2);
    update_to_ctx_map(ctx, core);
    copy_cu_data(&ctx->map_cu_data[core->lcu_num], cd, 0, 0, ctx->log2_max_cuwh, ctx->log2_max_cuwh, ctx->log2_max_cuwh, 0,
                 xeve_get_default_tree_cons(), ctx->sps.chroma_format_idc);
    const int xs = x0 >> 2, ys = y0 >> 2, w = XEVE_MIN(16, ctx->w_scu - xs), h = XEVE_MIN(16, ctx->h_scu - ys);
    for(int j = 0; j < h; j++)
        for(int i = 0; i < w; i++) MCU_CLR_COD(ctx->map_scu[(size_t)(ys + j) * ctx->w_scu + xs + i]);
    return XEVE_OK;
}

static int shadow_lcu(XEVE_CTX *ctx, XEVE_CORE *core)
{
    Shadow *s = (Shadow *)ctx->pf;
    if(core->lcu_num == 0) {
        DropIn *d = s->owner;
        Plan   *p;
        if(s->out) {
            if(*s->n_out >= s->cap_out) return XEVE_ERR;
            p = &s->out[(*s->n_out)++];
        }
        else {
            if(d->n_plan == d->cap_plan) {
                d->cap_plan = d->cap_plan ? 2 * d->cap_plan : 64;
                d->plan = (Plan *)realloc(d->plan, (size_t)d->cap_plan * sizeof(Plan));
            }
            p = &d->plan[d->n_plan++];
        }
        memset(p, 0, sizeof(*p));
        plan_from_ctx(ctx, core, &p->pp);
        p->pp.parallel_rows = d->parallel_rows;
        p->input = (int)ctx->pico->pic_icnt;
        p->rec = p->org = -1;
    }
    return inject_lcu(ctx, core, g_dummy_scu[ctx->slice_type == SLICE_I], g_dummy_coef);
}
static int shadow_loop_filter(XEVE_CTX *ctx, XEVE_CORE *core) { (void)ctx; (void)core; return XEVE_OK; }

static int img_addref(XEVE_IMGB *i) { return ++i->refcnt; }
static int img_getref(XEVE_IMGB *i) { return i->refcnt; }
static int img_release(XEVE_IMGB *i) { return --i->refcnt; }

static XEVE_CTX *shadow_create(DropIn *d, Shadow *s)
{
    XEVE_CDSC c = d->cdsc;
    int       err = 0;
    c.param.w = c.param.h = 64;
    c.param.threads = 1;
    c.max_bs_buf_size = 1 << 20;
    XEVE_CTX *sh = (XEVE_CTX *)xeve_create_host(&c, &err);
    if(!sh) return NULL;
    sh->pf = s;
    sh->fn_mode_analyze_lcu = shadow_lcu;
    sh->fn_loop_filter = shadow_loop_filter;
    return sh;
}
static void shadow_delete(XEVE_CTX *sh)
{
    if(!sh) return;
    sh->pf = NULL;
    xeve_delete((XEVE)sh);
}
/* one dummy frame into a shadow context, then every picture it can code */
/* time stamp of input frame idx: the application's for the frames it has pushed, extrapolated for the ones the shadow runs ahead by */
static XEVE_MTIME ts_of(const DropIn *d, int idx)
{
    if(idx < d->n_ts) return d->ts[idx];
    if(d->n_ts >= 2) return d->ts[d->n_ts - 1] + (XEVE_MTIME)(idx - (d->n_ts - 1)) * (d->ts[d->n_ts - 1] - d->ts[d->n_ts - 2]);
    return (d->n_ts == 1 ? d->ts[0] : 0) + idx;
}
/* push = -1: encode only; otherwise the index of the dummy frame to push first */
static int shadow_step(DropIn *d, XEVE_CTX *sh, int push)
{
    XEVE_BITB bitb;
    XEVE_STAT stat;
    if(push >= 0) {
        XEVE_IMGB img;
        const int depth = d->cdsc.param.codec_bit_depth, bps = depth > 8 ? 2 : 1;
        memset(&img, 0, sizeof(img));
        img.cs = XEVE_CS_SET(XEVE_CF_YCBCR420, depth, 0);
        img.np = 3;
        for(int c = 0; c < 3; c++) {
            const int cw = c ? 32 : 64;
            img.w[c] = img.aw[c] = cw; img.h[c] = img.ah[c] = cw; img.s[c] = cw * bps; img.e[c] = cw;
            img.a[c] = d->dummy_buf;
        }
        img.addref = img_addref; img.getref = img_getref; img.release = img_release; img.refcnt = 1;
        img.ts[XEVE_TS_PTS] = ts_of(d, push);
        if(XEVE_FAILED(xeve_push((XEVE)sh, &img))) return XEVE_ERR;
    }
    memset(&bitb, 0, sizeof(bitb));
    bitb.addr = d->sh_bs; bitb.bsize = 1 << 20;
    const int ret = xeve_encode((XEVE)sh, &bitb, &stat);
    return ret;
}

/* ---- device pictures ---------------------------------------------------------------------------------------------------------------- */
static int org_of(DropIn *d, int input)
{
    for(int i = 0; i < d->n_org; i++) if(d->org[i].input == input) return d->org[i].handle;
    return -1;
}
static void org_release(DropIn *d, int input)
{
    for(int i = 0; i < d->n_org; i++)
        if(d->org[i].input == input) {
            d->free_org[d->n_free_org++] = d->org[i].handle;
            d->org[i] = d->org[--d->n_org];
            return;
        }
}
static int rec_of_poc(DropIn *d, int poc)   /* the latest picture with that POC */
{
    int best = -1, idx = -1;
    for(int i = 0; i < d->n_rec; i++) if(d->rec[i].poc == poc && d->rec[i].idx > idx) { best = d->rec[i].handle; idx = d->rec[i].idx; }
    return best;
}
static void rec_drop(DropIn *d, int i)
{
    d->free_rec[d->n_free_rec++] = d->rec[i].handle;
    d->rec[i] = d->rec[--d->n_rec];
}
/* pictures the reference's DPB has released are not referenced by any picture later in coding order, and every earlier picture has
 * completed (the real context fetched it): their device pictures can be reused at once */
static void rec_collect(DropIn *d, int idx_now)
{
    for(int i = d->n_rec - 1; i >= 0; i--) {
        RecEnt *e = &d->rec[i];
        if(!e->rpic || e->idx >= idx_now) continue;
        int dead = !e->rpic->is_ref || (int)e->rpic->poc != e->poc;
        for(int j = 0; j < d->n_rec && !dead; j++) dead = j != i && d->rec[j].rpic == e->rpic && d->rec[j].idx > e->idx;
        if(dead) rec_drop(d, i);
    }
}

static int enqueue_ready(DropIn *d)
{
    while(d->next_enq < d->n_plan && d->plan[d->next_enq].input < d->pushed && !d->failed) {
        Plan *p = &d->plan[d->next_enq];
        xb200_picture *pp = &p->pp;
        if(d->n_rec >= DI_MAXPIC - 1) { DI_FAIL(d, "too many live device pictures"); return XEVE_ERR; }
        for(int l = 0; l < 2; l++)
            for(int k = 0; k < XB200_MAX_REFP; k++) {
                pp->ref_pic[l][k] = -1;
                if(pp->ref_poc[l][k] < 0 || pp->slice_type == SLICE_I) continue;
                if((pp->ref_pic[l][k] = rec_of_poc(d, pp->ref_poc[l][k])) < 0) {
                    DI_FAIL(d, "POC %d: reference POC %d is not on the device", pp->poc, pp->ref_poc[l][k]);
                    return XEVE_ERR;
                }
            }
        if((p->org = org_of(d, p->input)) < 0) { DI_FAIL(d, "POC %d: input frame %d is not on the device", pp->poc, p->input); return XEVE_ERR; }
        int32_t h = -1;
        if(d->n_free_rec) h = d->free_rec[--d->n_free_rec];
        else if(d->E->pic_create(d->dev, 1, &h) != XB200_OK) { DI_FAIL(d, "xb200_pic_create failed"); return XEVE_ERR; }
        p->rec = h;
        pp->cur_pic = p->org; pp->rec_pic = p->rec; pp->unfiltered_pic = -1;
        const int r = d->E->analyze_picture(d->dev, pp);
        if(r != XB200_OK) { DI_FAIL(d, "xb200_analyze_picture(POC %d) = %d", pp->poc, r); return XEVE_ERR; }
        RecEnt *e = &d->rec[d->n_rec++];
        e->poc = pp->poc; e->handle = p->rec; e->idx = d->next_enq; e->rpic = NULL;
        p->state = 1;
        d->next_enq++;
    }
    return XEVE_OK;
}
/* wait for an enqueued picture and throw its result away (its plan turned out to be a wrong guess) */
static void discard(DropIn *d, int idx)
{
    Plan *p = &d->plan[idx];
    if(p->state != 1) return;
    d->E->picture_fetch(d->dev, p->rec, NULL, NULL, NULL, NULL, NULL);
    for(int i = 0; i < d->n_rec; i++) if(d->rec[i].idx == idx) { rec_drop(d, i); break; }
    p->state = 0; p->rec = -1;
    d->stats.replanned++;
}

/* ---- hooks of the real context ---------------------------------------------------------------------------------------------------------- */
static int hook_push(XEVE_CTX *ctx, XEVE_IMGB *img)
{
    DropIn *d = (DropIn *)ctx->pf;
    const int ret = d->real_push(ctx, img);
    if(XEVE_FAILED(ret) || d->failed) return d->failed ? XEVE_ERR : ret;
    {   /* the picture as the reference stored it (xeve_imgb_cpy: internal bit depth, s16) -> device */
        XEVE_PIC *pic = PIC_ORIG(ctx);
        int32_t   h = -1;
        if(d->n_org >= DI_MAXPIC - 1) { DI_FAIL(d, "too many input pictures on the device"); return XEVE_ERR; }
        if(d->n_free_org) h = d->free_org[--d->n_free_org];
        else if(d->E->pic_create(d->dev, 0, &h) != XB200_OK) { DI_FAIL(d, "xb200_pic_create failed"); return XEVE_ERR; }
        const void *const planes[3] = {pic->y, pic->u, pic->v};
        const int32_t     stride[3] = {pic->s_l * 2, pic->s_c * 2, pic->s_c * 2};
        if(d->E->pic_upload(d->dev, h, planes, stride, ctx->param.codec_bit_depth, XB200_MEM_HOST) != XB200_OK) {
            DI_FAIL(d, "xb200_pic_upload failed");
            return XEVE_ERR;
        }
        d->org[d->n_org].input = (int)ctx->pic_icnt; d->org[d->n_org].handle = h; d->n_org++;
        d->pushed = (int)ctx->pic_icnt + 1;
        if(d->n_ts == d->cap_ts) {
            d->cap_ts = d->cap_ts ? 2 * d->cap_ts : 256;
            d->ts = (XEVE_MTIME *)realloc(d->ts, (size_t)d->cap_ts * sizeof(XEVE_MTIME));
        }
        d->ts[d->n_ts++] = img->ts[XEVE_TS_PTS];
    }
    if(!d->tail_done && !d->sync_mode && d->shadow) {   /* the shadow context stays DI_LEAD frames ahead */
        while(d->shadow_pushed < d->pushed + DI_LEAD) {
            if(XEVE_FAILED(shadow_step(d, d->shadow, d->shadow_pushed))) { DI_FAIL(d, "shadow context failed"); return XEVE_ERR; }
            d->shadow_pushed++;
        }
    }
    if(XEVE_FAILED(enqueue_ready(d))) return XEVE_ERR;
    return ret;
}

/* end of the stream: the reference codes the remaining pictures in "bumping" mode and restructures a partial last GOP.  A fresh
 * shadow context fed with exactly the N frames of the stream yields the exact plan; guesses that differ are redone. */
static int replan_tail(DropIn *d, int n_frames)
{
    Shadow    s2;
    Plan     *p2 = (Plan *)calloc((size_t)n_frames + 1, sizeof(Plan));
    int       n2 = 0, val = 1, size = sizeof(int), ret = XEVE_OK;
    s2.owner = d; s2.out = p2; s2.n_out = &n2; s2.cap_out = n_frames + 1;
    XEVE_CTX *sh = shadow_create(d, &s2);
    if(!sh) { free(p2); return XEVE_ERR; }
    for(int i = 0; i < n_frames && !XEVE_FAILED(ret); i++) ret = shadow_step(d, sh, i);
    xeve_config((XEVE)sh, XEVE_CFG_SET_FORCE_OUT, &val, &size);
    while(!XEVE_FAILED(ret) && ret != XEVE_OK_NO_MORE_FRM) ret = shadow_step(d, sh, -1);
    shadow_delete(sh);
    if(XEVE_FAILED(ret) || n2 != n_frames) { free(p2); DI_FAIL(d, "tail plan: %d pictures for %d frames (ret %d)", n2, n_frames, ret); return XEVE_ERR; }
    /* the first picture whose guess differs from the exact plan, and everything after it in coding order, is redone: wait for the
     * enqueued ones (latest first -- a later picture may read an earlier one), then enqueue the exact tail */
    int first_bad = d->next_fetch;
    while(first_bad < n_frames && first_bad < d->n_plan && same_plan(&d->plan[first_bad].pp, &p2[first_bad].pp) &&
          d->plan[first_bad].input == p2[first_bad].input)
        first_bad++;
    for(int j = d->n_plan - 1; j >= first_bad; j--) discard(d, j);
    if(d->next_enq > first_bad) d->next_enq = first_bad;
    if(d->cap_plan < n_frames) { d->cap_plan = n_frames; d->plan = (Plan *)realloc(d->plan, (size_t)d->cap_plan * sizeof(Plan)); }
    for(int idx = first_bad; idx < n_frames; idx++) d->plan[idx] = p2[idx];
    d->n_plan = n_frames;
    free(p2);
    return enqueue_ready(d);
}

static int hook_enc(XEVE_CTX *ctx, XEVE_BITB *bitb, XEVE_STAT *stat)
{
    DropIn *d = (DropIn *)ctx->pf;
    if(d->failed) return XEVE_ERR;
    if(FORCE_OUT(ctx) && !d->tail_done) {
        d->tail_done = 1;
        if(!d->sync_mode && XEVE_FAILED(replan_tail(d, (int)ctx->pic_ticnt + 1))) return XEVE_ERR;
        /* pictures the encoder would have coded before the end of the input had xeve_encode not answered "not available" while the
         * device worked: with N frames pushed it codes N - frm_rnum pictures in normal mode, the rest while bumping */
        d->catchup = (int)ctx->pic_ticnt + 1 - (int)ctx->frm_rnum - d->next_fetch;
        if(d->catchup < 0) d->catchup = 0;
    }
    if(FORCE_OUT(ctx) && d->catchup > 0) {
        /* The reference derives "pictures coded so far" from the frames pushed (xeve_enc: pic_icnt - frm_rnum, with a pseudo push per
         * call while bumping, src_base/xeve_enc.c:607-620, 969-972), i.e. it assumes the lock step this encoder left.  Code the
         * pictures it is behind by exactly as it would have coded them then: the pseudo push undone, force-output off for the call. */
        ctx->pic_icnt--;
        ctx->param.force_output = 0;
        const int ret = d->real_enc(ctx, bitb, stat);
        ctx->param.force_output = 1;
        d->catchup--;
        return ret;
    }
    /* Do not block the caller's pushes while the device is still deciding the next picture: xeve_encode may answer
     * XEVE_OK_OUT_NOT_AVAILABLE (inc/xeve.h:52, the answer it gives while its own delay fills), the application pushes the next frame
     * and the pictures that frame makes codable are enqueued -- so the next GOPs overlap the current one on the device.  Bounded by
     * the reference's input ring; at the end of the stream (FORCE_OUT: no more pushes) encode blocks. */
    if(!FORCE_OUT(ctx) && d->lag_ok && d->next_fetch < d->n_plan && d->plan[d->next_fetch].state == 1) {
        const int lag = d->pushed - (d->next_fetch + (int)ctx->frm_rnum + 1);
        if(lag < DI_MAX_LAG && d->E->picture_ready(d->dev, d->plan[d->next_fetch].rec) == 0) {
            d->stats.deferred++;
            return XEVE_OK_OUT_NOT_AVAILABLE;
        }
    }
    return d->real_enc(ctx, bitb, stat);
}

/* wait for picture idx (enqueued) and take its records */
static int fetch_picture(DropIn *d, XEVE_CTX *ctx, int idx)
{
    if(XEVE_FAILED(enqueue_ready(d)) || d->plan[idx].state != 1) {
        if(!d->failed) DI_FAIL(d, "picture %d (POC %d) could not be enqueued", idx, (int)ctx->poc.poc_val);
        return XEVE_ERR;
    }
    rec_collect(d, idx);
    Plan *p = &d->plan[idx];
    xb200_picture_stat st;
    memset(&st, 0, sizeof(st));
    const double t0 = now_ms();
    const int    r = d->E->picture_fetch(d->dev, p->rec, d->scu, d->coef, NULL, NULL, &st);
    d->stats.wait_ms += now_ms() - t0;
    if(r != XB200_OK) { DI_FAIL(d, "xb200_picture_fetch(POC %d) = %d", p->pp.poc, r); return XEVE_ERR; }
    p->state = 2;
    d->cur_rec = p->rec;
    for(int i = 0; i < d->n_rec; i++) if(d->rec[i].idx == idx) d->rec[i].rpic = PIC_CURR(ctx);
    org_release(d, p->input);
    d->next_fetch++;
    d->stats.pictures++; d->stats.n_inter += st.n_inter; d->stats.n_intra += st.n_intra;
    d->stats.chain_ms += st.chain_ms; d->stats.filter_ms += st.filter_ms;
    return XEVE_OK;
}
/* The plan ran ahead on a guess that the encoder's own state does not confirm (a parameter changed through xeve_config while
 * encoding, a GOP structure the shadow context could not foresee): everything enqueued from this picture on is thrown away, and from
 * now on every picture is planned from the real context at its first CTU, enqueued and waited for -- correct, without look-ahead. */
static void enter_sync_mode(DropIn *d, int idx, const char *why)
{
    if(!getenv("XB200_QUIET")) fprintf(stderr, "xeve_b200 drop-in: picture %d: %s -- pictures are planned one at a time from here on\n", idx, why);
    for(int j = d->n_plan - 1; j >= idx; j--) discard(d, j);
    d->n_plan = idx;
    if(d->next_enq > idx) d->next_enq = idx;
    d->sync_mode = 1;
}
/* the picture the reference is about to code: make sure it is (being) decided with the reference's own parameters, wait, fetch */
static int hook_frame(XEVE_CTX *ctx)
{
    DropIn *d = (DropIn *)ctx->pf;
    d->cur_rec = -1;
    d->need_sync = 0;
    if(d->failed) return XEVE_ERR;
    const int idx = d->next_fetch;
    if(!d->sync_mode) {
        int ok = idx < d->n_plan;
        if(ok) {   /* what is known before the CTU loop must agree with the plan; the rest is checked at the first CTU */
            const xb200_picture *pp = &d->plan[idx].pp;
            ok = pp->poc == (int)ctx->poc.poc_val && pp->slice_type == ctx->slice_type && pp->tile_qp == ctx->sh->qp &&
                 d->plan[idx].input == (int)ctx->pico->pic_icnt;
            for(int l = 0; l < 2 && ok; l++) {
                if(ctx->slice_type == SLICE_I || (l == 1 && ctx->slice_type != SLICE_B)) continue;
                ok = pp->num_refp[l] == ctx->rpm.num_refp[l];
                for(int k = 0; k < pp->num_refp[l] && k < XB200_MAX_REFP && ok; k++) ok = pp->ref_poc[l][k] == (int)ctx->refp[k][l].poc;
            }
        }
        {   /* test knob: pretend the plan of picture n was wrong */
            const char *e = getenv("XB200_DROPIN_FORCE_SYNC_AT");
            if(e && atoi(e) == idx) ok = 0;
        }
        if(!ok) enter_sync_mode(d, idx, "the picture plan does not match the encoder's state");
    }
    if(d->sync_mode) d->need_sync = 1;   /* planned at the first CTU, where the lambdas exist */
    else if(XEVE_FAILED(fetch_picture(d, ctx, idx))) return XEVE_ERR;
    return d->real_frame ? d->real_frame(ctx) : XEVE_OK;
}
static int hook_lcu(XEVE_CTX *ctx, XEVE_CORE *core)
{
    DropIn *d = (DropIn *)ctx->pf;
    if(d->failed) return XEVE_ERR;
    if(core->lcu_num == 0 && d->need_sync) {   /* sync mode: plan from the encoder's own state, enqueue, wait */
        const int idx = d->next_fetch;
        d->need_sync = 0;
        if(idx >= d->cap_plan) {
            d->cap_plan = d->cap_plan ? 2 * d->cap_plan : 64;
            d->plan = (Plan *)realloc(d->plan, (size_t)d->cap_plan * sizeof(Plan));
        }
        Plan *p = &d->plan[idx];
        memset(p, 0, sizeof(*p));
        plan_from_ctx(ctx, core, &p->pp);
        p->pp.parallel_rows = d->parallel_rows;
        p->input = (int)ctx->pico->pic_icnt;
        p->rec = p->org = -1;
        d->n_plan = idx + 1;
        d->next_enq = idx;
        if(XEVE_FAILED(fetch_picture(d, ctx, idx))) return XEVE_ERR;
    }
    if(d->cur_rec < 0) return XEVE_ERR;
    if(core->lcu_num == 0) {   /* the lambdas exist now: the plan the device used must be the encoder's */
        xb200_picture cur;
        plan_from_ctx(ctx, core, &cur);
        if(!same_plan(&cur, &d->plan[d->next_fetch - 1].pp)) {
            DI_FAIL(d, "POC %d: picture parameters differ from the plan the device used", cur.poc);
            return XEVE_ERR;
        }
    }
    return inject_lcu(ctx, core, d->scu + (size_t)core->lcu_num * 256, d->coef + (size_t)core->lcu_num * 6144);
}
static int hook_loop_filter(XEVE_CTX *ctx, XEVE_CORE *core)
{
    DropIn *d = (DropIn *)ctx->pf;
    (void)core;
    if(d->failed || d->cur_rec < 0) return XEVE_ERR;
    if(d->want_recon) {   /* deblocked picture -> the reference's picture buffer (its own border expansion follows, xeve_pic_finish) */
        XEVE_PIC      *pic = PIC_CURR(ctx);
        int16_t *const planes[3] = {pic->y, pic->u, pic->v};
        const int32_t  stride[3] = {pic->s_l, pic->s_c, pic->s_c};
        if(d->E->pic_download(d->dev, d->cur_rec, 0, planes, stride) != XB200_OK) { DI_FAIL(d, "xb200_pic_download failed"); return XEVE_ERR; }
    }
    return XEVE_OK;
}
/* The reference writes its parameters, `threads` among them, into an SEI message of the first access unit (xeve_param2string via
 * ctx->fn_enc_header, src_base/xeve_enc.c:2533-2537).  The host context runs single-threaded here (it decides nothing; its worker
 * threads would only busy-wait next to the other streams), the decisions are made as `threads` coder-state chains on the device: the
 * stream is labelled like the reference run it reproduces. */
static int hook_header(XEVE_CTX *ctx)
{
    DropIn   *d = (DropIn *)ctx->pf;
    const int keep = ctx->param.threads;
    ctx->param.threads = d->label_threads;
    const int ret = d->real_header(ctx);
    ctx->param.threads = keep;
    return ret;
}
static void dropin_free(DropIn *d)
{
    if(!d) return;
    if(d->shadow) shadow_delete(d->shadow);
    if(d->dev) d->E->destroy(d->dev);
    free(d->plan); free(d->scu); free(d->coef); free(d->sh_bs); free(d->dummy_buf); free(d->s1); free(d->ts);
    free(d);
}
static void hook_flush(XEVE_CTX *ctx)
{
    DropIn *d = (DropIn *)ctx->pf;
    void (*org)(XEVE_CTX *) = d->real_flush;
    ctx->pf = NULL;               /* xeve_platform_deinit expects it unused (src_base/xeve_enc.c:840) */
    dropin_free(d);
    if(org) org(ctx);
}

XB200_API int xeve_b200_get_stats(void *id, xeve_b200_stats *out)
{
    XEVE_CTX *ctx = (XEVE_CTX *)id;
    if(!ctx || !out) return XB200_ERR_INVALID_ARGUMENT;
    memset(out, 0, sizeof(*out));
    if(ctx->pf && ctx->fn_push == hook_push) { *out = ((DropIn *)ctx->pf)->stats; out->device_path = 1; }
    return XB200_OK;
}

/* why a configuration is outside the device path (NULL: inside) */
static const char *outside(const XEVE_PARAM *p)
{
    if(p->rc_type != 0) return "rate control (rc-type != CQP)";
    if(p->aq_mode || p->cutree || p->use_fcst) return "aq-mode / cu-tree / look-ahead";
    if(p->inter_slice_type != 0) return "P slices (inter-slice-type)";
    if(p->rdo_dbk_switch) return "rdo-deblk-switch (presets slow / placebo)";
    if(p->me_algo > 1) return "me-algo > 1";
    if(p->tile_rows * p->tile_columns > 1) return "tiles";
    if(p->chroma_format_idc != 1) return "chroma format other than 4:2:0";
    if(p->cu_qp_delta_area > 0 && p->aq_mode) return "delta QP";
    if(p->ibc_flag) return "intra block copy";
    if(p->max_cu_inter > 64 || p->max_cu_intra > 64 || p->min_cu_inter < 8 || p->min_cu_intra < 4) return "CU size limits";
    if(p->use_pic_sign) return NULL;   /* needs the reconstruction: forces the download, still inside */
    return NULL;
}

/* the lines a maintainer would add to xeve_platform_init */
static int install(XEVE_CTX *ctx, const XEVE_CDSC *cdsc, int label_threads)
{
    pthread_once(&g_dummy_once, dummy_init);
    DropIn *d = (DropIn *)calloc(1, sizeof(DropIn));
    if(!d) return XEVE_ERR_OUT_OF_MEMORY;
    d->real = ctx; d->cdsc = *cdsc; d->E = g_engine;
    d->cur_rec = -1;
    d->label_threads = label_threads < 1 ? 1 : label_threads;
    {
        XEVE_PINTER *pi = &ctx->pinter[0];
        xb200_seq    sq;
        memset(&sq, 0, sizeof(sq));
        sq.w = ctx->w; sq.h = ctx->h; sq.bit_depth = ctx->param.codec_bit_depth;
        sq.me_level = pi->me_level; sq.hpel_cnt = pi->search_pattern_hpel_cnt; sq.qpel_cnt = pi->search_pattern_qpel_cnt;
        sq.me_complexity = pi->me_complexity;
        sq.min_clip[0] = pi->min_clip[0]; sq.min_clip[1] = pi->min_clip[1]; sq.max_clip[0] = pi->max_clip[0]; sq.max_clip[1] = pi->max_clip[1];
        sq.merge_num = ctx->param.merge_num; sq.me_range = ctx->param.me_range; sq.gop_size = ctx->param.gop_size;
        sq.rdoq = ctx->param.rdoq; sq.tool_iqt = ctx->param.tool_iqt;
        const char *dv = getenv("XB200_DEVICE");
        const int   r = d->E->create(&d->dev, dv ? atoi(dv) : 0, &sq);
        if(r != XB200_OK) {
            fprintf(stderr, "xeve_b200 drop-in: no device context (xb200_create = %d) -- this library has no CPU path\n", r);
            dropin_free(d);
            return XEVE_ERR_UNSUPPORTED;
        }
    }
    const int h_lcu = (ctx->h + ctx->max_cuwh - 1) >> ctx->log2_max_cuwh;
    d->n_lcu = (int)ctx->f_lcu;
    d->parallel_rows = d->label_threads > h_lcu ? h_lcu : d->label_threads;
    if(d->parallel_rows < 1) d->parallel_rows = 1;
    d->scu = (xb200_scu_rec *)malloc((size_t)d->n_lcu * 256 * sizeof(xb200_scu_rec));
    d->coef = (int16_t *)malloc((size_t)d->n_lcu * 6144 * sizeof(int16_t));
    d->sh_bs = (uint8_t *)malloc(1 << 20);
    d->dummy_buf = calloc(64 * 64, 2);
    d->lag_ok = !(getenv("XB200_DROPIN_SYNC") && atoi(getenv("XB200_DROPIN_SYNC")));   /* 1: xeve_encode always blocks on the device */
    d->want_recon = ctx->param.use_pic_sign || (getenv("XB200_DROPIN_RECON") && atoi(getenv("XB200_DROPIN_RECON")));
    d->s1 = (Shadow *)calloc(1, sizeof(Shadow));
    if(d->s1) { d->s1->owner = d; d->shadow = shadow_create(d, d->s1); }
    if(!d->scu || !d->coef || !d->sh_bs || !d->dummy_buf || !d->shadow) { dropin_free(d); return XEVE_ERR_OUT_OF_MEMORY; }
    ctx->pf = d;
    d->real_push = ctx->fn_push; ctx->fn_push = hook_push;
    d->real_flush = ctx->fn_flush; ctx->fn_flush = hook_flush;
    d->real_enc = ctx->fn_enc; ctx->fn_enc = hook_enc;
    d->real_frame = ctx->fn_mode_analyze_frame; ctx->fn_mode_analyze_frame = hook_frame;
    d->real_header = ctx->fn_enc_header; ctx->fn_enc_header = hook_header;
    ctx->fn_mode_analyze_lcu = hook_lcu;
    ctx->fn_loop_filter = hook_loop_filter;
    return XEVE_OK;
}

/* inc/xeve.h:600 */
XEVE xeve_create(XEVE_CDSC *cdsc, int *err)
{
    XEVE id = xeve_create_host(cdsc, err);
    if(!id) return NULL;
    const char *why = outside(&((XEVE_CTX *)id)->param);   /* judged on the parameters as the reference completed them */
    if(why) {
        if(!getenv("XB200_QUIET")) fprintf(stderr, "xeve_b200 drop-in: %s is outside the device path -- this encoder runs the reference's host code\n", why);
        return id;
    }
    /* inside the path the host context decides nothing: it runs single-threaded (see hook_header), `threads` goes to the device */
    const int threads = cdsc->param.threads;
    XEVE_CDSC host = *cdsc;
    if(threads > 1) {
        xeve_delete(id);
        host.param.threads = 1;
        id = xeve_create_host(&host, err);
        if(!id) return NULL;
    }
    const int r = install((XEVE_CTX *)id, &host, threads);
    if(r != XEVE_OK) {
        xeve_delete(id);
        if(err) *err = r;
        return NULL;
    }
    return id;
}
