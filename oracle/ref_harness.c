/*
 * ref_harness.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A thin driver around the UNMODIFIED reference encoder (mpeg5/xeve, Baseline profile),
 * compiled against the reference's own headers where they lie under /root/reference and
 * linked with oracle/_ref/libxeveb_ref.so (see oracle/Makefile.ref).  It does three things:
 *
 *   1. probes   -- call the reference's kernel tables (C / SSE / AVX2 variants) on caller
 *                  buffers: SAD, SSD, DIFF, SATD, luma/chroma MC, fwd/inv transform stages;
 *   2. tracing  -- run a real encode through the public API with logging wrappers installed on
 *                  the reference's own operator hooks (pi->fn_me, pi->fn_mc, ctx->fn_tq; see
 *                  reference src_base/xeve_type.h:448-453, 981-982) and record every call's
 *                  inputs (the "work list") plus outputs / output hashes;
 *   3. replay   -- run the reference's functions over a work list on N host threads: this is
 *                  both the expected-output generator for the parity tests and the CPU arm
 *                  (`bench.py --impl reference`, cpu_baseline.kind = "reference").
 *
 * Everything in this file is original harness code; the reference is used only through its
 * headers, exported symbols and function-pointer hooks.
 */
#define _GNU_SOURCE
#include "xeve_type.h"
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define RH_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * record layouts (mirrored as numpy dtypes in xeve_b200/refharness.py)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int32_t  poc;           /* POC of the picture being coded */
    int32_t  cur_pic;       /* index into the picture table (original picture) */
    int32_t  ref_pic;       /* index into the picture table (reference picture) */
    int32_t  ref_poc;
    int16_t  x, y;
    uint8_t  log2w, log2h, lidx, bi;
    int8_t   refi;
    uint8_t  num_refp;
    int16_t  mvp[2];
    int16_t  mv_in[2];
    uint32_t lambda_mv;
    int32_t  mot_bits_in[2];
    int32_t  max_search_range;
    int32_t  gop_size;
    int32_t  org_bi_off;    /* element offset into the s16 side buffer, -1 if bi == 0 */
    /* outputs */
    int16_t  mv_out[2];
    uint32_t cost;
    int32_t  mot_bits_out[2];
} RH_ME_REC; /* 72 bytes */

typedef struct {
    int32_t  poc;
    int32_t  ref_pic[2];    /* picture-table index per list, -1 if refi invalid */
    int32_t  ref_poc[2];
    int16_t  x, y, w, h;
    int8_t   refi[2];
    int16_t  mv[2][2];
    uint64_t out_hash;      /* FNV-1a over pred[0] Y,U,V as produced in situ */
} RH_MC_REC;

typedef struct {
    int32_t  poc;
    uint8_t  log2w, log2h, slice_type, is_intra;
    uint8_t  run_stats, qp[3];
    int32_t  rate_idx;      /* index into the rate-table array */
    int64_t  in_off;        /* element offset of the 3 input planes (Y: w*h, U,V: w*h/4 each) */
    double   lambda[3];
    int32_t  nnz[3];
    uint64_t out_hash;      /* FNV-1a over the three output planes */
} RH_TQ_REC;

typedef struct {
    int32_t cbf_all[2], cbf_luma[2], cbf_cb[2], cbf_cr[2];
    int32_t run[NUM_CTX_CC_RUN][2];
    int32_t level[NUM_CTX_CC_LEVEL][2];
    int32_t last[NUM_CTX_CC_LAST][2];
} RH_RATES;

typedef struct {
    int32_t poc, kind;      /* kind 0 = original, 1 = reconstructed reference (padded) */
    int32_t w_l, h_l, w_c, h_c, s_l, s_c, pad_l, pad_c;
    int64_t off_y, off_u, off_v; /* element offsets of the *buffer start* (incl. padding) */
} RH_PIC;

typedef struct {            /* planes handed in by the caller for replay */
    int16_t *y, *u, *v;     /* top-left of the active area */
    int32_t  s_l, s_c, w_l, h_l, poc;
} RH_PLANES;

typedef struct {
    int32_t w, h, bit_depth, me_level, hpel_cnt, qpel_cnt, me_complexity;
    int32_t min_clip[2], max_clip[2];
    int32_t merge_num, me_range, gop_size, rdoq, tool_iqt;
} RH_CONST;

/* One ctx->fn_pinter_analyze_cu call (layout == xb200_cu_item of include/xeve_b200.h). */
#define RH_MAXR 4
typedef struct {
    int32_t  poc, cur_pic;
    int16_t  x, y;
    uint8_t  log2_cuw, log2_cuh, slice_type, ctx_skip, ctx_pred_mode, all_preds;
    uint8_t  num_refp[2];
    uint8_t  qp[3], pad0_;
    int32_t  max_search_range;   /* pi->max_search_range */
    int32_t  ref_pic[2][RH_MAXR], ref_poc[2][RH_MAXR];   /* [lidx][refi] */
    uint32_t lambda_mv;
    int32_t  rate_idx, state_in, state_out;
    double   lambda[3], dist_chroma_weight[2];
    int16_t  mvp[2][4][2];       /* xeve_get_motion candidates per list */
    int8_t   refi_pred[2][4];
    int16_t  mv_dir[2][2];       /* xeve_get_mv_dir (B slices) */
    int64_t  out_off;            /* element offset of the coef / rec output slots */
    /* results */
    double   cost;
    uint8_t  best_idx, pad1_;    /* PRED_L0 0, L1 1, BI 2, SKIP 3, DIR 4 */
    int8_t   refi[2];
    uint8_t  mvp_idx[2];
    int16_t  mv[2][2], mvd[2][2];
    int32_t  nnz[3];
    uint64_t coef_hash, rec_hash;  /* FNV-1a over coef / rec Y,U,V (test bookkeeping) */
    int32_t  me_first, me_cnt;     /* this CU's slice of the ME trace (test bookkeeping) */
} RH_CU_REC;

/* ------------------------------------------------------------------------------------------
 * growable buffers
 * ---------------------------------------------------------------------------------------- */
typedef struct { uint32_t range; uint16_t m[68]; } RH_SBAC;
static void sbac_pack(const XEVE_SBAC *d, RH_SBAC *s);
static void sbac_unpack(const RH_SBAC *s, XEVE_SBAC *d);
typedef struct { void *p; size_t n, cap, esz; } vec_t;
static void *vec_push(vec_t *v, size_t cnt)
{
    if(v->n + cnt > v->cap) {
        size_t nc = v->cap ? v->cap * 2 : 1024;
        while(nc < v->n + cnt) nc *= 2;
        v->p   = realloc(v->p, nc * v->esz);
        v->cap = nc;
    }
    void *r = (char *)v->p + v->n * v->esz;
    v->n += cnt;
    return r;
}
static void vec_reset(vec_t *v, size_t esz) { v->n = 0; v->esz = esz; }

static uint64_t fnv1a(uint64_t h, const void *data, size_t bytes)
{
    const uint8_t *p = data;
    for(size_t i = 0; i < bytes; i++) { h ^= p[i]; h *= 0x100000001b3ULL; }
    return h;
}
#define FNV_INIT 0xcbf29ce484222325ULL

/* ------------------------------------------------------------------------------------------
 * trace state
 * ---------------------------------------------------------------------------------------- */
enum { RH_T_ME = 1, RH_T_MC = 2, RH_T_TQ = 4 };

/* One set of trace state per process (g_T: hooks may run in the reference's worker threads, which have no thread-local state) or,
 * for encodes started with RH_T_ISOLATED (threads = 1, so every hook runs in the calling thread), one per calling thread: several
 * such encodes can then run concurrently in one process (one stream per host thread). */
typedef struct {
    XEVE_CTX *ctx;
    int       mask, pic_lo, pic_hi; /* record while pic_lo <= ctx->pic_cnt <= pic_hi */
    u32 (*org_me)(XEVE_PINTER *, int, int, int, int, s8 *, int, s16 *, s16 *, int, int);
    void (*org_mc)(XEVE_CTX *, XEVE_CORE *, int, int, int, int, s8 *, s16 (*)[MV_D], XEVE_REFP (*)[REFP_NUM],
                   pel (*)[N_C][MAX_CU_DIM], int, int, s16 (*)[REFP_NUM][MV_D]);
    int (*org_tq)(XEVE_CTX *, XEVE_CORE *, s16 (*)[MAX_CU_DIM], int, int, int, int *, int, int);
    double (*org_cu)(XEVE_CTX *, XEVE_CORE *, int, int, int, int, XEVE_MODE *, s16 (*)[MAX_CU_DIM], pel **, int *);
    vec_t    cu, cu_sbac; /* cu_sbac: coder states named by RH_CU_REC::state_in / state_out */
    int (*org_lf)(XEVE_CTX *, XEVE_CORE *);
    void (*org_df_unit)(XEVE_CTX *, XEVE_PIC *, int, int, int, int, int, XEVE_CORE *, int);
    double (*org_intra)(XEVE_CTX *, XEVE_CORE *, int, int, int, int, XEVE_MODE *, s16 (*)[MAX_CU_DIM], pel **, int *);
    vec_t    intra;    /* RH_INTRA_REC; coder states go to cu_sbac, neighbour samples to samp */
    vec_t    df, df_cu, df_maps; /* deblocking: one record per picture, CU rectangles, frame maps (bytes) */
    int (*org_lcu)(XEVE_CTX *, XEVE_CORE *);
    int (*org_frame)(XEVE_CTX *);
    vec_t    lcu;      /* RH_LCU_REC: one per ctx->fn_mode_analyze_lcu call */
    int      df_collect;
    vec_t    me, mc, tq, rates, pics, samp, sbac; /* sbac[i]: coder state rates[i] was derived from */
    /* samp: s16 side buffer (pictures, org_bi, tq inputs) */
    RH_CONST cst;
    RH_RATES last_rates;
    int      have_rates;
} RH_STATE_T;
static RH_STATE_T           g_T;
static __thread RH_STATE_T *tl_T;
#define T (*(tl_T ? tl_T : &g_T))
#define RH_T_ISOLATED 4096

static int tracing(int kind)
{
    if(!(T.mask & kind) || !T.ctx) return 0;
    int c = (int)T.ctx->pic_cnt;
    return c >= T.pic_lo && c <= T.pic_hi;
}

static int find_or_add_pic(XEVE_PIC *pic, int poc, int kind)
{
    RH_PIC *tab = T.pics.p;
    for(size_t i = 0; i < T.pics.n; i++)
        if(tab[i].poc == poc && tab[i].kind == kind) return (int)i;
    RH_PIC r;
    memset(&r, 0, sizeof(r));
    r.poc = poc; r.kind = kind;
    r.w_l = pic->w_l; r.h_l = pic->h_l; r.w_c = pic->w_c; r.h_c = pic->h_c;
    r.s_l = pic->s_l; r.s_c = pic->s_c; r.pad_l = pic->pad_l; r.pad_c = pic->pad_c;
    size_t ny = (size_t)pic->s_l * (pic->h_l + 2 * pic->pad_l);
    size_t nc = (size_t)pic->s_c * (pic->h_c + 2 * pic->pad_c);
    r.off_y = (int64_t)T.samp.n; memcpy(vec_push(&T.samp, ny), pic->buf_y, ny * sizeof(s16));
    r.off_u = (int64_t)T.samp.n; memcpy(vec_push(&T.samp, nc), pic->buf_u, nc * sizeof(s16));
    r.off_v = (int64_t)T.samp.n; memcpy(vec_push(&T.samp, nc), pic->buf_v, nc * sizeof(s16));
    *(RH_PIC *)vec_push(&T.pics, 1) = r;
    return (int)T.pics.n - 1;
}

static u32 hook_me(XEVE_PINTER *pi, int x, int y, int log2_cuw, int log2_cuh, s8 *refi, int lidx, s16 mvp[MV_D],
                   s16 mv[MV_D], int bi, int bit_depth_luma)
{
    if(!tracing(RH_T_ME)) return T.org_me(pi, x, y, log2_cuw, log2_cuh, refi, lidx, mvp, mv, bi, bit_depth_luma);
    RH_ME_REC r;
    memset(&r, 0, sizeof(r));
    XEVE_PIC *rp = pi->refp[*refi][lidx].pic;
    r.poc = pi->poc;
    r.cur_pic = find_or_add_pic(pi->pic_o, pi->poc, 0);
    r.ref_pic = find_or_add_pic(rp, (int)rp->poc, 1);
    r.ref_poc = (int)pi->refp[*refi][lidx].poc;
    r.x = x; r.y = y; r.log2w = log2_cuw; r.log2h = log2_cuh; r.lidx = lidx; r.bi = bi; r.refi = *refi;
    r.num_refp = pi->num_refp;
    r.mvp[0] = mvp[0]; r.mvp[1] = mvp[1]; r.mv_in[0] = mv[0]; r.mv_in[1] = mv[1];
    r.lambda_mv = pi->lambda_mv;
    r.mot_bits_in[0] = pi->mot_bits[0]; r.mot_bits_in[1] = pi->mot_bits[1];
    r.max_search_range = pi->max_search_range; r.gop_size = pi->gop_size;
    r.org_bi_off = -1;
    if(bi) {
        size_t n = (size_t)1 << (log2_cuw + log2_cuh);
        r.org_bi_off = (int32_t)T.samp.n;
        memcpy(vec_push(&T.samp, n), pi->org_bi, n * sizeof(s16));
    }
    u32 cost = T.org_me(pi, x, y, log2_cuw, log2_cuh, refi, lidx, mvp, mv, bi, bit_depth_luma);
    r.mv_out[0] = mv[0]; r.mv_out[1] = mv[1]; r.cost = cost;
    r.mot_bits_out[0] = pi->mot_bits[0]; r.mot_bits_out[1] = pi->mot_bits[1];
    *(RH_ME_REC *)vec_push(&T.me, 1) = r;
    return cost;
}

static void hook_mc(XEVE_CTX *ctx, XEVE_CORE *core, int x, int y, int w, int h, s8 refi[REFP_NUM], s16 (*mv)[MV_D],
                    XEVE_REFP (*refp)[REFP_NUM], pel pred[REFP_NUM][N_C][MAX_CU_DIM], int a, int b,
                    s16 (*c)[REFP_NUM][MV_D])
{
    T.org_mc(ctx, core, x, y, w, h, refi, mv, refp, pred, a, b, c);
    if(!tracing(RH_T_MC)) return;
    RH_MC_REC r;
    memset(&r, 0, sizeof(r));
    r.poc = ctx->poc.poc_val; r.x = x; r.y = y; r.w = w; r.h = h;
    for(int l = 0; l < 2; l++) {
        r.refi[l] = refi[l]; r.mv[l][0] = mv[l][0]; r.mv[l][1] = mv[l][1];
        r.ref_pic[l] = -1; r.ref_poc[l] = -1;
        if(REFI_IS_VALID(refi[l])) {
            XEVE_PIC *rp = refp[refi[l]][l].pic;
            r.ref_pic[l] = find_or_add_pic(rp, (int)rp->poc, 1);
            r.ref_poc[l] = (int)rp->poc;
        }
    }
    uint64_t hsh = FNV_INIT;
    hsh = fnv1a(hsh, pred[0][Y_C], (size_t)w * h * 2);
    hsh = fnv1a(hsh, pred[0][U_C], (size_t)w * h / 2);
    hsh = fnv1a(hsh, pred[0][V_C], (size_t)w * h / 2);
    r.out_hash = hsh;
    *(RH_MC_REC *)vec_push(&T.mc, 1) = r;
}

static void grab_rates(XEVE_CORE *core, RH_RATES *o)
{
    memcpy(o->cbf_all, core->rdoq_est_cbf_all, sizeof(o->cbf_all));
    memcpy(o->cbf_luma, core->rdoq_est_cbf_luma, sizeof(o->cbf_luma));
    memcpy(o->cbf_cb, core->rdoq_est_cbf_cb, sizeof(o->cbf_cb));
    memcpy(o->cbf_cr, core->rdoq_est_cbf_cr, sizeof(o->cbf_cr));
    memcpy(o->run, core->rdoq_est_run, sizeof(o->run));
    memcpy(o->level, core->rdoq_est_level, sizeof(o->level));
    memcpy(o->last, core->rdoq_est_last, sizeof(o->last));
}
static void put_rates(XEVE_CORE *core, const RH_RATES *o)
{
    memcpy(core->rdoq_est_cbf_all, o->cbf_all, sizeof(o->cbf_all));
    memcpy(core->rdoq_est_cbf_luma, o->cbf_luma, sizeof(o->cbf_luma));
    memcpy(core->rdoq_est_cbf_cb, o->cbf_cb, sizeof(o->cbf_cb));
    memcpy(core->rdoq_est_cbf_cr, o->cbf_cr, sizeof(o->cbf_cr));
    memcpy(core->rdoq_est_run, o->run, sizeof(o->run));
    memcpy(core->rdoq_est_level, o->level, sizeof(o->level));
    memcpy(core->rdoq_est_last, o->last, sizeof(o->last));
}

static int rate_index(XEVE_CORE *core, int log2_cuw, int log2_cuh)
{
    RH_RATES cur;
    grab_rates(core, &cur);
    if(!T.have_rates || memcmp(&cur, &T.last_rates, sizeof(cur))) {
        *(RH_RATES *)vec_push(&T.rates, 1) = cur;
        sbac_pack(&core->s_curr_best[log2_cuw - 2][log2_cuh - 2], (RH_SBAC *)vec_push(&T.sbac, 1));
        T.last_rates = cur;
        T.have_rates = 1;
    }
    return (int)T.rates.n - 1;
}

static int hook_tq(XEVE_CTX *ctx, XEVE_CORE *core, s16 coef[N_C][MAX_CU_DIM], int log2_cuw, int log2_cuh,
                   int slice_type, int nnz[N_C], int is_intra, int run_stats)
{
    if(!tracing(RH_T_TQ)) return T.org_tq(ctx, core, coef, log2_cuw, log2_cuh, slice_type, nnz, is_intra, run_stats);
    RH_TQ_REC r;
    memset(&r, 0, sizeof(r));
    size_t ny = (size_t)1 << (log2_cuw + log2_cuh), nc = ny >> 2;
    r.poc = ctx->poc.poc_val; r.log2w = log2_cuw; r.log2h = log2_cuh; r.slice_type = slice_type;
    r.is_intra = is_intra; r.run_stats = run_stats;
    r.qp[0] = core->qp_y; r.qp[1] = core->qp_u; r.qp[2] = core->qp_v;
    for(int i = 0; i < 3; i++) r.lambda[i] = core->lambda[i];
    r.rate_idx = rate_index(core, log2_cuw, log2_cuh);
    r.in_off = (int64_t)T.samp.n;
    s16 *dst = vec_push(&T.samp, ny + 2 * nc);
    memcpy(dst, coef[Y_C], ny * 2);
    memcpy(dst + ny, coef[U_C], nc * 2);
    memcpy(dst + ny + nc, coef[V_C], nc * 2);
    int ret = T.org_tq(ctx, core, coef, log2_cuw, log2_cuh, slice_type, nnz, is_intra, run_stats);
    for(int i = 0; i < 3; i++) r.nnz[i] = nnz[i];
    uint64_t hsh = FNV_INIT;
    hsh = fnv1a(hsh, coef[Y_C], ny * 2);
    hsh = fnv1a(hsh, coef[U_C], nc * 2);
    hsh = fnv1a(hsh, coef[V_C], nc * 2);
    r.out_hash = hsh;
    *(RH_TQ_REC *)vec_push(&T.tq, 1) = r;
    return ret;
}

#define RH_T_CU 8
#define RH_T_CU_TIME 16 /* no records: only the time spent inside the reference's xeve_pinter_analyze_cu and the call count */
static double g_cu_secs;
static int64_t g_cu_calls;
RH_API double rh_cu_time(int64_t *calls) { if(calls) *calls = g_cu_calls; return g_cu_secs; }
static double now_s(void);
static double hook_cu(XEVE_CTX *ctx, XEVE_CORE *core, int x, int y, int log2_cuw, int log2_cuh, XEVE_MODE *mi,
                      s16 coef[N_C][MAX_CU_DIM], pel *rec[N_C], int s_rec[N_C])
{
    if(tracing(RH_T_CU_TIME) && !tracing(RH_T_CU)) {
        const double t0 = now_s();
        const double c = T.org_cu(ctx, core, x, y, log2_cuw, log2_cuh, mi, coef, rec, s_rec);
        g_cu_secs += now_s() - t0;
        g_cu_calls++;
        return c;
    }
    if(!tracing(RH_T_CU)) return T.org_cu(ctx, core, x, y, log2_cuw, log2_cuh, mi, coef, rec, s_rec);
    XEVE_PINTER *pi = &ctx->pinter[core->thread_cnt];
    RH_CU_REC    r;
    memset(&r, 0, sizeof(r));
    r.poc = ctx->poc.poc_val;
    r.cur_pic = find_or_add_pic(pi->pic_o, r.poc, 0);
    r.x = x; r.y = y; r.log2_cuw = log2_cuw; r.log2_cuh = log2_cuh; r.slice_type = pi->slice_type;
    r.ctx_skip = core->ctx_flags[CNID_SKIP_FLAG]; r.ctx_pred_mode = core->ctx_flags[CNID_PRED_MODE];
    r.all_preds = xeve_check_all_preds(core->tree_cons);
    r.qp[0] = core->qp_y; r.qp[1] = core->qp_u; r.qp[2] = core->qp_v;
    for(int l = 0; l < 2; l++) {
        r.num_refp[l] = ctx->rpm.num_refp[l];
        for(int k = 0; k < RH_MAXR; k++) {
            r.ref_pic[l][k] = -1; r.ref_poc[l][k] = -1;
            if(k < r.num_refp[l] && (l == 0 || pi->slice_type == SLICE_B)) {
                XEVE_PIC *rp = pi->refp[k][l].pic;
                r.ref_pic[l][k] = find_or_add_pic(rp, (int)rp->poc, 1);
                r.ref_poc[l][k] = (int)pi->refp[k][l].poc;
            }
        }
    }
    r.lambda_mv = pi->lambda_mv; r.max_search_range = pi->max_search_range;
    r.rate_idx = rate_index(core, log2_cuw, log2_cuh);
    r.state_in = (int)T.cu_sbac.n;
    sbac_pack(&core->s_curr_best[log2_cuw - 2][log2_cuh - 2], (RH_SBAC *)vec_push(&T.cu_sbac, 1));
    for(int i = 0; i < 3; i++) r.lambda[i] = core->lambda[i];
    r.dist_chroma_weight[0] = core->dist_chroma_weight[0]; r.dist_chroma_weight[1] = core->dist_chroma_weight[1];
    if(pi->slice_type == SLICE_B)
        xeve_get_mv_dir(pi->refp[0], ctx->poc.poc_val,
                        core->scup + ((1 << (log2_cuw - MIN_CU_LOG2)) - 1) + ((1 << (log2_cuh - MIN_CU_LOG2)) - 1) * ctx->w_scu,
                        core->scup, ctx->w_scu, ctx->h_scu, r.mv_dir, 0);
    r.me_first = (int)T.me.n;
    size_t ny = (size_t)1 << (log2_cuw + log2_cuh), nc = ny >> 2;
    r.out_off = -1;

    double cost = T.org_cu(ctx, core, x, y, log2_cuw, log2_cuh, mi, coef, rec, s_rec);

    r.me_cnt = (int)T.me.n - r.me_first;
    for(int l = 0; l < 2; l++)
        for(int k = 0; k < 4; k++) {
            r.mvp[l][k][0] = pi->mvp[l][k][0]; r.mvp[l][k][1] = pi->mvp[l][k][1];
            r.refi_pred[l][k] = pi->refi_pred[l][k];
        }
    r.cost = cost;
    int both = REFI_IS_VALID(mi->refi[0]) && REFI_IS_VALID(mi->refi[1]);
    r.best_idx = core->cu_mode == MODE_SKIP ? PRED_SKIP : core->cu_mode == MODE_DIR ? PRED_DIR
               : both ? PRED_BI : REFI_IS_VALID(mi->refi[0]) ? PRED_L0 : PRED_L1;
    for(int l = 0; l < 2; l++) {
        r.refi[l] = mi->refi[l]; r.mvp_idx[l] = mi->mvp_idx[l];
        r.mv[l][0] = mi->mv[l][0]; r.mv[l][1] = mi->mv[l][1]; r.mvd[l][0] = mi->mvd[l][0]; r.mvd[l][1] = mi->mvd[l][1];
    }
    for(int i = 0; i < 3; i++) r.nnz[i] = core->nnz[i];
    uint64_t hsh = FNV_INIT;
    hsh = fnv1a(hsh, coef[Y_C], ny * 2); hsh = fnv1a(hsh, coef[U_C], nc * 2); hsh = fnv1a(hsh, coef[V_C], nc * 2);
    r.coef_hash = hsh;
    hsh = FNV_INIT;
    hsh = fnv1a(hsh, rec[Y_C], ny * 2); hsh = fnv1a(hsh, rec[U_C], nc * 2); hsh = fnv1a(hsh, rec[V_C], nc * 2);
    r.rec_hash = hsh;
    r.state_out = (int)T.cu_sbac.n;
    sbac_pack(&core->s_next_best[log2_cuw - 2][log2_cuh - 2], (RH_SBAC *)vec_push(&T.cu_sbac, 1));
    *(RH_CU_REC *)vec_push(&T.cu, 1) = r;
    return cost;
}

/* ------------------------------------------------------------------------------------------
 * intra analysis trace: ctx->fn_pintra_analyze_cu (src_base/xeve_pintra.c:544-698)
 * ---------------------------------------------------------------------------------------- */
#define RH_T_INTRA 64
#define RH_T_INTRA_TIME 128 /* no records: only the time spent inside the reference's pintra_analyze_cu and the call count */
static double  g_intra_secs;
static int64_t g_intra_calls;
RH_API double rh_intra_time(int64_t *calls) { if(calls) *calls = g_intra_calls; return g_intra_secs; }
typedef struct {                /* == xb200_intra_item */
    int32_t  poc, cur_pic;
    int16_t  x, y;
    uint8_t  log2_cuw, log2_cuh, slice_type, ctx_skip, ctx_pred_mode, all_preds;
    uint8_t  qp[3];
    uint8_t  mpm[5];
    uint8_t  pad0_[2];
    uint32_t inter_satd;
    int32_t  rate_idx, state_in, state_out;
    uint16_t cm_ipm_in[2], cm_ipm_out[2];
    double   lambda[3], sqrt_lambda0, dist_chroma_weight[2];
    int64_t  nb_off, out_off;
    double   cost;
    int32_t  dist_cu;
    int8_t   ipm[2];
    uint8_t  pad1_[2];
    int32_t  nnz[3];
    uint64_t coef_hash, rec_hash;
} RH_INTRA_REC;
RH_API int rh_sizeof_intra(void) { return sizeof(RH_INTRA_REC); }

/* neighbour samples as xeve_get_nbr left them in core->nb: per plane left[-1 .. 2n-1] then up[-1 .. 2n-1] */
static int64_t push_nbr(XEVE_CORE *core, int cuw, int cuh)
{
    int64_t off = (int64_t)T.samp.n;
    for(int c = 0; c < 3; c++) {
        int w = c ? cuw >> 1 : cuw, h = c ? cuh >> 1 : cuh, n = w + h + 1;
        memcpy(vec_push(&T.samp, n), core->nb[c][0] + 2 - 1, n * sizeof(s16));
        memcpy(vec_push(&T.samp, n), core->nb[c][1] + h - 1, n * sizeof(s16));
    }
    return off;
}

static double hook_intra(XEVE_CTX *ctx, XEVE_CORE *core, int x, int y, int log2_cuw, int log2_cuh, XEVE_MODE *mi,
                         s16 coef[N_C][MAX_CU_DIM], pel *rec[N_C], int s_rec[N_C])
{
    if(tracing(RH_T_INTRA_TIME) && !tracing(RH_T_INTRA)) {
        const double t0 = now_s();
        const double c = T.org_intra(ctx, core, x, y, log2_cuw, log2_cuh, mi, coef, rec, s_rec);
        g_intra_secs += now_s() - t0;
        g_intra_calls++;
        return c;
    }
    if(!tracing(RH_T_INTRA)) return T.org_intra(ctx, core, x, y, log2_cuw, log2_cuh, mi, coef, rec, s_rec);
    XEVE_PINTRA *pi = &ctx->pintra[core->thread_cnt];
    RH_INTRA_REC r;
    memset(&r, 0, sizeof(r));
    r.poc = ctx->poc.poc_val;
    r.cur_pic = find_or_add_pic(pi->pic_o, r.poc, 0);
    r.x = x; r.y = y; r.log2_cuw = log2_cuw; r.log2_cuh = log2_cuh; r.slice_type = pi->slice_type;
    r.ctx_skip = core->ctx_flags[CNID_SKIP_FLAG]; r.ctx_pred_mode = core->ctx_flags[CNID_PRED_MODE];
    r.all_preds = xeve_check_all_preds(core->tree_cons);
    r.qp[0] = core->qp_y; r.qp[1] = core->qp_u; r.qp[2] = core->qp_v;
    r.inter_satd = core->inter_satd;
    r.rate_idx = rate_index(core, log2_cuw, log2_cuh);
    XEVE_SBAC *sin = &core->s_curr_best[log2_cuw - 2][log2_cuh - 2];
    r.state_in = (int)T.cu_sbac.n;
    sbac_pack(sin, (RH_SBAC *)vec_push(&T.cu_sbac, 1));
    memcpy(r.cm_ipm_in, sin->ctx.intra_dir, 4);
    for(int i = 0; i < 3; i++) r.lambda[i] = core->lambda[i];
    r.sqrt_lambda0 = core->sqrt_lambda[0];
    r.dist_chroma_weight[0] = core->dist_chroma_weight[0]; r.dist_chroma_weight[1] = core->dist_chroma_weight[1];
    r.out_off = -1;

    double cost = T.org_intra(ctx, core, x, y, log2_cuw, log2_cuh, mi, coef, rec, s_rec);

    r.nb_off = push_nbr(core, 1 << log2_cuw, 1 << log2_cuh);
    memcpy(r.mpm, core->mpm_b_list, 5);
    r.cost = cost;
    if(cost < MAX_COST) {
        size_t ny = (size_t)1 << (log2_cuw + log2_cuh), nc = ny >> 2;
        r.dist_cu = core->dist_cu;
        r.ipm[0] = core->ipm[0]; r.ipm[1] = core->ipm[1];
        for(int i = 0; i < 3; i++) r.nnz[i] = core->nnz[i];
        uint64_t hsh = FNV_INIT;
        hsh = fnv1a(hsh, coef[Y_C], ny * 2); hsh = fnv1a(hsh, coef[U_C], nc * 2); hsh = fnv1a(hsh, coef[V_C], nc * 2);
        r.coef_hash = hsh;
        hsh = FNV_INIT;
        hsh = fnv1a(hsh, rec[Y_C], ny * 2); hsh = fnv1a(hsh, rec[U_C], nc * 2); hsh = fnv1a(hsh, rec[V_C], nc * 2);
        r.rec_hash = hsh;
        memcpy(r.cm_ipm_out, core->s_temp_best.ctx.intra_dir, 4);
    }
    r.state_out = (int)T.cu_sbac.n;
    sbac_pack(&core->s_temp_best, (RH_SBAC *)vec_push(&T.cu_sbac, 1));
    *(RH_INTRA_REC *)vec_push(&T.intra, 1) = r;
    return cost;
}

/* the reference's xeve_get_avail_intra + xeve_get_nbr (Y, U, V) + xeve_get_mpm on caller-provided planes and maps */
typedef struct { int16_t x, y; uint8_t log2_cuw, log2_cuh; uint8_t mpm[5]; uint8_t pad_; uint16_t avail; uint16_t pad2_; int64_t nb_off; } RH_NBR_REC;
RH_API int rh_sizeof_nbr(void) { return sizeof(RH_NBR_REC); }
RH_API void rh_intra_nbr(const RH_PLANES *pl, RH_NBR_REC *items, int n, u32 *map_scu, s8 *map_ipm, int w_scu, int h_scu, int cip,
                         int bit_depth, s16 *side)
{
    static pel nb[N_C][N_REF][MAX_CU_SIZE * 3];
    u8 *tidx = calloc((size_t)w_scu * h_scu, 1);
    for(int i = 0; i < n; i++) {
        RH_NBR_REC *it = &items[i];
        const int   x = it->x, y = it->y, cuw = 1 << it->log2_cuw, cuh = 1 << it->log2_cuh, xs = x >> 2, ys = y >> 2, scup = xs + ys * w_scu;
        u16 av = xeve_get_avail_intra(xs, ys, w_scu, h_scu, scup, it->log2_cuw, it->log2_cuh, map_scu, tidx);
        it->avail = av;
        xeve_get_nbr(x, y, cuw, cuh, pl->y + y * pl->s_l + x, pl->s_l, av, nb, scup, map_scu, w_scu, h_scu, Y_C, cip, tidx, bit_depth, 1);
        xeve_get_nbr(x >> 1, y >> 1, cuw >> 1, cuh >> 1, pl->u + (y >> 1) * pl->s_c + (x >> 1), pl->s_c, av, nb, scup, map_scu, w_scu, h_scu,
                     U_C, cip, tidx, bit_depth, 1);
        xeve_get_nbr(x >> 1, y >> 1, cuw >> 1, cuh >> 1, pl->v + (y >> 1) * pl->s_c + (x >> 1), pl->s_c, av, nb, scup, map_scu, w_scu, h_scu,
                     V_C, cip, tidx, bit_depth, 1);
        u8 *mpm;
        xeve_get_mpm(xs, ys, cuw, cuh, map_scu, map_ipm, scup, w_scu, &mpm, tidx);
        memcpy(it->mpm, mpm, 5);
        s16 *out = side + it->nb_off;
        for(int c = 0; c < 3; c++) {
            int w = c ? cuw >> 1 : cuw, h = c ? cuh >> 1 : cuh, m = w + h + 1;
            memcpy(out, nb[c][0] + 2 - 1, m * sizeof(s16)); out += m;
            memcpy(out, nb[c][1] + h - 1, m * sizeof(s16)); out += m;
        }
    }
    free(tidx);
}

/* ------------------------------------------------------------------------------------------
 * CTU decision trace: ctx->fn_mode_analyze_lcu (mode_analyze_lcu -> mode_coding_tree, src_base/xeve_mode.c:2007-2374,
 * 2521-2608).  One record per CTU: the picture-level inputs of the decision pass, the coder state the CTU starts from
 * (core->s_curr_best[CTU], loaded from the bitstream coder, src_base/xeve_enc.c:139) and the one it leaves
 * (core->s_next_best[CTU]).  The colocated MV maps of the picture are stored once, with the first CTU.
 * ---------------------------------------------------------------------------------------- */
#define RH_T_LCU 256
typedef struct { RH_SBAC s; uint16_t ipm[2], split, pad_; } RH_STATE;
typedef struct {
    int32_t  poc, slice_type, lcu_num, x_pel, y_pel, tile_qp, cur_pic;
    int32_t  num_refp[2], ref_pic[2][RH_MAXR], ref_poc[2][RH_MAXR], col_list_poc0;
    int32_t  max_cu_inter, min_cu_inter, max_cu_intra, min_cu_intra, cip;
    int32_t  qp[3];
    uint32_t lambda_mv;
    int32_t  max_search_range, parallel_rows; /* ctx->parallel_rows: CTU rows y, y + n, ... share one coder-state chain */
    double   lambda[3], sqrt_lambda0, dist_chroma_weight[2];
    int64_t  col_off[2];            /* s16 offset in samp of refp[0][l].map_mv (f_scu x [2][2]); -1: none */
    RH_STATE state_in, state_out;
} RH_LCU_REC;
static void state_pack(const XEVE_SBAC *d, RH_STATE *s)
{
    memset(s, 0, sizeof(*s));
    sbac_pack(d, &s->s);
    memcpy(s->ipm, d->ctx.intra_dir, 4);
    memcpy(&s->split, d->ctx.split_cu_flag, 2);
}
/* Decision injection: ctx->fn_mode_analyze_lcu replaced by "take the CTU's decisions from outside" -- the reference-side binding
 * of a picture-level decision engine.  The records (oracle/xeve_oracle.h xo_scu_rec, coefficient planes, the
 * picture before deblocking) are put where mode_analyze_lcu leaves its own results: core->cu_data_best[CTU] ->
 * mode_cpy_rec_to_ref -> update_to_ctx_map -> ctx->map_cu_data[lcu] (src_base/xeve_mode.c:2521-2608).  The reference's own
 * entropy coder, loop filter and picture management then run unchanged on them. */
#define RH_T_INJECT 512
typedef struct { uint8_t mode, log2; int8_t ipm, refi[2]; uint8_t mvp_idx[2], pad_; int16_t mv[2][2], mvd[2][2]; int32_t nnz[3]; } RH_SCU_REC;
typedef struct { int32_t poc, s_l, s_c, pad_; const RH_SCU_REC *scu; const int16_t *coef, *rec_y, *rec_u, *rec_v; } RH_INJECT_PIC;
static const RH_INJECT_PIC *g_inject;
static int g_inject_n;
static int64_t g_inject_ctus;
RH_API void rh_inject(const RH_INJECT_PIC *pics, int n) { g_inject = pics; g_inject_n = n; g_inject_ctus = 0; }
/* Lazy variant: the records of a picture are asked for when the reference starts coding it (ctx->fn_mode_analyze_frame, the
 * picture-level hook the reference calls before its CTU loop, src_base/xeve_enc.c:322) -- a decision engine that runs ahead of the
 * entropy coder hands each picture over as soon as it is decided.  rec_* may be NULL (the reconstruction stays with the engine). */
typedef int (*rh_fetch_fn)(int poc, RH_INJECT_PIC *out);
static __thread rh_fetch_fn   g_fetch;      /* per calling thread: lazy hand-over needs threads = 1 (hooks run in the caller) */
static __thread RH_INJECT_PIC g_fetched;
static __thread int           g_fetched_ok;
static __thread int64_t       tl_inject_ctus;
RH_API void rh_inject_lazy(rh_fetch_fn f) { g_fetch = f; g_fetched_ok = 0; tl_inject_ctus = 0; }
RH_API int64_t rh_inject_lazy_count(void) { return tl_inject_ctus; }
/* The reference writes its parameters, `threads` among them, into an SEI message of the first access unit (xeve_param2string,
 * src_base/xeve_enc.c:2533-2537 via xeve_eco_emitsei).  A host pass that runs single-threaded over decisions made as T coder-state
 * chains labels the stream with T, like the reference run it reproduces: ctx->fn_enc_header sees param.threads = T, nothing else does. */
static __thread int g_label_threads;
static __thread int (*g_org_header)(XEVE_CTX *);
RH_API void rh_label_threads(int t) { g_label_threads = t; }
static int hook_header(XEVE_CTX *ctx)
{
    const int keep = ctx->param.threads;
    if(g_label_threads > 0) ctx->param.threads = g_label_threads;
    const int ret = g_org_header(ctx);
    ctx->param.threads = keep;
    return ret;
}
/* Plan mode (RH_T_PLAN): the reference's own control plane run dry -- every CTU gets a trivial decision (8x8 SKIP / 8x8 intra DC
 * units, no residual), so a sequence is "coded" in milliseconds per picture while slice types, POCs, QPs, lambdas and reference lists
 * come out exactly as a real encode computes them (constant QP: none of them depends on a decision).  One RH_LCU_REC (CTU 0) and one
 * RH_DF_REC (parameters only) per picture, in coding order: the picture-level inputs of xb200_analyze_picture for a whole clip. */
#define RH_T_PLAN 1024
#define RH_T_NO_LF 2048  /* ctx->fn_loop_filter does nothing: the deblocked picture lives on the device */
static RH_SCU_REC *g_plan_scu[2];   /* [inter, intra] dummy records of one CTU */
static int16_t    *g_plan_zero;
RH_API int64_t rh_inject_count(void) { return g_inject_ctus; }
RH_API int rh_sizeof_inject(int what) { return what == 0 ? sizeof(RH_SCU_REC) : sizeof(RH_INJECT_PIC); }

static void inject_split(XEVE_CTX *ctx, XEVE_CU_DATA *cd, const RH_SCU_REC *scu, int x0, int y0, int x, int y, int log2, int cud, int cup)
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
static int inject_lcu(XEVE_CTX *ctx, XEVE_CORE *core)
{
    const RH_INJECT_PIC *ip = NULL;
    RH_INJECT_PIC        plan;
    for(int i = 0; i < g_inject_n; i++) if(g_inject[i].poc == (int)ctx->poc.poc_val) ip = &g_inject[i];
    if(!ip && g_fetched_ok && g_fetched.poc == (int)ctx->poc.poc_val) ip = &g_fetched;
    if(!ip && (T.mask & RH_T_PLAN)) {
        memset(&plan, 0, sizeof(plan));
        plan.poc = (int)ctx->poc.poc_val;
        plan.scu = g_plan_scu[ctx->slice_type == SLICE_I] - (size_t)core->lcu_num * 256;
        plan.coef = g_plan_zero - (size_t)core->lcu_num * 6144;
        ip = &plan;
    }
    if(!ip) return XEVE_ERR;
    const int L = ctx->log2_max_cuwh - 2, x0 = core->x_pel, y0 = core->y_pel, q = ctx->tile[core->tile_idx].qp;
    const int bdc = ctx->sps.bit_depth_chroma_minus8;
    const int qp_y = GET_LUMA_QP(q, ctx->sps.bit_depth_luma_minus8);
    const int qp_u = ctx->qp_chroma_dynamic[0][XEVE_CLIP3(-6 * bdc, 57, q + ctx->sh->qp_u_offset)] + 6 * bdc;
    const int qp_v = ctx->qp_chroma_dynamic[1][XEVE_CLIP3(-6 * bdc, 57, q + ctx->sh->qp_v_offset)] + 6 * bdc;
    const RH_SCU_REC *scu = ip->scu + (size_t)core->lcu_num * 256;
    const int16_t    *coef = ip->coef + (size_t)core->lcu_num * 6144;
    XEVE_CU_DATA     *cd = &core->cu_data_best[L][L];
    init_cu_data(cd, ctx->log2_max_cuwh, ctx->log2_max_cuwh, ctx->qp, ctx->qp, ctx->qp);
    inject_split(ctx, cd, scu, x0, y0, x0, y0, ctx->log2_max_cuwh, 0, 0);
    static const u8 mode_of[4] = {MODE_SKIP, MODE_DIR, MODE_INTER, MODE_INTRA};
    for(int i = 0; i < 256; i++) {
        const RH_SCU_REC *r = &scu[i];
        if(x0 + (i & 15) * 4 >= ctx->w || y0 + (i >> 4) * 4 >= ctx->h) continue;
        const int mode = mode_of[r->mode];
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
    if(ip->rec_y) {
        for(int j = 0; j < 64 && y0 + j < ctx->h; j++)
            memcpy(cd->reco[Y_C] + j * 64, ip->rec_y + (size_t)(y0 + j) * ip->s_l + x0, 2 * XEVE_MIN(64, ctx->w - x0));
        for(int j = 0; j < 32 && y0 / 2 + j < ctx->h / 2; j++) {
            memcpy(cd->reco[U_C] + j * 32, ip->rec_u + (size_t)(y0 / 2 + j) * ip->s_c + x0 / 2, 2 * XEVE_MIN(32, (ctx->w - x0) / 2));
            memcpy(cd->reco[V_C] + j * 32, ip->rec_v + (size_t)(y0 / 2 + j) * ip->s_c + x0 / 2, 2 * XEVE_MIN(32, (ctx->w - x0) / 2));
        }
        mode_cpy_rec_to_ref(core, x0, y0, ctx->max_cuwh, ctx->max_cuwh, PIC_MODE(ctx), xeve_get_default_tree_cons(), ctx->sps.chroma_format_idc);
    }
    update_to_ctx_map(ctx, core);
    copy_cu_data(&ctx->map_cu_data[core->lcu_num], cd, 0, 0, ctx->log2_max_cuwh, ctx->log2_max_cuwh, ctx->log2_max_cuwh, 0,
                 xeve_get_default_tree_cons(), ctx->sps.chroma_format_idc);
    const int xs = x0 >> 2, ys = y0 >> 2, w = XEVE_MIN(16, ctx->w_scu - xs), h = XEVE_MIN(16, ctx->h_scu - ys);
    for(int j = 0; j < h; j++)
        for(int i = 0; i < w; i++) MCU_CLR_COD(ctx->map_scu[(size_t)(ys + j) * ctx->w_scu + xs + i]);
    if(g_fetch) tl_inject_ctus++;
    else __atomic_fetch_add(&g_inject_ctus, 1, __ATOMIC_RELAXED);
    return XEVE_OK;
}
static void lcu_fill_picture_fields(XEVE_CTX *ctx, XEVE_CORE *core, RH_LCU_REC *r);
static void plan_push_df(XEVE_CTX *ctx);
static int hook_analyze_frame(XEVE_CTX *ctx)
{
    if((T.mask & RH_T_INJECT) && g_fetch) {
        g_fetched_ok = 0;
        memset(&g_fetched, 0, sizeof(g_fetched));
        if(g_fetch((int)ctx->poc.poc_val, &g_fetched) != 0) return XEVE_ERR;
        g_fetched_ok = 1;
    }
    return T.org_frame ? T.org_frame(ctx) : XEVE_OK;
}
static int hook_lcu(XEVE_CTX *ctx, XEVE_CORE *core)
{
    if(T.mask & RH_T_PLAN) {
        if(core->lcu_num == 0) {
            static pthread_mutex_t pm = PTHREAD_MUTEX_INITIALIZER;
            pthread_mutex_lock(&pm);
            RH_LCU_REC r;
            memset(&r, 0, sizeof(r));
            lcu_fill_picture_fields(ctx, core, &r);
            *(RH_LCU_REC *)vec_push(&T.lcu, 1) = r;
            plan_push_df(ctx);
            pthread_mutex_unlock(&pm);
        }
        return inject_lcu(ctx, core);
    }
    if((T.mask & RH_T_INJECT) && (g_inject || g_fetch)) return inject_lcu(ctx, core);
    if(!tracing(RH_T_LCU)) return T.org_lcu(ctx, core);
    static int64_t col_off[2];
    static pthread_mutex_t mtx = PTHREAD_MUTEX_INITIALIZER;   /* with threads > 1 the CTU rows run concurrently */
    pthread_mutex_lock(&mtx);
    XEVE_PINTER *pi = &ctx->pinter[core->thread_cnt];
    XEVE_PINTRA *pa = &ctx->pintra[core->thread_cnt];
    RH_LCU_REC   r;
    memset(&r, 0, sizeof(r));
    const int L = ctx->log2_max_cuwh - 2;
    r.poc = ctx->poc.poc_val; r.slice_type = ctx->slice_type; r.lcu_num = core->lcu_num;
    r.x_pel = core->x_pel; r.y_pel = core->y_pel; r.tile_qp = ctx->tile[core->tile_idx].qp;
    r.cur_pic = find_or_add_pic(pa->pic_o, r.poc, 0);
    for(int l = 0; l < 2; l++) {
        r.num_refp[l] = ctx->rpm.num_refp[l];
        for(int k = 0; k < RH_MAXR; k++) {
            r.ref_pic[l][k] = -1; r.ref_poc[l][k] = -1;
            if(ctx->slice_type != SLICE_I && k < r.num_refp[l] && (l == 0 || ctx->slice_type == SLICE_B)) {
                XEVE_PIC *rp = ctx->refp[k][l].pic;
                r.ref_pic[l][k] = find_or_add_pic(rp, (int)rp->poc, 1);
                r.ref_poc[l][k] = (int)ctx->refp[k][l].poc;
            }
        }
    }
    if(core->lcu_num == 0) {
        col_off[0] = col_off[1] = -1;
        if(ctx->slice_type != SLICE_I)
            for(int l = 0; l < 2; l++) {
                if(l == 1 && ctx->slice_type != SLICE_B) break;
                if(!ctx->refp[0][l].map_mv) continue;
                col_off[l] = (int64_t)T.samp.n;
                memcpy(vec_push(&T.samp, (size_t)ctx->f_scu * 4), ctx->refp[0][l].map_mv, (size_t)ctx->f_scu * 8);
            }
    }
    r.col_off[0] = col_off[0]; r.col_off[1] = col_off[1];
    if(ctx->slice_type == SLICE_B && ctx->refp[0][REFP_1].list_poc) r.col_list_poc0 = (int)ctx->refp[0][REFP_1].list_poc[0];
    r.max_cu_inter = ctx->param.max_cu_inter; r.min_cu_inter = ctx->param.min_cu_inter;
    r.max_cu_intra = ctx->param.max_cu_intra; r.min_cu_intra = ctx->param.min_cu_intra;
    r.cip = ctx->pps.constrained_intra_pred_flag;
    {   /* mode_cu_init, src_base/xeve_mode.c:776-783, with core->qp == tile qp (no delta QP) */
        const int q = r.tile_qp, bdc = ctx->sps.bit_depth_chroma_minus8;
        r.qp[0] = GET_LUMA_QP(q, ctx->sps.bit_depth_luma_minus8);
        r.qp[1] = ctx->qp_chroma_dynamic[0][XEVE_CLIP3(-6 * bdc, 57, q + ctx->sh->qp_u_offset)] + 6 * bdc;
        r.qp[2] = ctx->qp_chroma_dynamic[1][XEVE_CLIP3(-6 * bdc, 57, q + ctx->sh->qp_v_offset)] + 6 * bdc;
    }
    r.lambda_mv = pi->lambda_mv; r.max_search_range = pi->max_search_range;
    for(int i = 0; i < 3; i++) r.lambda[i] = core->lambda[i];
    r.sqrt_lambda0 = core->sqrt_lambda[0];
    r.dist_chroma_weight[0] = core->dist_chroma_weight[0]; r.dist_chroma_weight[1] = core->dist_chroma_weight[1];
    r.parallel_rows = ctx->parallel_rows;
    state_pack(&core->s_curr_best[L][L], &r.state_in);
    pthread_mutex_unlock(&mtx);
    int ret = T.org_lcu(ctx, core);
    state_pack(&core->s_next_best[L][L], &r.state_out);
    pthread_mutex_lock(&mtx);
    *(RH_LCU_REC *)vec_push(&T.lcu, 1) = r;
    pthread_mutex_unlock(&mtx);
    return ret;
}
RH_API int rh_sizeof_lcu(void) { return sizeof(RH_LCU_REC); }
/* the picture-level fields of an RH_LCU_REC without touching any picture data (plan mode): reference pictures are named by POC only
 * (ref_pic = 0 marks a valid entry), no colocated maps, no coder states */
static void lcu_fill_picture_fields(XEVE_CTX *ctx, XEVE_CORE *core, RH_LCU_REC *r)
{
    XEVE_PINTER *pi = &ctx->pinter[core->thread_cnt];
    r->poc = ctx->poc.poc_val; r->slice_type = ctx->slice_type; r->lcu_num = core->lcu_num;
    r->x_pel = core->x_pel; r->y_pel = core->y_pel; r->tile_qp = ctx->tile[core->tile_idx].qp;
    r->cur_pic = -1;
    for(int l = 0; l < 2; l++) {
        r->num_refp[l] = ctx->rpm.num_refp[l];
        for(int k = 0; k < RH_MAXR; k++) {
            r->ref_pic[l][k] = -1; r->ref_poc[l][k] = -1;
            if(ctx->slice_type != SLICE_I && k < r->num_refp[l] && (l == 0 || ctx->slice_type == SLICE_B)) {
                r->ref_pic[l][k] = 0;
                r->ref_poc[l][k] = (int)ctx->refp[k][l].poc;
            }
        }
    }
    r->col_off[0] = r->col_off[1] = -1;
    if(ctx->slice_type == SLICE_B && ctx->refp[0][REFP_1].list_poc) r->col_list_poc0 = (int)ctx->refp[0][REFP_1].list_poc[0];
    r->max_cu_inter = ctx->param.max_cu_inter; r->min_cu_inter = ctx->param.min_cu_inter;
    r->max_cu_intra = ctx->param.max_cu_intra; r->min_cu_intra = ctx->param.min_cu_intra;
    r->cip = ctx->pps.constrained_intra_pred_flag;
    {
        const int q = r->tile_qp, bdc = ctx->sps.bit_depth_chroma_minus8;
        r->qp[0] = GET_LUMA_QP(q, ctx->sps.bit_depth_luma_minus8);
        r->qp[1] = ctx->qp_chroma_dynamic[0][XEVE_CLIP3(-6 * bdc, 57, q + ctx->sh->qp_u_offset)] + 6 * bdc;
        r->qp[2] = ctx->qp_chroma_dynamic[1][XEVE_CLIP3(-6 * bdc, 57, q + ctx->sh->qp_v_offset)] + 6 * bdc;
    }
    r->lambda_mv = pi->lambda_mv; r->max_search_range = pi->max_search_range;
    for(int i = 0; i < 3; i++) r->lambda[i] = core->lambda[i];
    r->sqrt_lambda0 = core->sqrt_lambda[0];
    r->dist_chroma_weight[0] = core->dist_chroma_weight[0]; r->dist_chroma_weight[1] = core->dist_chroma_weight[1];
    r->parallel_rows = ctx->parallel_rows;
}

/* ------------------------------------------------------------------------------------------
 * deblocking trace: ctx->fn_loop_filter with the picture before / after, the frame
 * maps it reads and the CU rectangles xeve_deblock_tree enumerates (ctx->fn_deblock_unit)
 * ---------------------------------------------------------------------------------------- */
#define RH_T_DF 32
typedef struct { int16_t x, y; uint8_t log2_cuw, log2_cuh, pad_[2]; } RH_DF_CU;
typedef struct {                /* == xb200_df_pic */
    int32_t w_scu, h_scu, qp_u_offset, qp_v_offset;
    int32_t chroma_qp[2][70];   /* ctx->qp_chroma_dynamic[c][q] at [q + 6 * (bit_depth - 8)] */
} RH_DF_PIC;
typedef struct {
    int32_t   poc, pre_pic, post_pic, on;
    int64_t   cu_first, cu_cnt;
    int64_t   maps_off;         /* byte offset: map_scu u32[f] | map_refi s8[f][2] | map_mv s16[f][2][2] */
    RH_DF_PIC pp;
} RH_DF_REC;

static int ilog2(int v) { int l = 0; while((1 << l) < v) l++; return l; }

static void hook_df_unit(XEVE_CTX *ctx, XEVE_PIC *pic, int x, int y, int cuw, int cuh, int is_hor_edge, XEVE_CORE *core, int bf)
{
    if(T.df_collect && !is_hor_edge) {
        RH_DF_CU c = {(int16_t)x, (int16_t)y, (uint8_t)ilog2(cuw), (uint8_t)ilog2(cuh), {0, 0}};
        *(RH_DF_CU *)vec_push(&T.df_cu, 1) = c;
    }
    T.org_df_unit(ctx, pic, x, y, cuw, cuh, is_hor_edge, core, bf);
}

static void df_fill_pic(XEVE_CTX *ctx, RH_DF_PIC *pp)
{
    const int bdo = 6 * (ctx->param.codec_bit_depth - 8);
    memset(pp, 0, sizeof(*pp));
    pp->w_scu = ctx->w_scu; pp->h_scu = ctx->h_scu;
    pp->qp_u_offset = ctx->sh->qp_u_offset; pp->qp_v_offset = ctx->sh->qp_v_offset;
    for(int c = 0; c < 2; c++)
        for(int q = -bdo; q <= 57; q++) pp->chroma_qp[c][q + bdo] = ctx->qp_chroma_dynamic[c][q];
}

static void plan_push_df(XEVE_CTX *ctx)
{
    RH_DF_REC r;
    memset(&r, 0, sizeof(r));
    r.poc = ctx->poc.poc_val;
    r.on  = ctx->sh->deblocking_filter_on;
    r.pre_pic = r.post_pic = -1;
    r.maps_off = -1;
    df_fill_pic(ctx, &r.pp);
    *(RH_DF_REC *)vec_push(&T.df, 1) = r;
}
static int hook_loop_filter(XEVE_CTX *ctx, XEVE_CORE *core)
{
    if(T.mask & (RH_T_PLAN | RH_T_NO_LF)) return XEVE_OK;
    if(!tracing(RH_T_DF)) return T.org_lf(ctx, core);
    RH_DF_REC r;
    memset(&r, 0, sizeof(r));
    r.poc = ctx->poc.poc_val;
    r.on  = ctx->sh->deblocking_filter_on;
    r.pre_pic = find_or_add_pic(PIC_CURR(ctx), r.poc, 2);
    df_fill_pic(ctx, &r.pp);
    const size_t f = ctx->f_scu;
    r.maps_off = (int64_t)T.df_maps.n;
    uint8_t *m = vec_push(&T.df_maps, f * 14);
    memcpy(m, ctx->map_scu, f * 4); memcpy(m + f * 4, ctx->map_refi, f * 2); memcpy(m + f * 6, ctx->map_mv, f * 8);
    r.cu_first = (int64_t)T.df_cu.n;
    T.df_collect = 1;
    int ret = T.org_lf(ctx, core);
    T.df_collect = 0;
    r.cu_cnt = (int64_t)T.df_cu.n - r.cu_first;
    r.post_pic = find_or_add_pic(PIC_CURR(ctx), r.poc, 3);
    *(RH_DF_REC *)vec_push(&T.df, 1) = r;
    return ret;
}

/* The reference's own edge filters (exported xeve_deblock_cu_ver / _hor, src_base/xeve_df.c:253-471) driven like
 * xeve_loop_filter + xeve_deblock (src_base/xeve_enc.c:2355-2414, xeve_df.c:522-573): all vertical edges, then all
 * horizontal ones, CUs in the order given.  Planes are filtered in place. */
RH_API int rh_sizeof_df(int what) { return what == 0 ? sizeof(RH_DF_REC) : what == 1 ? sizeof(RH_DF_CU) : sizeof(RH_DF_PIC); }
RH_API double rh_deblock(const RH_PLANES *pl, const RH_DF_CU *cus, int64_t n, const RH_DF_PIC *pp, const u32 *map_scu_in,
                         s8 (*map_refi)[REFP_NUM], s16 (*map_mv)[REFP_NUM][MV_D], int bit_depth)
{
    XEVE_PIC pic;
    memset(&pic, 0, sizeof(pic));
    pic.y = pl->y; pic.u = pl->u; pic.v = pl->v; pic.s_l = pl->s_l; pic.s_c = pl->s_c; pic.w_l = pl->w_l; pic.h_l = pl->h_l;
    pic.w_c = pl->w_l / 2; pic.h_c = pl->h_l / 2;
    pic.pic_qp_u_offset = pp->qp_u_offset; pic.pic_qp_v_offset = pp->qp_v_offset;
    const size_t f = (size_t)pp->w_scu * pp->h_scu;
    u32 *map_scu = malloc(f * 4);
    u8  *tidx = calloc(f, 1);
    memcpy(map_scu, map_scu_in, f * 4);
    const int bdo = 6 * (bit_depth - 8);
    int *tbl[2] = {(int *)pp->chroma_qp[0] + bdo, (int *)pp->chroma_qp[1] + bdo};
    double t0 = now_s();
    for(int is_hor = 0; is_hor <= 1; is_hor++) {
        for(size_t i = 0; i < f; i++) MCU_CLR_COD(map_scu[i]);
        for(int64_t i = 0; i < n; i++) {
            const RH_DF_CU *c = &cus[i];
            if(is_hor)
                xeve_deblock_cu_hor(&pic, c->x, c->y, 1 << c->log2_cuw, 1 << c->log2_cuh, map_scu, map_refi, map_mv, pp->w_scu,
                                    xeve_get_default_tree_cons(), tidx, 0, bit_depth, bit_depth, 1, tbl);
            else
                xeve_deblock_cu_ver(&pic, c->x, c->y, 1 << c->log2_cuw, 1 << c->log2_cuh, map_scu, map_refi, map_mv, pp->w_scu,
                                    NULL, xeve_get_default_tree_cons(), tidx, 0, bit_depth, bit_depth, 1, tbl);
        }
    }
    double t = now_s() - t0;
    free(map_scu); free(tidx);
    return t;
}

/* ------------------------------------------------------------------------------------------
 * encoder session
 * ---------------------------------------------------------------------------------------- */
static int img_addref(XEVE_IMGB *i) { return ++i->refcnt; }
static int img_getref(XEVE_IMGB *i) { return i->refcnt; }
static int img_release(XEVE_IMGB *i) { return --i->refcnt; }

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static XEVE make_encoder(int w, int h, int in_depth, int preset, int qp, int threads, int bframes,
                         const char *extra, int *err)
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
    /* extra "name=value;name=value" overrides through the reference's own parser */
    if(extra && *extra) {
        char *dup = strdup(extra), *save = NULL;
        for(char *tok = strtok_r(dup, ";", &save); tok; tok = strtok_r(NULL, ";", &save)) {
            char *eq = strchr(tok, '=');
            if(!eq) continue;
            *eq = 0;
            if(xeve_param_parse(p, tok, eq + 1) != XEVE_OK) fprintf(stderr, "rh: bad param %s\n", tok);
        }
        free(dup);
    }
    cdsc.max_bs_buf_size = 16 * 1024 * 1024;
    if(xeve_param_check(p) != XEVE_OK) { *err = -2; return NULL; }
    return xeve_create(&cdsc, err);
}

/* Encode `nframes` frames of planar I420 held in memory; returns seconds spent inside
 * xeve_encode (the quantity the reference app reports, app/xeve_app.c:1236-1246), or <0. */
RH_API double rh_encode_clip(const void *yuv, int nframes, int w, int h, int in_depth, int preset, int qp,
                             int threads, int bframes, const char *extra, int trace_mask, int pic_lo, int pic_hi,
                             uint8_t *bs_out, int64_t bs_cap, int64_t *bs_len)
{
    int  err = 0;
    XEVE id  = make_encoder(w, h, in_depth, preset, qp, threads, bframes, extra, &err);
    if(!id) return -1.0;
    XEVE_CTX *ctx = (XEVE_CTX *)id;
    RH_STATE_T *own = NULL;
    if((trace_mask & RH_T_ISOLATED) && threads == 1) { own = calloc(1, sizeof(RH_STATE_T)); tl_T = own; }
    trace_mask &= ~RH_T_ISOLATED;

    vec_reset(&T.me, sizeof(RH_ME_REC)); vec_reset(&T.mc, sizeof(RH_MC_REC)); vec_reset(&T.tq, sizeof(RH_TQ_REC));
    vec_reset(&T.rates, sizeof(RH_RATES)); vec_reset(&T.pics, sizeof(RH_PIC)); vec_reset(&T.samp, sizeof(s16)); vec_reset(&T.sbac, sizeof(RH_SBAC)); vec_reset(&T.cu, sizeof(RH_CU_REC)); vec_reset(&T.cu_sbac, sizeof(RH_SBAC));
    vec_reset(&T.intra, sizeof(RH_INTRA_REC)); vec_reset(&T.lcu, sizeof(RH_LCU_REC));
    vec_reset(&T.df, sizeof(RH_DF_REC)); vec_reset(&T.df_cu, sizeof(RH_DF_CU)); vec_reset(&T.df_maps, 1); T.df_collect = 0;
    T.have_rates = 0;
    g_cu_secs = 0; g_cu_calls = 0; g_intra_secs = 0; g_intra_calls = 0;
    T.ctx = ctx; T.mask = trace_mask; T.pic_lo = pic_lo; T.pic_hi = pic_hi;
    if(trace_mask) {
        T.org_me = ctx->pinter[0].fn_me; T.org_mc = ctx->pinter[0].fn_mc; T.org_tq = ctx->fn_tq;
        for(int i = 0; i < ctx->param.threads; i++) { ctx->pinter[i].fn_me = hook_me; ctx->pinter[i].fn_mc = hook_mc; }
        ctx->fn_tq = hook_tq;
        T.org_cu = ctx->fn_pinter_analyze_cu;
        ctx->fn_pinter_analyze_cu = hook_cu;
        T.org_intra = ctx->fn_pintra_analyze_cu; ctx->fn_pintra_analyze_cu = hook_intra;
        T.org_lf = ctx->fn_loop_filter; T.org_df_unit = ctx->fn_deblock_unit;
        ctx->fn_loop_filter = hook_loop_filter; ctx->fn_deblock_unit = hook_df_unit;
        T.org_lcu = ctx->fn_mode_analyze_lcu; ctx->fn_mode_analyze_lcu = hook_lcu;
        T.org_frame = ctx->fn_mode_analyze_frame; ctx->fn_mode_analyze_frame = hook_analyze_frame;
        if(g_label_threads > 0) { g_org_header = ctx->fn_enc_header; ctx->fn_enc_header = hook_header; }
    }
    if(trace_mask & RH_T_PLAN) {     /* dummy decisions of one CTU: 8x8 units, SKIP with zero motion / intra DC, no residual */
        for(int k = 0; k < 2; k++) {
            if(!g_plan_scu[k]) g_plan_scu[k] = calloc(256, sizeof(RH_SCU_REC));
            for(int i = 0; i < 256; i++) {
                RH_SCU_REC *u = &g_plan_scu[k][i];
                memset(u, 0, sizeof(*u));
                u->log2 = 3;
                if(k) { u->mode = 3; u->refi[0] = u->refi[1] = -1; }
                else  { u->mode = 0; u->refi[0] = 0; u->refi[1] = 0; }
            }
        }
        if(!g_plan_zero) g_plan_zero = calloc(6144, sizeof(int16_t));
    }
    XEVE_PINTER *pi = &ctx->pinter[0];
    T.cst.w = w; T.cst.h = h; T.cst.bit_depth = ctx->param.codec_bit_depth; T.cst.me_level = pi->me_level;
    T.cst.hpel_cnt = pi->search_pattern_hpel_cnt; T.cst.qpel_cnt = pi->search_pattern_qpel_cnt;
    T.cst.me_complexity = pi->me_complexity;
    T.cst.min_clip[0] = pi->min_clip[0]; T.cst.min_clip[1] = pi->min_clip[1];
    T.cst.max_clip[0] = pi->max_clip[0]; T.cst.max_clip[1] = pi->max_clip[1];
    T.cst.merge_num = ctx->param.merge_num; T.cst.me_range = ctx->param.me_range; T.cst.gop_size = ctx->param.gop_size;
    T.cst.rdoq = ctx->param.rdoq; T.cst.tool_iqt = ctx->param.tool_iqt;

    int       bps = in_depth > 8 ? 2 : 1;
    size_t    fsz = (size_t)w * h * 3 / 2 * bps;
    XEVE_IMGB img;
    uint8_t  *bs = malloc(16 * 1024 * 1024);
    XEVE_BITB bitb;
    XEVE_STAT stat;
    memset(&bitb, 0, sizeof(bitb));
    bitb.addr = bs; bitb.bsize = 16 * 1024 * 1024;
    double  t_enc = 0;
    int64_t total = 0;
    int     pushed = 0, bumping = 0, ret;
    while(1) {
        if(!bumping) {
            if(pushed < nframes) {
                const uint8_t *f = (const uint8_t *)yuv + fsz * pushed;
                memset(&img, 0, sizeof(img));
                img.cs = XEVE_CS_SET(XEVE_CF_YCBCR420, in_depth, 0);
                img.np = 3;
                for(int c = 0; c < 3; c++) {
                    int cw = c ? w / 2 : w, ch = c ? h / 2 : h;
                    img.w[c] = img.aw[c] = cw; img.h[c] = img.ah[c] = ch; img.s[c] = cw * bps; img.e[c] = ch;
                }
                img.a[0] = (void *)f; img.a[1] = (void *)(f + (size_t)w * h * bps);
                img.a[2] = (void *)(f + (size_t)w * h * bps * 5 / 4);
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
        double t0 = now_s();
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
    free(bs);
    T.ctx = NULL;
    xeve_delete(id);
    if(own) {   /* private state of an isolated encode: nothing to hand back but the bitstream */
        vec_t *vs[] = {&own->cu, &own->cu_sbac, &own->intra, &own->df, &own->df_cu, &own->df_maps, &own->lcu, &own->me, &own->mc, &own->tq,
                       &own->rates, &own->pics, &own->samp, &own->sbac};
        for(size_t i = 0; i < sizeof(vs) / sizeof(vs[0]); i++) free(vs[i]->p);
        tl_T = NULL;
        free(own);
    }
    return t_enc;
}

RH_API int64_t rh_trace_get(int what, void **ptr)
{
    vec_t *v = what == 0 ? &T.me : what == 1 ? &T.mc : what == 2 ? &T.tq : what == 3 ? &T.rates
             : what == 4 ? &T.pics : what == 5 ? &T.samp : what == 6 ? &T.sbac : what == 7 ? &T.cu : what == 8 ? &T.cu_sbac
             : what == 9 ? &T.df : what == 10 ? &T.df_cu : what == 11 ? &T.df_maps : what == 13 ? &T.lcu : &T.intra;
    *ptr = v->p;
    return (int64_t)v->n;
}
RH_API void rh_trace_const(RH_CONST *out) { *out = T.cst; }
RH_API int  rh_sizeof(int what)
{
    switch(what) {
    case 0: return sizeof(RH_ME_REC);
    case 1: return sizeof(RH_MC_REC);
    case 2: return sizeof(RH_TQ_REC);
    case 3: return sizeof(RH_RATES);
    case 4: return sizeof(RH_PIC);
    case 5: return sizeof(RH_PLANES);
    case 6: return sizeof(RH_CONST);
    case 7: return sizeof(RH_CU_REC);
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------
 * utility context: a small real encoder instance whose ctx supplies the function pointers,
 * err_scale table and parameter block the replay functions need
 * ---------------------------------------------------------------------------------------- */
static XEVE_CTX *g_util;
static XEVE_CTX *util_ctx(void)
{
    if(!g_util) {
        int err = 0;
        g_util = (XEVE_CTX *)make_encoder(176, 144, 8, XEVE_PRESET_FAST, 32, 1, -1, NULL, &err);
        /* the SPS is normally filled when the first picture is coded (src_base/xeve_enc.c:894) */
        if(g_util) xeve_set_sps(g_util, &g_util->sps);
    }
    return g_util;
}

/* ------------------------------------------------------------------------------------------
 * kernel probes.  variant: 0 = C table, 1 = SSE, 2 = AVX2
 * ---------------------------------------------------------------------------------------- */
#include "xeve_sad_sse.h"
#include "xeve_sad_avx.h"
#include "xeve_mc_sse.h"
#include "xeve_mc_avx.h"
#include "xeve_itdq_sse.h"
#include "xeve_itdq_avx.h"
#include "xeve_tq_avx.h"

RH_API int rh_sad(int variant, int log2w, int log2h, void *s1, void *s2, int st1, int st2, int bd)
{
    const XEVE_FN_SAD(*t)[8] = variant == 2 ? xeve_tbl_sad_16b_avx : variant == 1 ? xeve_tbl_sad_16b_sse : xeve_tbl_sad_16b;
    XEVE_FN_SAD f = t[log2w][log2h];
    if(!f) f = xeve_tbl_sad_16b[0][0];
    return f(1 << log2w, 1 << log2h, s1, s2, st1, st2, bd);
}
RH_API int64_t rh_ssd(int variant, int log2w, int log2h, void *s1, void *s2, int st1, int st2, int bd)
{
    const XEVE_FN_SSD(*t)[8] = variant ? xeve_tbl_ssd_16b_sse : xeve_tbl_ssd_16b;
    XEVE_FN_SSD f = t[log2w][log2h];
    if(!f) f = xeve_tbl_ssd_16b[0][0];
    return f(1 << log2w, 1 << log2h, s1, s2, st1, st2, bd);
}
RH_API void rh_diff(int variant, int log2w, int log2h, void *s1, void *s2, int st1, int st2, int sd, s16 *d, int bd)
{
    const XEVE_FN_DIFF(*t)[8] = variant ? xeve_tbl_diff_16b_sse : xeve_tbl_diff_16b;
    XEVE_FN_DIFF f = t[log2w][log2h];
    if(!f) f = xeve_tbl_diff_16b[0][0];
    f(1 << log2w, 1 << log2h, s1, s2, st1, st2, sd, d, bd);
}
RH_API int rh_satd(int variant, int w, int h, void *s1, void *s2, int st1, int st2, int bd)
{
    return (variant ? xeve_tbl_satd_16b_sse : xeve_tbl_satd_16b)[0](w, h, s1, s2, st1, st2, bd);
}
/* gmv in 1/16 (luma) or 1/32 (chroma) pel, absolute; ref = top-left of active area */
RH_API void rh_mc_l(int variant, s16 *ref, int gmv_x, int gmv_y, int s_ref, int s_pred, s16 *pred, int w, int h, int bd)
{
    const XEVE_MC_L(*t)[2] = variant == 2 ? xeve_tbl_mc_l_avx : variant == 1 ? xeve_tbl_mc_l_sse : xeve_tbl_mc_l;
    t[(gmv_x & 15) ? 1 : 0][(gmv_y & 15) ? 1 : 0](ref, gmv_x, gmv_y, s_ref, s_pred, pred, w, h, bd, xeve_tbl_mc_l_coeff);
}
RH_API void rh_mc_c(int variant, s16 *ref, int gmv_x, int gmv_y, int s_ref, int s_pred, s16 *pred, int w, int h, int bd)
{
    const XEVE_MC_C(*t)[2] = variant == 2 ? xeve_tbl_mc_c_avx : variant == 1 ? xeve_tbl_mc_c_sse : xeve_tbl_mc_c;
    t[(gmv_x & 31) ? 1 : 0][(gmv_y & 31) ? 1 : 0](ref, gmv_x, gmv_y, s_ref, s_pred, pred, w, h, bd, xeve_tbl_mc_c_coeff);
}
/* whole-block forward transform exactly as the reference sequences its two stages
 * (src_base/xeve_tq.c:396-404), with a selectable stage table */
RH_API void rh_fwd_transform(int variant, s16 *coef, int log2w, int log2h, int bd)
{
    const XEVE_TXB *t = variant == 2 ? xeve_tbl_txb_avx : xeve_tbl_txb;
    s32 *tb = malloc(sizeof(s32) * MAX_TR_DIM);
    int  s1 = xeve_get_transform_shift(log2w, 0, bd), s2 = xeve_get_transform_shift(log2h, 1, bd);
    t[log2w - 1](coef, tb, 0, 1 << log2h, 0);
    t[log2h - 1](tb, coef, s1 + s2, 1 << log2w, 1);
    free(tb);
}
RH_API void rh_inv_transform(int variant, s16 *coef, int log2w, int log2h, int bd)
{
    const XEVE_ITXB *t = variant == 2 ? xeve_tbl_itxb_avx : variant == 1 ? xeve_tbl_itxb_sse : xeve_tbl_itxb;
    s32 *tb = malloc(sizeof(s32) * MAX_TR_DIM);
    t[log2h - 1](coef, tb, 0, 1 << log2w, 0);
    t[log2w - 1](tb, coef, ITX_SHIFT1 + ITX_SHIFT2(bd), 1 << log2h, 1);
    free(tb);
}
RH_API void rh_recon(s16 *coef, s16 *pred, int is_coef, int w, int h, int s_rec, s16 *rec, int bd)
{
    xeve_recon_blk(coef, pred, is_coef, w, h, s_rec, rec, bd);
}
RH_API void rh_average(s16 *a, s16 *b, s16 *d, int w, int h) { xeve_average_16b_no_clip(a, b, d, w, w, w, w, h); }

/* constant tables (for checking the restatement's generated tables) */
RH_API const void *rh_table(int which, int *bytes)
{
    switch(which) {
    case 0: *bytes = sizeof(xeve_tbl_tm64); return xeve_tbl_tm64;
    case 1: *bytes = sizeof(xeve_tbl_tm32); return xeve_tbl_tm32;
    case 2: *bytes = sizeof(xeve_tbl_tm16); return xeve_tbl_tm16;
    case 3: *bytes = sizeof(xeve_tbl_tm8); return xeve_tbl_tm8;
    case 4: *bytes = sizeof(xeve_tbl_tm4); return xeve_tbl_tm4;
    case 5: *bytes = sizeof(xeve_tbl_tm2); return xeve_tbl_tm2;
    case 6: *bytes = 4096; return xeve_tbl_mv_bits - 2047;
    case 7: *bytes = sizeof(xeve_tbl_refi_bits); return xeve_tbl_refi_bits;
    case 8: *bytes = sizeof(xeve_tbl_scan); return xeve_tbl_scan;
    case 9: *bytes = sizeof(xeve_tbl_dq_scale_b); return xeve_tbl_dq_scale_b;
    case 10: *bytes = sizeof(xeve_quant_scale); return xeve_quant_scale;
    case 11: *bytes = sizeof(xeve_tbl_mc_l_coeff); return xeve_tbl_mc_l_coeff;
    case 12: *bytes = sizeof(xeve_tbl_mc_c_coeff); return xeve_tbl_mc_c_coeff;
    case 13: *bytes = sizeof(util_ctx()->err_scale); return util_ctx()->err_scale;
    case 14: *bytes = sizeof(xeve_tbl_df_st); return xeve_tbl_df_st;
    case 15: *bytes = sizeof(xeve_tbl_mpm); return xeve_tbl_mpm;
    }
    *bytes = 0;
    return NULL;
}

/* ------------------------------------------------------------------------------------------
 * replay: run the reference's own functions over a work list on `nthreads` host threads
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int              tid, nthreads, n;
    const RH_CONST  *cst;
    const RH_PLANES *planes;
    const s16       *side;
    /* ME */
    const RH_ME_REC *me_in;
    RH_ME_REC       *me_out;
    /* MC */
    const RH_MC_REC *mc_in;
    s16             *mc_pred; /* per item: w*h*3/2 at offsets mc_off[i] */
    const int64_t   *mc_off;
    uint64_t        *mc_hash;
    /* TQ */
    const RH_TQ_REC *tq_in;
    const RH_RATES  *rates;
    s16             *tq_coef;  /* output, same offsets as in_off - base */
    s16             *tq_resi;  /* optional: dequant + inverse transform output */
    s16             *tq_rec;   /* unused */
    int32_t         *tq_nnz;   /* n*3 */
    int              do_itdq;
} job_t;

static void fill_pic(XEVE_PIC *p, const RH_PLANES *pl)
{
    memset(p, 0, sizeof(*p));
    p->y = pl->y; p->u = pl->u; p->v = pl->v; p->s_l = pl->s_l; p->s_c = pl->s_c;
    p->w_l = pl->w_l; p->h_l = pl->h_l; p->w_c = pl->w_l / 2; p->h_c = pl->h_l / 2; p->poc = pl->poc;
}

static void *me_worker(void *arg)
{
    job_t       *j   = arg;
    XEVE_CTX    *u   = util_ctx();
    XEVE_PINTER *pi  = calloc(1, sizeof(XEVE_PINTER));
    XEVE_REFP (*refp)[REFP_NUM] = calloc(XEVE_MAX_NUM_REF_PICS, sizeof(XEVE_REFP[REFP_NUM]));
    XEVE_PIC     cur, ref;
    XEVE_PINTER *tp  = &u->pinter[0];
    pi->fn_me = (T.org_me && tp->fn_me == hook_me) ? T.org_me : tp->fn_me;
    pi->search_pattern_hpel = tp->search_pattern_hpel; pi->search_pattern_qpel = tp->search_pattern_qpel;
    pi->mc_l_coeff = xeve_tbl_mc_l_coeff; pi->mc_c_coeff = xeve_tbl_mc_c_coeff;
    pi->search_pattern_hpel_cnt = j->cst->hpel_cnt; pi->search_pattern_qpel_cnt = j->cst->qpel_cnt;
    pi->me_level = j->cst->me_level; pi->me_complexity = j->cst->me_complexity;
    pi->min_clip[0] = j->cst->min_clip[0]; pi->min_clip[1] = j->cst->min_clip[1];
    pi->max_clip[0] = j->cst->max_clip[0]; pi->max_clip[1] = j->cst->max_clip[1];
    pi->refp = refp;
    for(int i = j->tid; i < j->n; i += j->nthreads) {
        const RH_ME_REC *r = &j->me_in[i];
        RH_ME_REC       *o = &j->me_out[i];
        fill_pic(&cur, &j->planes[r->cur_pic]);
        fill_pic(&ref, &j->planes[r->ref_pic]);
        pi->pic_o = &cur; pi->o[0] = cur.y; pi->o[1] = cur.u; pi->o[2] = cur.v;
        pi->s_o[0] = cur.s_l; pi->s_o[1] = pi->s_o[2] = cur.s_c;
        refp[r->refi][r->lidx].pic = &ref; refp[r->refi][r->lidx].poc = r->ref_poc;
        pi->poc = r->poc; pi->gop_size = r->gop_size; pi->max_search_range = r->max_search_range;
        pi->lambda_mv = r->lambda_mv; pi->num_refp = r->num_refp;
        pi->mot_bits[0] = r->mot_bits_in[0]; pi->mot_bits[1] = r->mot_bits_in[1];
        if(r->bi) memcpy(pi->org_bi, j->side + r->org_bi_off, sizeof(s16) << (r->log2w + r->log2h));
        s8  refi = r->refi;
        s16 mvp[2] = {r->mvp[0], r->mvp[1]}, mv[2] = {r->mv_in[0], r->mv_in[1]};
        *o = *r;
        o->cost = pi->fn_me(pi, r->x, r->y, r->log2w, r->log2h, &refi, r->lidx, mvp, mv, r->bi, j->cst->bit_depth);
        o->mv_out[0] = mv[0]; o->mv_out[1] = mv[1];
        o->mot_bits_out[0] = pi->mot_bits[0]; o->mot_bits_out[1] = pi->mot_bits[1];
    }
    free(refp);
    free(pi);
    return NULL;
}

static double run_jobs(job_t *proto, void *(*fn)(void *), int nthreads)
{
    if(nthreads < 1) nthreads = 1;
    util_ctx();
    pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
    job_t     *jb = malloc(sizeof(job_t) * nthreads);
    double     t0 = now_s();
    for(int t = 0; t < nthreads; t++) {
        jb[t] = *proto; jb[t].tid = t; jb[t].nthreads = nthreads;
        pthread_create(&th[t], NULL, fn, &jb[t]);
    }
    for(int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    double dt = now_s() - t0;
    free(th);
    free(jb);
    return dt;
}

/* returns wall seconds */
RH_API double rh_replay_me(const RH_CONST *cst, const RH_PLANES *planes, const s16 *side, const RH_ME_REC *in,
                           RH_ME_REC *out, int n, int nthreads)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.cst = cst; j.planes = planes; j.side = side; j.me_in = in; j.me_out = out; j.n = n;
    return run_jobs(&j, me_worker, nthreads);
}

static void *mc_worker(void *arg)
{
    job_t *j = arg;
    XEVE_REFP (*refp)[REFP_NUM] = calloc(XEVE_MAX_NUM_REF_PICS, sizeof(XEVE_REFP[REFP_NUM]));
    pel (*pred)[N_C][MAX_CU_DIM] = malloc(sizeof(pel) * 2 * N_C * MAX_CU_DIM);
    XEVE_PIC pic[2];
    for(int i = j->tid; i < j->n; i += j->nthreads) {
        const RH_MC_REC *r = &j->mc_in[i];
        s8  refi[2] = {r->refi[0], r->refi[1]};
        s16 mv[2][2] = {{r->mv[0][0], r->mv[0][1]}, {r->mv[1][0], r->mv[1][1]}};
        for(int l = 0; l < 2; l++)
            if(REFI_IS_VALID(refi[l])) {
                fill_pic(&pic[l], &j->planes[r->ref_pic[l]]);
                refp[refi[l]][l].pic = &pic[l]; refp[refi[l]][l].poc = pic[l].poc;
            }
        xeve_mc(r->x, r->y, j->cst->w, j->cst->h, r->w, r->h, refi, mv, refp, pred, j->cst->bit_depth,
                j->cst->bit_depth, 1);
        size_t ny = (size_t)r->w * r->h, nc = ny / 4;
        if(j->mc_pred) {
            s16 *d = j->mc_pred + j->mc_off[i];
            memcpy(d, pred[0][Y_C], ny * 2); memcpy(d + ny, pred[0][U_C], nc * 2); memcpy(d + ny + nc, pred[0][V_C], nc * 2);
        }
        if(j->mc_hash) {
            uint64_t h = FNV_INIT;
            h = fnv1a(h, pred[0][Y_C], ny * 2); h = fnv1a(h, pred[0][U_C], nc * 2); h = fnv1a(h, pred[0][V_C], nc * 2);
            j->mc_hash[i] = h;
        }
    }
    free(pred);
    free(refp);
    return NULL;
}

RH_API double rh_replay_mc(const RH_CONST *cst, const RH_PLANES *planes, const RH_MC_REC *in, int n, s16 *pred_out,
                           const int64_t *pred_off, uint64_t *hash_out, int nthreads)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.cst = cst; j.planes = planes; j.mc_in = in; j.n = n; j.mc_pred = pred_out; j.mc_off = pred_off; j.mc_hash = hash_out;
    return run_jobs(&j, mc_worker, nthreads);
}

static void *tq_worker(void *arg)
{
    job_t     *j = arg;
    XEVE_CTX  *u = util_ctx();
    XEVE_CORE *core = calloc(1, sizeof(XEVE_CORE));
    s16 (*coef)[MAX_CU_DIM] = malloc(sizeof(s16) * N_C * MAX_CU_DIM);
    int (*org_tq)(XEVE_CTX *, XEVE_CORE *, s16 (*)[MAX_CU_DIM], int, int, int, int *, int, int) =
        (u->fn_tq == hook_tq) ? T.org_tq : u->fn_tq;
    core->ctx = u;
    int last_rate = -1;
    for(int i = j->tid; i < j->n; i += j->nthreads) {
        const RH_TQ_REC *r = &j->tq_in[i];
        size_t ny = (size_t)1 << (r->log2w + r->log2h), nc = ny >> 2;
        const s16 *src = j->side + r->in_off;
        memcpy(coef[Y_C], src, ny * 2); memcpy(coef[U_C], src + ny, nc * 2); memcpy(coef[V_C], src + ny + nc, nc * 2);
        core->qp_y = r->qp[0]; core->qp_u = r->qp[1]; core->qp_v = r->qp[2];
        for(int c = 0; c < 3; c++) core->lambda[c] = r->lambda[c];
        core->log2_cuw = r->log2w; core->log2_cuh = r->log2h;
        if(r->rate_idx != last_rate) { put_rates(core, &j->rates[r->rate_idx]); last_rate = r->rate_idx; }
        int nnz[3];
        org_tq(u, core, coef, r->log2w, r->log2h, r->slice_type, nnz, r->is_intra, r->run_stats);
        s16 *d = j->tq_coef + r->in_off;
        memcpy(d, coef[Y_C], ny * 2); memcpy(d + ny, coef[U_C], nc * 2); memcpy(d + ny + nc, coef[V_C], nc * 2);
        for(int c = 0; c < 3; c++) j->tq_nnz[i * 3 + c] = nnz[c];
        if(j->do_itdq && j->tq_resi) {
            u->fn_itdp(u, core, coef, core->nnz_sub);
            s16 *e = j->tq_resi + r->in_off;
            memcpy(e, coef[Y_C], ny * 2); memcpy(e + ny, coef[U_C], nc * 2); memcpy(e + ny + nc, coef[V_C], nc * 2);
        }
    }
    free(coef);
    free(core);
    return NULL;
}

/* side: the input planes; coef_out / resi_out use the same element offsets (in_off) */
RH_API double rh_replay_tq(const RH_CONST *cst, const s16 *side, const RH_TQ_REC *in, const RH_RATES *rates, int n,
                           s16 *coef_out, int32_t *nnz_out, s16 *resi_out, int nthreads)
{
    job_t j;
    memset(&j, 0, sizeof(j));
    j.cst = cst; j.side = side; j.tq_in = in; j.rates = rates; j.n = n; j.tq_coef = coef_out; j.tq_nnz = nnz_out;
    j.tq_resi = resi_out; j.do_itdq = resi_out != NULL;
    util_ctx()->param.rdoq = cst->rdoq; /* the utility context follows the traced encode's quantiser choice */
    return run_jobs(&j, tq_worker, nthreads);
}

/* ------------------------------------------------------------------------------------------
 * replay of the distortion/transform body of pinter_residue_rdo (src_base/xeve_pinter.c:961-1056)
 * with the reference's own functions: xeve_mc -> xeve_func_diff -> xeve_func_ssd -> ctx->fn_tq ->
 * ctx->fn_itdp -> xeve_recon_blk -> xeve_func_ssd.  Record layout = xb200_residue_item.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    RH_MC_REC mc;
    int32_t   cur_pic;
    uint8_t   slice_type, run_stats, qp[3], pad_[3];
    int32_t   rate_idx;
    double    lambda[3];
    int64_t   out_off;
    int32_t   nnz[3];
    int64_t   dist_pred[3];
    int64_t   dist_rec[3];
} RH_RES_REC;

typedef struct {
    int              tid, nthreads, n;
    const RH_CONST  *cst;
    const RH_PLANES *planes;
    const RH_RATES  *rates;
    RH_RES_REC      *items;
    s16             *coef, *rec;
} resjob_t;

static void *res_worker(void *arg)
{
    resjob_t  *j = arg;
    XEVE_CTX  *u = util_ctx();
    XEVE_CORE *core = calloc(1, sizeof(XEVE_CORE));
    XEVE_REFP (*refp)[REFP_NUM] = calloc(XEVE_MAX_NUM_REF_PICS, sizeof(XEVE_REFP[REFP_NUM]));
    pel (*pred)[N_C][MAX_CU_DIM] = malloc(sizeof(pel) * 2 * N_C * MAX_CU_DIM);
    s16 (*coef)[MAX_CU_DIM] = malloc(sizeof(s16) * N_C * MAX_CU_DIM);
    s16 (*resi)[MAX_CU_DIM] = malloc(sizeof(s16) * N_C * MAX_CU_DIM);
    XEVE_PIC pic[2];
    int (*org_tq)(XEVE_CTX *, XEVE_CORE *, s16 (*)[MAX_CU_DIM], int, int, int, int *, int, int) =
        (u->fn_tq == hook_tq) ? T.org_tq : u->fn_tq;
    core->ctx = u;
    int last_rate = -1, bd = j->cst->bit_depth;
    for(int i = j->tid; i < j->n; i += j->nthreads) {
        RH_RES_REC *r = &j->items[i];
        s8  refi[2] = {r->mc.refi[0], r->mc.refi[1]};
        s16 mv[2][2] = {{r->mc.mv[0][0], r->mc.mv[0][1]}, {r->mc.mv[1][0], r->mc.mv[1][1]}};
        for(int l = 0; l < 2; l++)
            if(REFI_IS_VALID(refi[l])) {
                fill_pic(&pic[l], &j->planes[r->mc.ref_pic[l]]);
                refp[refi[l]][l].pic = &pic[l]; refp[refi[l]][l].poc = pic[l].poc;
            }
        const int w = r->mc.w, h = r->mc.h, x = r->mc.x, y = r->mc.y;
        int l2 = 0; while((1 << l2) < w) l2++;
        xeve_mc(x, y, j->cst->w, j->cst->h, w, h, refi, mv, refp, pred, bd, bd, 1);
        const RH_PLANES *o = &j->planes[r->cur_pic];
        s16 *org[3] = {o->y + y * o->s_l + x, o->u + (y >> 1) * o->s_c + (x >> 1), o->v + (y >> 1) * o->s_c + (x >> 1)};
        int  so[3] = {o->s_l, o->s_c, o->s_c}, lw[3] = {l2, l2 - 1, l2 - 1};
        for(int c = 0; c < 3; c++) {
            xeve_diff_16b(lw[c], lw[c], org[c], pred[0][c], so[c], 1 << lw[c], 1 << lw[c], coef[c], bd);
            r->dist_pred[c] = xeve_ssd_16b(lw[c], lw[c], pred[0][c], org[c], 1 << lw[c], so[c], bd);
        }
        core->qp_y = r->qp[0]; core->qp_u = r->qp[1]; core->qp_v = r->qp[2];
        for(int c = 0; c < 3; c++) core->lambda[c] = r->lambda[c];
        core->log2_cuw = l2; core->log2_cuh = l2;
        if(r->rate_idx != last_rate) { put_rates(core, &j->rates[r->rate_idx]); last_rate = r->rate_idx; }
        int nnz[3];
        org_tq(u, core, coef, l2, l2, r->slice_type, nnz, 0, r->run_stats);
        size_t ny = (size_t)w * h, nc = ny >> 2, off[3] = {0, ny, ny + nc};
        for(int c = 0; c < 3; c++) {
            memcpy(j->coef + r->out_off + off[c], coef[c], (c ? nc : ny) * 2);
            memcpy(resi[c], coef[c], (c ? nc : ny) * 2);
        }
        u->fn_itdp(u, core, resi, core->nnz_sub);
        for(int c = 0; c < 3; c++) {
            s16 *rc = j->rec + r->out_off + off[c];
            xeve_recon_blk(resi[c], pred[0][c], nnz[c], 1 << lw[c], 1 << lw[c], 1 << lw[c], rc, bd);
            r->nnz[c] = nnz[c];
            r->dist_rec[c] = nnz[c] ? xeve_ssd_16b(lw[c], lw[c], rc, org[c], 1 << lw[c], so[c], bd) : r->dist_pred[c];
        }
    }
    free(resi); free(coef); free(pred); free(refp); free(core);
    return NULL;
}

RH_API double rh_replay_residue(const RH_CONST *cst, const RH_PLANES *planes, const RH_RATES *rates, RH_RES_REC *items, int n,
                                s16 *coef, s16 *rec, int nthreads)
{
    if(nthreads < 1) nthreads = 1;
    util_ctx()->param.rdoq = cst->rdoq;
    pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
    resjob_t  *jb = malloc(sizeof(resjob_t) * nthreads);
    double     t0 = now_s();
    for(int t = 0; t < nthreads; t++) {
        jb[t] = (resjob_t){t, nthreads, n, cst, planes, rates, items, coef, rec};
        pthread_create(&th[t], NULL, res_worker, &jb[t]);
    }
    for(int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    double dt = now_s() - t0;
    free(th); free(jb);
    return dt;
}
RH_API int rh_sizeof_res(void) { return sizeof(RH_RES_REC); }

/* ------------------------------------------------------------------------------------------
 * MV-predictor probe: the reference's xeve_get_avail_inter / xeve_get_motion / xeve_get_mv_dir on
 * caller-provided SCU maps (record layout = xb200_mvp_item / xb200_mvp_pic)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int16_t  x_scu, y_scu;
    uint8_t  log2_cuw, log2_cuh, lidx, pad_;
    uint16_t avail;
    int8_t   refi[4];
    int16_t  mvp[4][2];
    int16_t  mv_dir[2][2];
} RH_MVP_REC;
typedef struct { int32_t w_scu, h_scu, poc, ref_poc[2], col_list_poc0; } RH_MVP_PIC;

RH_API void rh_mvp(RH_MVP_REC *items, int n, const RH_MVP_PIC *pp, u32 *map_scu, s16 (*map_mv)[REFP_NUM][MV_D],
                   s16 (*col0)[REFP_NUM][MV_D], s16 (*col1)[REFP_NUM][MV_D])
{
    XEVE_REFP refp[XEVE_MAX_NUM_REF_PICS][REFP_NUM];
    memset(refp, 0, sizeof(refp));
    u32 list_poc[XEVE_MAX_NUM_REF_PICS] = {0};
    list_poc[0] = pp->col_list_poc0;
    refp[0][REFP_0].map_mv = col0; refp[0][REFP_0].poc = pp->ref_poc[0];
    refp[0][REFP_1].map_mv = col1; refp[0][REFP_1].poc = pp->ref_poc[1]; refp[0][REFP_1].list_poc = list_poc;
    refp[0][REFP_0].list_poc = list_poc;
    u8 *tidx = calloc((size_t)pp->w_scu * pp->h_scu, 1);
    for(int i = 0; i < n; i++) {
        RH_MVP_REC *r = &items[i];
        int cuw = 1 << r->log2_cuw, cuh = 1 << r->log2_cuh, scup = r->x_scu + r->y_scu * pp->w_scu;
        r->avail = xeve_get_avail_inter(r->x_scu, r->y_scu, pp->w_scu, pp->h_scu, scup, cuw, cuh, map_scu, tidx);
        xeve_get_motion(scup, r->lidx, NULL, map_mv, refp, cuw, cuh, pp->w_scu, r->avail, r->refi, r->mvp);
        xeve_get_mv_dir(refp[0], pp->poc, scup + ((cuw >> 2) - 1) + ((cuh >> 2) - 1) * pp->w_scu, scup, pp->w_scu, pp->h_scu,
                        r->mv_dir, 0);
    }
    free(tidx);
}
RH_API int rh_sizeof_mvp(void) { return sizeof(RH_MVP_REC); }

/* ------------------------------------------------------------------------------------------
 * RDO bit counting probes: the reference's xeve_rdo_bit_cnt_* on a scratch core, exactly the
 * SBAC_LOAD -> xeve_sbac_bit_reset -> count -> xeve_get_bit_number sequence of xeve_pinter.c
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint8_t  kind, slice_type, log2_cuw, log2_cuh, pidx, ch, ctx_skip, ctx_pred_mode;
    int8_t   refi[2];
    uint8_t  mvp_idx[2], num_refp[2], all_preds, pad_;
    int16_t  mvd[2][2];
    int32_t  nnz[3], state_in, state_out;
    int64_t  coef_off;
    uint32_t bits, pad2_;
} RH_BITS_REC;

static void sbac_unpack(const RH_SBAC *s, XEVE_SBAC *d)
{
    XEVE_SBAC_CTX *c = &d->ctx;
    d->range = s->range;
    memcpy(c->skip_flag, s->m + 0, 4); memcpy(c->pred_mode, s->m + 2, 6); memcpy(c->direct_mode_flag, s->m + 5, 2);
    memcpy(c->inter_dir, s->m + 6, 4); memcpy(c->refi, s->m + 8, 4); memcpy(c->mvp_idx, s->m + 10, 6);
    memcpy(c->mvd, s->m + 13, 2); memcpy(c->cbf_all, s->m + 14, 2); memcpy(c->cbf_luma, s->m + 15, 2);
    memcpy(c->cbf_cb, s->m + 16, 2); memcpy(c->cbf_cr, s->m + 17, 2); memcpy(c->run, s->m + 18, 48);
    memcpy(c->last, s->m + 42, 4); memcpy(c->level, s->m + 44, 48);
}
static void sbac_pack(const XEVE_SBAC *d, RH_SBAC *s)
{
    const XEVE_SBAC_CTX *c = &d->ctx;
    s->range = d->range;
    memcpy(s->m + 0, c->skip_flag, 4); memcpy(s->m + 2, c->pred_mode, 6); memcpy(s->m + 5, c->direct_mode_flag, 2);
    memcpy(s->m + 6, c->inter_dir, 4); memcpy(s->m + 8, c->refi, 4); memcpy(s->m + 10, c->mvp_idx, 6);
    memcpy(s->m + 13, c->mvd, 2); memcpy(s->m + 14, c->cbf_all, 2); memcpy(s->m + 15, c->cbf_luma, 2);
    memcpy(s->m + 16, c->cbf_cb, 2); memcpy(s->m + 17, c->cbf_cr, 2); memcpy(s->m + 18, c->run, 48);
    memcpy(s->m + 42, c->last, 4); memcpy(s->m + 44, c->level, 48);
}

RH_API int rh_rdo_bits(RH_BITS_REC *items, int n, RH_SBAC *states, const s16 *coef_buf)
{
    XEVE_CTX  *ctx = util_ctx();
    XEVE_CORE *core = calloc(1, sizeof(XEVE_CORE));
    s16(*coef)[MAX_CU_DIM] = calloc(N_C, sizeof(*coef));
    if(!ctx || !core || !coef) return -1;
    ctx->pps.cu_qp_delta_enabled_flag = 0;
    ctx->sps.tool_admvp = 0;
    ctx->sps.chroma_format_idc = 1;
    core->bs_temp.pdata[1] = &core->s_temp_run;
    for(int i = 0; i < n; i++) {
        RH_BITS_REC *it = &items[i];
        memset(&core->s_temp_run, 0, sizeof(core->s_temp_run));
        /* an arbitrary legal low/pending state: the count must not depend on it */
        core->s_temp_run.code = 0x12345u + 977u * (u32)i;
        core->s_temp_run.code_bits = 1 + i % 11;
        sbac_unpack(&states[it->state_in], &core->s_temp_run);
        core->s_temp_run.is_bitcount = 1;
        xeve_sbac_bit_reset(&core->s_temp_run);
        core->log2_cuw = it->log2_cuw; core->log2_cuh = it->log2_cuh;
        core->cuw = 1 << it->log2_cuw; core->cuh = 1 << it->log2_cuh;
        core->ctx_flags[CNID_SKIP_FLAG] = it->ctx_skip; core->ctx_flags[CNID_PRED_MODE] = it->ctx_pred_mode;
        core->tree_cons = xeve_get_default_tree_cons();
        if(!it->all_preds) core->tree_cons.mode_cons = eOnlyInter;
        memset(core->nnz_sub, 0, sizeof(core->nnz_sub));
        for(int c = 0; c < 3; c++) core->nnz[c] = core->nnz_sub[c][0] = it->nnz[c];
        ctx->rpm.num_refp[0] = it->num_refp[0]; ctx->rpm.num_refp[1] = it->num_refp[1];
        if(it->kind == 1 || it->kind == 3) {
            int ny = 1 << (it->log2_cuw + it->log2_cuh), nc = ny >> 2;
            memcpy(coef[0], coef_buf + it->coef_off, ny * 2);
            memcpy(coef[1], coef_buf + it->coef_off + ny, nc * 2);
            memcpy(coef[2], coef_buf + it->coef_off + ny + nc, nc * 2);
        }
        switch(it->kind) {
        case 0: xeve_rdo_bit_cnt_cu_skip(ctx, core, it->slice_type, 0, it->mvp_idx[0], it->mvp_idx[1], 0, 0); break;
        case 1: xeve_rdo_bit_cnt_cu_inter(ctx, core, it->slice_type, 0, it->refi, it->mvd, coef, it->pidx, it->mvp_idx, 0, 0, NULL); break;
        case 2: xeve_rdo_bit_cnt_mvp(ctx, core, it->slice_type, it->refi, it->mvd, it->pidx, it->mvp_idx[0]); break;
        default: xeve_rdo_bit_cnt_cu_inter_comp(core, coef, it->ch, it->pidx, ctx, core->tree_cons); break;
        }
        it->bits = xeve_get_bit_number(&core->s_temp_run);
        if(it->state_out >= 0) sbac_pack(&core->s_temp_run, &states[it->state_out]);
    }
    free(coef);
    free(core);
    return 0;
}
RH_API int rh_sizeof_bits(void) { return sizeof(RH_BITS_REC); }
