/*
 * xeve_b200.h -- C ABI of the B200-native XEVE inter-search + transform hot path.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no torch or
 * CUDA types.  Each entry point replaces one of the reference's internal operator hooks for the
 * path xeve_pinter -> xeve_sad / xeve_mc -> xeve_tq / xeve_itdq, in *batched* form: instead of one
 * synchronous call per block, the caller hands over a work list (one record per call the
 * reference would have made) and gets every result back.  Record fields carry exactly the
 * arguments and the XEVE_PINTER / XEVE_CORE fields the reference hook reads.
 *
 *   entry point              replaces (reference file:line)
 *   -----------------------  ---------------------------------------------------------------
 *   xb200_pic_upload         xeve_imgb_cpy bit-depth convert (src_base/xeve_util.c:1552-1704)
 *                            + xeve_picbuf_expand padding     (src_base/xeve_util.c:190-248)
 *   xb200_sad / xb200_ssd    xeve_func_sad/ssd[log2w][log2h] (src_base/xeve_sad.h:40-68); xeve_func_diff
 *                            is fused into xb200_residue
 *   xb200_satd               xeve_func_satd[0] = xeve_had     (src_base/xeve_sad.c:1043-1140)
 *   xb200_me                 pi->fn_me = pinter_me_epzs        (src_base/xeve_pinter.c:699-869,
 *                            hook declared src_base/xeve_type.h:448)
 *   xb200_mc                 pi->fn_mc -> xeve_mc              (src_base/xeve_mc.c:465-610,
 *                            hook declared src_base/xeve_type.h:452)
 *   xb200_bi_org             get_org_bi after fn_mc            (src_base/xeve_pinter.c:143-156, 1620-1622)
 *   xb200_tq                 ctx->fn_tq = xeve_sub_block_tq    (src_base/xeve_tq.c:750-864,
 *                            hook declared src_base/xeve_type.h:982)
 *   xb200_itdq               ctx->fn_itdp = xeve_itdq          (src_base/xeve_itdq.c:499-580)
 *   xb200_recon              ctx->fn_recon -> xeve_recon_blk   (src_base/xeve_recon.c:35-57)
 *   xb200_residue            the distortion/transform body of pinter_residue_rdo
 *                            (src_base/xeve_pinter.c:961-1056): fn_mc -> xeve_diff_pred -> SSD ->
 *                            fn_tq -> fn_itdp -> fn_recon -> SSD, fused on the device
 *   xb200_analyze_intra      ctx->fn_pintra_analyze_cu = pintra_analyze_cu (src_base/xeve_pintra.c:544-698,
 *                            hook declared src_base/xeve_type.h:966) over xeve_ipred (src_base/xeve_ipred.c:99-228)
 *   xb200_deblock            ctx->fn_loop_filter = xeve_loop_filter (src_base/xeve_enc.c:2355-2414) ->
 *                            xeve_deblock / _cu_ver / _cu_hor   (src_base/xeve_df.c:34-573)
 *                            + ctx->fn_picbuf_expand            (src_base/xeve_enc.c:1274)
 *
 * Status codes are the reference's own (inc/xeve.h:50-74): XB200_OK == XEVE_OK == 0, errors
 * negative.  All functions are synchronous on return (results are in the caller's buffers).
 * `mem` selects where the caller's *bulk* buffers live: XB200_MEM_HOST (the FFI case: staged
 * through pinned memory, copies included in the call) or XB200_MEM_DEVICE (already in HBM).
 * The library never falls back to a CPU implementation: without a usable sm_100 device
 * xb200_create fails with XB200_ERR_UNSUPPORTED.
 */
#ifndef XEVE_B200_H_
#define XEVE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XB200_API __attribute__((visibility("default")))

/* status codes (values of inc/xeve.h:50-74) */
#define XB200_OK                   (0)
#define XB200_ERR                  (-1)
#define XB200_ERR_INVALID_ARGUMENT (-101)
#define XB200_ERR_OUT_OF_MEMORY    (-102)
#define XB200_ERR_UNSUPPORTED      (-104)
#define XB200_ERR_UNEXPECTED       (-105)

#define XB200_MEM_HOST   0
#define XB200_MEM_DEVICE 1

#define XB200_PAD_L 144 /* PIC_PAD_SIZE_L, src_base/xeve_def.h:380 */
#define XB200_PAD_C 72  /* PIC_PAD_SIZE_C, src_base/xeve_def.h:381 */

#define XB200_NUM_CTX_CC_RUN   24 /* src_base/xeve_def.h:722-724 */
#define XB200_NUM_CTX_CC_LEVEL 24
#define XB200_NUM_CTX_CC_LAST  2

typedef struct xb200_ctx xb200_ctx; /* opaque */

/* Sequence-level constants of the search: the XEVE_PINTER / XEVE_PARAM fields that stay fixed
 * for an encoder instance (src_base/xeve_pinter.c:2087-2133). */
typedef struct {
    int32_t w, h;            /* picture size in luma samples (param.w/h aligned up to 8) */
    int32_t bit_depth;       /* internal bit depth (10; src_base/xeve_enc.c:2298) */
    int32_t me_level;        /* param.me_sub: 1 int-pel, 2 half-pel, 3 quarter-pel */
    int32_t hpel_cnt;        /* param.me_sub_pos */
    int32_t qpel_cnt;        /* param.me_sub_pos */
    int32_t me_complexity;   /* param.me_algo */
    int32_t min_clip[2];     /* -MAX_CU_SIZE + 1 */
    int32_t max_clip[2];     /* param.w - 1, param.h - 1 */
    int32_t merge_num, me_range, gop_size, rdoq, tool_iqt;
} xb200_seq;

/* One pi->fn_me call (in/out record). */
typedef struct {
    int32_t  poc;            /* pi->poc */
    int32_t  cur_pic;        /* picture handle of the original (pi->o) */
    int32_t  ref_pic;        /* picture handle of pi->refp[refi][lidx].pic */
    int32_t  ref_poc;        /* pi->refp[refi][lidx].poc */
    int16_t  x, y;
    uint8_t  log2_cuw, log2_cuh, lidx, bi;
    int8_t   refi;
    uint8_t  num_refp;       /* pi->num_refp */
    int16_t  mvp[2];
    int16_t  mv_in[2];       /* mv[] on entry (start MV when bi != 0) */
    uint32_t lambda_mv;      /* pi->lambda_mv */
    int32_t  mot_bits_in[2]; /* pi->mot_bits[] on entry */
    int32_t  max_search_range; /* pi->max_search_range */
    int32_t  gop_size;       /* pi->gop_size */
    int32_t  org_bi_off;     /* bi != 0: element offset of pi->org_bi (cuw*cuh s16) in `side`; else -1 */
    /* results */
    int16_t  mv_out[2];      /* mv[] on return (quarter-pel, relative to the CU) */
    uint32_t cost;           /* return value */
    int32_t  mot_bits_out[2];/* pi->mot_bits[] on return */
} xb200_me_item;

/* One pi->fn_mc call.  The prediction of item i (Y: w*h, then U, V: w*h/4 each, s16, stride = block
 * width, exactly pred[0][Y_C..V_C]) is written at element offset pred_off[i] of the output. */
typedef struct {
    int32_t  poc;
    int32_t  ref_pic[2];     /* picture handle per list, -1 when refi[l] is invalid */
    int32_t  ref_poc[2];
    int16_t  x, y, w, h;
    int8_t   refi[2];
    int16_t  mv[2][2];
    uint64_t out_hash;       /* unused by the library (test bookkeeping) */
} xb200_mc_item;

/* CABAC-derived rate tables RDOQ reads (XEVE_CORE::rdoq_est_*, src_base/xeve_type.h:737-747). */
typedef struct {
    int32_t cbf_all[2], cbf_luma[2], cbf_cb[2], cbf_cr[2];
    int32_t run[XB200_NUM_CTX_CC_RUN][2];
    int32_t level[XB200_NUM_CTX_CC_LEVEL][2];
    int32_t last[XB200_NUM_CTX_CC_LAST][2];
} xb200_rates;

/* One ctx->fn_tq call.  Input/output planes of item i live at element offset in_off of the
 * coefficient buffer: Y (cuw*cuh) then U, V ((cuw/2)*(cuh/2) each), s16, in place. */
typedef struct {
    int32_t  poc;
    uint8_t  log2_cuw, log2_cuh, slice_type, is_intra;
    uint8_t  run_stats, qp[3];     /* core->qp_y/u/v (already + 6*(bd-8)) */
    int32_t  rate_idx;             /* index into the xb200_rates array */
    int64_t  in_off;
    double   lambda[3];            /* core->lambda[] */
    int32_t  nnz[3];               /* result */
    uint64_t out_hash;             /* unused by the library */
} xb200_tq_item;

/* Fused residue work item = the distortion/transform body of pinter_residue_rdo for one
 * candidate mode of one CU.  Outputs are written to per-item slots of size 3/2*cuw*cuh at
 * element offset out_off in each of the coef / rec buffers.
 * Preconditions (checked on the device for host-buffer calls, XB200_ERR_INVALID_ARGUMENT otherwise): the CU (mc.x, mc.y, mc.w, mc.h)
 * lies inside the current picture, out_off is even and out_off + 3/2 w h fits the buffers, rate_idx < n_rates, valid picture handles
 * (reference pictures padded).  xb200_me likewise requires the CU inside the current picture and 1 <= max_search_range <= 256. */
typedef struct {
    xb200_mc_item mc;              /* prediction to build (x, y, w, h, refi, mv, refs) */
    int32_t  cur_pic;              /* picture handle of the original picture */
    uint8_t  slice_type, run_stats, qp[3], pad_[3];
    int32_t  rate_idx;
    double   lambda[3];
    int64_t  out_off;
    /* results */
    int32_t  nnz[3];
    int64_t  dist_pred[3];         /* SSD(pred, org)  -- dist[0][] / dist_no_resi[] */
    int64_t  dist_rec[3];          /* SSD(rec, org) where nnz != 0, else = dist_pred -- dist[1][] */
} xb200_residue_item;

/* Element-wise kernel probe (XEVE_FN_SAD / SSD / DIFF / SATD argument lists). Buffers are the
 * picture planes; a block is addressed by picture handle, plane and (x, y) which may reach
 * into the padding. */
typedef struct {
    int32_t pic1, pic2;
    int16_t x1, y1, x2, y2;
    uint8_t plane1, plane2, log2w, log2h;
} xb200_blk_item;

/* MV-predictor inputs of one CU: xeve_get_avail_inter (src_base/xeve_util.c:652-715,
 * single tile), xeve_get_motion (:526-573) for one list, and the temporal-direct MVs of xeve_get_mv_dir
 * (:619-650).  Maps are SCU-granular (4x4) frame maps as the reference keeps them (ctx->map_scu u32 flags,
 * ctx->map_mv [f_scu][2][2] s16, the colocated refp[0][lidx].map_mv). */
typedef struct {
    int16_t  x_scu, y_scu;
    uint8_t  log2_cuw, log2_cuh, lidx, pad_;
    /* results */
    uint16_t avail;          /* AVAIL_* bit mask (src_base/xeve_def.h:402-425) */
    int8_t   refi[4];
    int16_t  mvp[4][2];      /* left, up, up-right (or (1,1)), colocated */
    int16_t  mv_dir[2][2];   /* xeve_get_mv_dir for the SCU at the CU's bottom-right corner */
} xb200_mvp_item;

typedef struct {             /* picture-level inputs of xb200_mvp */
    int32_t w_scu, h_scu;
    int32_t poc;             /* current picture */
    int32_t ref_poc[2];      /* refp[0][REFP_0].poc, refp[0][REFP_1].poc */
    int32_t col_list_poc0;   /* refp[0][REFP_1].list_poc[0] */
} xb200_mvp_pic;

/* ---- device-side rate estimation: CABAC bit counting for inter RDO ------------------------------
 * The RDO bit count of the reference is the arithmetic coder run in "bitcount" mode after xeve_sbac_bit_reset
 * (src_base/xeve_mode.c:39-55); it equals the number of renormalisation shifts, so only the coder's range and
 * the context models matter.  xb200_sbac carries exactly those (models of the Baseline inter syntax, 2-byte
 * SBAC_CTX_MODEL = state << 1 | mps, src_base/xeve_def.h:727-790). */
enum {
    XB200_CM_SKIP_FLAG = 0, XB200_CM_PRED_MODE = 2, XB200_CM_DIRECT = 5, XB200_CM_INTER_DIR = 6, XB200_CM_REFI = 8,
    XB200_CM_MVP_IDX = 10, XB200_CM_MVD = 13, XB200_CM_CBF_ALL = 14, XB200_CM_CBF_LUMA = 15, XB200_CM_CBF_CB = 16,
    XB200_CM_CBF_CR = 17, XB200_CM_RUN = 18, XB200_CM_LAST = 42, XB200_CM_LEVEL = 44, XB200_CM_COUNT = 68
};
typedef struct {
    uint32_t range;                  /* XEVE_SBAC::range */
    uint16_t m[XB200_CM_COUNT];      /* context models, laid out by the XB200_CM_* offsets */
} xb200_sbac;

/* One RDO bit-count call: SBAC_LOAD(state) -> xeve_sbac_bit_reset -> syntax -> xeve_get_bit_number.
 * kind 0: xeve_rdo_bit_cnt_cu_skip (src_base/xeve_mode.c:283-302)      kind 1: xeve_rdo_bit_cnt_cu_inter (:185-281)
 * kind 2: xeve_rdo_bit_cnt_mvp (:57-77, mvp_idx[0] used for both lists)  kind 3: xeve_rdo_bit_cnt_cu_inter_comp (:160-183) */
typedef struct {
    uint8_t  kind, slice_type, log2_cuw, log2_cuh;
    uint8_t  pidx;                   /* PRED_L0 0, PRED_L1 1, PRED_BI 2, PRED_SKIP 3, PRED_DIR 4 */
    uint8_t  ch;                     /* kind 3: component */
    uint8_t  ctx_skip, ctx_pred_mode;/* core->ctx_flags[CNID_SKIP_FLAG], [CNID_PRED_MODE] */
    int8_t   refi[2];
    uint8_t  mvp_idx[2];
    uint8_t  num_refp[2];
    uint8_t  all_preds, pad_;        /* xeve_check_all_preds(core->tree_cons) */
    int16_t  mvd[2][2];
    int32_t  nnz[3];                 /* core->nnz_sub[c][0] */
    int32_t  state_in;               /* index of the input coder state */
    int32_t  state_out;              /* index where the coder state after the call is stored, -1: discard */
    int64_t  coef_off;               /* element offset of the coefficient planes (Y | U | V), kinds 1 and 3 */
    uint32_t bits;                   /* result: xeve_get_bit_number */
    uint32_t pad2_;
} xb200_bits_item;

/* One ctx->fn_pinter_analyze_cu call (src_base/xeve_pinter.c:1839-2056): best inter mode of one CU among SKIP, DIRECT (B),
 * L0, L1 and BI by RD cost.  The neighbourhood enters through the xeve_get_motion candidates (xb200_mvp) and the input
 * coder state; costs are IEEE doubles evaluated in the reference's operation order (no fused multiply-add). */
#define XB200_MAX_REFP 4
typedef struct {
    int32_t  poc, cur_pic;          /* POC and original-picture handle of the picture being coded */
    int16_t  x, y;
    uint8_t  log2_cuw, log2_cuh, slice_type, ctx_skip, ctx_pred_mode, all_preds;
    uint8_t  num_refp[2];           /* ctx->rpm.num_refp[] */
    uint8_t  qp[3], pad0_;          /* core->qp_y/u/v */
    int32_t  max_search_range;      /* pi->max_search_range (src_base/xeve_pinter.c:2093) */
    int32_t  ref_pic[2][XB200_MAX_REFP], ref_poc[2][XB200_MAX_REFP]; /* pi->refp[refi][lidx] as [lidx][refi] */
    uint32_t lambda_mv;             /* pi->lambda_mv */
    int32_t  rate_idx;              /* RDOQ rate tables of the input coder state (xb200_rdoq_rates) */
    int32_t  state_in, state_out;   /* coder state slots: core->s_curr_best[..] in, core->s_next_best[..] out */
    double   lambda[3], dist_chroma_weight[2];
    int16_t  mvp[2][4][2];          /* pi->mvp[lidx][] (== pi->mvp_scale[lidx][refi][]) */
    int8_t   refi_pred[2][4];       /* pi->refi_pred[lidx][] */
    int16_t  mv_dir[2][2];          /* xeve_get_mv_dir result (B slices) */
    int64_t  out_off;               /* element offset of this CU's coef / rec slots (3/2 * cuw * cuh each) */
    /* results */
    double   cost;                  /* return value */
    uint8_t  best_idx, pad1_;       /* PRED_L0 0, PRED_L1 1, PRED_BI 2, PRED_SKIP 3, PRED_DIR 4 */
    int8_t   refi[2];               /* XEVE_MODE::refi, mvp_idx, mv, mvd */
    uint8_t  mvp_idx[2];
    int16_t  mv[2][2], mvd[2][2];
    int32_t  nnz[3];                /* core->nnz */
    uint64_t coef_hash, rec_hash;   /* unused by the library (test bookkeeping) */
    int32_t  me_first, me_cnt;      /* unused by the library (test bookkeeping) */
} xb200_cu_item;

/* ---- intra analysis of one CU -----------------------------------------------------------------------
 * One ctx->fn_pintra_analyze_cu call (src_base/xeve_pintra.c:544-698), Baseline profile: the five intra modes (DC, HOR,
 * VER, UL, UR; xeve_ipred, src_base/xeve_ipred.c:99-228) ranked by SATD + mode bits (make_ipred_list, :308-374, pruned
 * against core->inter_satd), a luma RDO per surviving mode (transform, RDOQ, CABAC bit count, reconstruction, SSD;
 * pintra_residue_rdo :69-272), chroma with the winning mode, and the final CU bit count.  The neighbourhood enters as
 * the reference samples xeve_get_nbr (src_base/xeve_ipred.c:33-97) assembled from the picture being reconstructed: at
 * element offset nb_off of `side`, per plane c (size n = cuw, cuw/2, cuw/2): left[-1 .. 2n-1] then up[-1 .. 2n-1]
 * (2n+1 samples each, unavailable ones already replaced by 1 << (bit_depth-1)); 8*cuw + 6 samples per CU.
 * The two ctx.intra_dir context models of the coder states travel in the item (xb200_sbac holds the inter syntax). */
typedef struct {
    int32_t  poc, cur_pic;          /* POC and original-picture handle of the picture being coded */
    int16_t  x, y;
    uint8_t  log2_cuw, log2_cuh, slice_type, ctx_skip, ctx_pred_mode, all_preds;
    uint8_t  qp[3];                 /* core->qp_y/u/v */
    uint8_t  mpm[5];                /* core->mpm_b_list[ipm] = xeve_tbl_mpm[ipm_l][ipm_u][ipm] (xeve_get_mpm, src_base/xeve_ipred.c:230-252) */
    uint8_t  pad0_[2];
    uint32_t inter_satd;            /* core->inter_satd (src_base/xeve_mode.c:1247-1258), UINT32_MAX when there is no inter mode */
    int32_t  rate_idx;              /* RDOQ rate tables of the input coder state */
    int32_t  state_in, state_out;   /* coder state slots: core->s_curr_best[..] in, core->s_temp_best out */
    uint16_t cm_ipm_in[2];          /* ctx.intra_dir models of the input state */
    uint16_t cm_ipm_out[2];         /* result: the same models after the CU */
    double   lambda[3], sqrt_lambda0, dist_chroma_weight[2];
    int64_t  nb_off;                /* neighbour samples in `side` */
    int64_t  out_off;               /* element offset of this CU's coef / rec slots (3/2 * cuw * cuh each) */
    /* results */
    double   cost;                  /* return value (MAX_COST 1.7e308 when every mode was pruned) */
    int32_t  dist_cu;               /* core->dist_cu */
    int8_t   ipm[2];                /* core->ipm[] */
    uint8_t  pad1_[2];
    int32_t  nnz[3];
    uint64_t coef_hash, rec_hash;   /* unused by the library (test bookkeeping) */
} xb200_intra_item;

/* Inputs of xb200_analyze_intra taken from the picture being reconstructed: xeve_get_avail_intra (src_base/xeve_util.c:717-772),
 * xeve_get_nbr for Y, U, V (src_base/xeve_ipred.c:33-97) and xeve_get_mpm (:230-252), one tile.  The samples of item i are
 * written at element offset nb_off of `side` in the layout xb200_intra_item::nb_off expects (8 * cuw + 6 samples). */
typedef struct {
    int16_t  x, y;                  /* luma position of the CU */
    uint8_t  log2_cuw, log2_cuh;
    /* results */
    uint8_t  mpm[5];                /* xeve_tbl_mpm[ipm_left][ipm_up][0..4] */
    uint8_t  pad_;
    uint16_t avail;                 /* AVAIL_* bit mask of xeve_get_avail_intra (src_base/xeve_def.h:402-425) */
    uint16_t pad2_;
    int64_t  nb_off;                /* in: where the reference samples go */
} xb200_nbr_item;

/* ---- in-loop deblocking of a reconstructed picture --------------------------------------------------
 * One leaf CU of the coding tree, as xeve_deblock_tree hands it to ctx->fn_deblock_unit (src_base/xeve_df.c:575-639). */
typedef struct {
    int16_t x, y;                   /* luma position */
    uint8_t log2_cuw, log2_cuh, pad_[2];
} xb200_df_cu;

typedef struct {                    /* picture-level inputs of xb200_deblock */
    int32_t w_scu, h_scu;           /* ctx->w_scu / h_scu (4x4 units) */
    int32_t qp_u_offset, qp_v_offset; /* sh->qp_u_offset / qp_v_offset (-> pic->pic_qp_u/v_offset, src_base/xeve_df.c:594-595) */
    int32_t chroma_qp[2][70];       /* ctx->qp_chroma_dynamic[c][q] stored at [q + 6 * (bit_depth - 8)], q <= 57 */
} xb200_df_pic;

/* ---- lifetime ------------------------------------------------------------------------------ */
XB200_API int  xb200_create(xb200_ctx **out, int device, const xb200_seq *seq);
XB200_API void xb200_destroy(xb200_ctx *c);
XB200_API const char *xb200_version(void);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
XB200_API int64_t xb200_launch_count(const xb200_ctx *c);

/* ---- pictures (device-resident, s16, 4:2:0) ------------------------------------------------ */
/* padded != 0: planes carry PIC_PAD_SIZE_L/C of edge-replicated border like xeve_picbuf_alloc +
 * xeve_picbuf_expand; padded == 0: original-picture layout, stride = w (src_base/xeve_enc.c:1942). */
XB200_API int xb200_pic_create(xb200_ctx *c, int padded, int32_t *handle);
XB200_API int xb200_pic_destroy(xb200_ctx *c, int32_t handle);
/* planes[3]: caller's Y, U, V; stride_bytes[3]; in_bit_depth 8 (u8 samples) or 10..16 (u16 LE).
 * Converts to the internal depth (left shift) and, for padded pictures, replicates the borders. */
XB200_API int xb200_pic_upload(xb200_ctx *c, int32_t handle, const void *const planes[3], const int32_t stride_bytes[3],
                               int in_bit_depth, int mem);
/* raw s16 upload/download of the active area (internal samples) */
XB200_API int xb200_pic_upload_s16(xb200_ctx *c, int32_t handle, const int16_t *const planes[3],
                                   const int32_t stride_elems[3], int mem);
XB200_API int xb200_pic_download(xb200_ctx *c, int32_t handle, int with_padding, int16_t *const planes[3],
                                 const int32_t stride_elems[3]);

/* ---- kernel probes ---------------------------------------------------------------------------- */
XB200_API int xb200_sad(xb200_ctx *c, const xb200_blk_item *items, int64_t n, int32_t *out, int mem);
XB200_API int xb200_ssd(xb200_ctx *c, const xb200_blk_item *items, int64_t n, int64_t *out, int mem);
XB200_API int xb200_satd(xb200_ctx *c, const xb200_blk_item *items, int64_t n, int32_t *out, int mem);

/* forward DCT-II of n contiguous N x N s16 blocks (N = 1 << log2n, 32 or 64; |x| <= 2048) on the tcgen05
 * tensor cores -- the transform stage of xb200_residue exposed for parity checks against
 * xeve_trans (src_base/xeve_tq.c:396-404).  Host buffers. */
XB200_API int xb200_fwd_dct_tc(xb200_ctx *c, const int16_t *in, int16_t *out, int64_t n, int log2n);

/* map_scu: u32[f_scu]; map_mv: s16[f_scu][2][2]; col_mv[l]: s16[f_scu][2][2] = refp[0][l].map_mv.  Host buffers. */
XB200_API int xb200_mvp(xb200_ctx *c, xb200_mvp_item *items, int64_t n, const xb200_mvp_pic *pic, const uint32_t *map_scu,
                        const int16_t *map_mv, const int16_t *col_mv0, const int16_t *col_mv1);

/* states: in/out array of coder states (entries named by state_out are written); coef: s16 buffer.  Host buffers. */
XB200_API int xb200_rdo_bits(xb200_ctx *c, xb200_bits_item *items, int64_t n, xb200_sbac *states, int64_t n_states,
                             const int16_t *coef, int64_t coef_elems);
/* xeve_rdoq_bit_est (src_base/xeve_mode.c:304-373): RDOQ rate tables of n coder states (fields not used by the
 * Baseline RDOQ -- sig_coeff, gtx, last_sig_coeff -- are not produced) */
XB200_API int xb200_rdoq_rates(xb200_ctx *c, const xb200_sbac *states, int64_t n, xb200_rates *rates);

/* xeve_pinter_analyze_cu over a list of CUs (host buffers).  Every item is decided independently from its own inputs
 * (MVP candidates, input coder state, rate tables): CUs whose inputs do not depend on each other -- different pictures of
 * a temporal layer, or a wavefront inside a picture -- go into one call.  coef / rec receive the winner's coefficient
 * planes and reconstruction at items[i].out_off (rec may be NULL); states[items[i].state_out] receives s_next_best. */
XB200_API int xb200_analyze_cu(xb200_ctx *c, xb200_cu_item *items, int64_t n, const xb200_rates *rates, int64_t n_rates,
                               xb200_sbac *states, int64_t n_states, int16_t *coef, int16_t *rec, int64_t elems);

/* ---- hot-path operators ----------------------------------------------------------------------- */
/* side: s16 buffer holding the org_bi blocks referenced by org_bi_off (may be NULL if none). */
XB200_API int xb200_me(xb200_ctx *c, xb200_me_item *items, int64_t n, const int16_t *side, int64_t side_elems, int mem);
XB200_API int xb200_mc(xb200_ctx *c, const xb200_mc_item *items, int64_t n, const int64_t *pred_off, int16_t *pred,
                       int64_t pred_elems, int mem);
/* bi-search target of analyze_bi (src_base/xeve_pinter.c:143-156, 1620-1622): luma prediction of
 * items[i] (fn_mc) folded into org_bi = 2*org - pred, written as a contiguous w*h block at element
 * offset off[i] of `side` -- the buffer xb200_me reads through org_bi_off. */
XB200_API int xb200_bi_org(xb200_ctx *c, const xb200_mc_item *items, int64_t n, const int32_t *cur_pic, const int64_t *off,
                           int16_t *side, int64_t side_elems, int mem);
XB200_API int xb200_tq(xb200_ctx *c, xb200_tq_item *items, int64_t n, const xb200_rates *rates, int64_t n_rates,
                       int16_t *coef, int64_t coef_elems, int mem);
/* dequantise + inverse transform the planes whose nnz[] is non-zero, in place (fn_itdp) */
XB200_API int xb200_itdq(xb200_ctx *c, const xb200_tq_item *items, int64_t n, int16_t *coef, int64_t coef_elems, int mem);
/* rec = clip(pred + resi) per plane where nnz != 0, else clip(pred) (fn_recon) */
XB200_API int xb200_recon(xb200_ctx *c, const xb200_tq_item *items, int64_t n, const int16_t *resi, const int16_t *pred,
                          int16_t *rec, int64_t elems, int mem);
/* rec may be NULL: the per-candidate reconstruction is then only used for dist_rec and not stored -- the
 * reference itself recomputes the reconstruction of the winning mode (src_base/xeve_pinter.c:2006-2038).
 * With XB200_MEM_HOST the coefficient planes are read back compacted: a plane whose nnz[] is 0 is not copied and the caller's
 * `coef` is left untouched there (the reference never reads the coefficients of such a plane); with XB200_MEM_DEVICE every plane
 * is written. */
XB200_API int xb200_residue(xb200_ctx *c, xb200_residue_item *items, int64_t n, const xb200_rates *rates,
                            int64_t n_rates, int16_t *coef, int16_t *rec, int64_t elems, int mem);

/* pintra_analyze_cu over a list of CUs (host buffers); items are independent of each other, like xb200_analyze_cu.
 * CU sizes 4x4 .. 32x32 (param.min_cu_intra .. max_cu_intra of every preset but placebo). */
XB200_API int xb200_analyze_intra(xb200_ctx *c, xb200_intra_item *items, int64_t n, const xb200_rates *rates, int64_t n_rates,
                                  xb200_sbac *states, int64_t n_states, const int16_t *side, int64_t side_elems, int16_t *coef,
                                  int16_t *rec, int64_t elems);

/* pic: device picture holding the reconstruction so far (PIC_MODE).  map_scu: u32[f_scu] (COD bit 31 = already coded, IF bit 15
 * = intra); map_ipm: s8[f_scu] luma intra modes; constrained_intra_pred: pps.constrained_intra_pred_flag.  Host buffers. */
XB200_API int xb200_intra_nbr(xb200_ctx *c, int32_t pic, xb200_nbr_item *items, int64_t n, const uint32_t *map_scu, const int8_t *map_ipm,
                              int w_scu, int h_scu, int constrained_intra_pred, int16_t *side, int64_t side_elems);

/* In-loop deblocking of picture `pic` (a padded picture holding the unfiltered reconstruction), in place: every CU's left
 * edge (x > 0), then every CU's top edge (y > 0), 4-sample segments with the filter strength derived from the two
 * SCUs' intra / cbf / motion data, exactly as xeve_loop_filter does with one tile and one slice.  `cus` lists the leaf
 * CUs in coding order (left and upper neighbours of a CU precede it), covering the picture.  map_scu: u32[f_scu]
 * (MCU_* bit layout, src_base/xeve_def.h:585-640; only IF, QP, CBFL, IBC are read), map_refi: s8[f_scu][2], map_mv:
 * s16[f_scu][2][2].  expand != 0 replicates the borders afterwards (xeve_picbuf_expand), which makes the picture a
 * usable reference.  `mem` applies to cus and the three maps. */
XB200_API int xb200_deblock(xb200_ctx *c, int32_t pic, const xb200_df_cu *cus, int64_t n, const xb200_df_pic *pp,
                            const uint32_t *map_scu, const int8_t *map_refi, const int16_t *map_mv, int expand, int mem);

/* ---- the decision pass of a whole picture on the device (north_star: "the CTU tile loop in xeve_enc is lifted to the device") -----------
 * xb200_analyze_picture runs, for one picture, everything between the picture-level set-up and the entropy coder:
 *   the CTU loop of xeve_pic / xeve_ctu_mt_core            (src_base/xeve_enc.c:103-175, 327-404)
 *   mode_analyze_lcu -> mode_coding_tree -> mode_coding_unit -> mode_check_inter / mode_check_intra
 *                                                          (src_base/xeve_mode.c:1170-1348, 2007-2374, 2521-2608)
 *   with init_cu_data / copy_cu_data / copy_to_cu_data, update_map_scu / clear_map_scu, mode_cpy_rec_to_ref, update_to_ctx_map,
 *   xeve_pinter_analyze_cu and pintra_analyze_cu for every CU the tree visits (the operators above, as device functions),
 *   then ctx->fn_loop_filter and ctx->fn_picbuf_expand     (src_base/xeve_enc.c:407, 1274)
 * on the device with no host round trip per CU or per CTU: every coder-state chain of every picture in flight is a task of ONE
 * long-lived grid of 128-thread chain workers (the "chain server", a device-side FIFO fed by a per-device scheduler thread that all
 * contexts of the process share), followed by the two loop-filter grids and the border expansion of the picture.
 * A chain is what the reference calls a worker thread: with parallel_rows = n (param.threads, at most the CTU rows) CTU rows
 * y, y + n, .. form one chain, each chain starts from the reset coder state and a CTU waits for its upper-right neighbour
 * (src_base/xeve_enc.c:128-132) -- bit-exact with the reference run with `threads = n`; n = 1 is the single-thread bitstream.
 * Pictures whose references are complete run concurrently (the library orders them with events on the reference pictures).
 *
 * What returns to the host is what the reference's entropy coder reads (ctx->map_cu_data[lcu]): per CTU 256
 * xb200_scu_rec (16 x 16 units of 4x4 luma samples, raster inside the CTU) and the coefficient planes (Y 64x64 | U 32x32 | V 32x32,
 * CU rectangles in place).  The reconstruction stays on the device as picture rec_pic (deblocked, borders replicated: a usable
 * reference picture) together with its motion maps.  Restrictions: SLICE_B and SLICE_I (a P slice needs the bitstream-order state
 * walk, src_base/xeve_eco.c:1519), quad-tree CUs 4..64, no delta QP, rdo_dbk_switch = 0 (presets fast / medium). */
typedef struct {            /* decision of one 4x4 unit of a CTU */
    uint8_t mode, log2;     /* 0 SKIP, 1 DIRECT, 2 INTER, 3 INTRA; log2 size of the leaf CU covering the unit */
    int8_t  ipm, refi[2];
    uint8_t mvp_idx[2], pad_;
    int16_t mv[2][2], mvd[2][2];
    int32_t nnz[3];
} xb200_scu_rec;
typedef struct { xb200_sbac s; uint16_t ipm[2], split, pad_; } xb200_state;  /* full coder state: + ctx.intra_dir[2], ctx.split_cu_flag[0] */
typedef struct {
    int32_t  poc, slice_type;            /* ctx->poc.poc_val; SLICE_B 0, SLICE_P 1, SLICE_I 2 */
    int32_t  cur_pic, rec_pic;           /* original picture (unpadded) and the padded picture that receives the reconstruction */
    int32_t  tile_qp;                    /* ctx->tile[0].qp (the QP field of map_scu) */
    int32_t  num_refp[2], ref_pic[2][XB200_MAX_REFP], ref_poc[2][XB200_MAX_REFP]; /* ctx->rpm.num_refp, ctx->refp[refi][lidx] as [lidx][refi] */
    int32_t  col_list_poc0;              /* ctx->refp[0][REFP_1].list_poc[0] */
    int32_t  max_cu_inter, min_cu_inter, max_cu_intra, min_cu_intra, cip; /* param.*; pps.constrained_intra_pred_flag */
    int32_t  qp[3];                      /* core->qp_y / qp_u / qp_v */
    uint32_t lambda_mv;                  /* pi->lambda_mv */
    int32_t  max_search_range;           /* pi->max_search_range */
    int32_t  parallel_rows;              /* ctx->parallel_rows */
    int32_t  deblock;                    /* sh->deblocking_filter_on */
    int32_t  unfiltered_pic;             /* -1, or a padded picture that receives a copy of the reconstruction before the loop filter */
    double   lambda[3], sqrt_lambda0, dist_chroma_weight[2];   /* core->lambda[], sqrt_lambda[0], dist_chroma_weight[] */
    xb200_df_pic df;                     /* loop-filter inputs (w_scu / h_scu are filled by the library) */
} xb200_picture;
typedef struct {            /* what xb200_picture_fetch reports besides the records */
    int64_t n_inter, n_intra;            /* CU analyses the decision pass ran */
    double  chain_ms, filter_ms;         /* device time of the decision kernel and of the loop filter + border expansion */
} xb200_picture_stat;

/* Queues the picture and returns at once; it runs when its reference pictures are complete and the workers can take all its chains.
 * Fails with XB200_ERR_UNSUPPORTED for P slices and out-of-range CU sizes.  A picture handle (rec_pic, cur_pic, reference pictures)
 * must not be reused or destroyed while a queued picture reads or writes it: fetch the pictures that reference it first. */
XB200_API int xb200_analyze_picture(xb200_ctx *c, const xb200_picture *pp);
/* Waits for picture rec_pic and copies its records to the caller (host buffers; any pointer may be NULL): scu [n_lcu * 256],
 * coef [n_lcu * 6144], ctu_states [n_lcu][2] (the coder state each CTU's decision pass started from / ended with), ctu_cost [n_lcu]. */
XB200_API int xb200_picture_fetch(xb200_ctx *c, int32_t rec_pic, xb200_scu_rec *scu, int16_t *coef, xb200_state *ctu_states,
                                  double *ctu_cost, xb200_picture_stat *stat);
/* 1 when picture rec_pic is complete (xb200_picture_fetch would not block), 0 while it is queued or running, < 0 for a handle with
 * no picture in flight.  Lets a caller that must not block (xeve_encode returning XEVE_OK_OUT_NOT_AVAILABLE) poll. */
XB200_API int xb200_picture_ready(xb200_ctx *c, int32_t rec_pic);
/* Frame maps of a decided picture as the reference holds them when the loop filter starts (host buffers, any may be NULL):
 * map_scu u32[f_scu], map_ipm s8[f_scu], map_refi s8[f_scu][2], map_mv s16[f_scu][2][2]. */
XB200_API int xb200_picture_maps(xb200_ctx *c, int32_t rec_pic, uint32_t *map_scu, int8_t *map_ipm, int8_t *map_refi, int16_t *map_mv);
/* Adopt a reference picture decided elsewhere (another GPU of the picture DAG, a test): its motion map (refp.map_mv, what later
 * pictures read as colocated MVs) is attached to padded picture `pic`, whose planes the caller uploads as usual. */
XB200_API int xb200_picture_adopt(xb200_ctx *c, int32_t pic, const int16_t *map_mv);
/* Debug aid: keep the records of the first cap_cu inter / cap_intra intra CU analyses of every following picture (in call order);
 * xb200_picture_log copies those of picture rec_pic to the caller and returns their counts through n[2]. */
XB200_API int xb200_picture_log_enable(xb200_ctx *c, int64_t cap_cu, int64_t cap_intra);
XB200_API int xb200_picture_log(xb200_ctx *c, int32_t rec_pic, xb200_cu_item *cu, xb200_intra_item *intra, int64_t n[2]);
/* Number of chain workers of the device (CTAs of the chain server for the default working set); pictures beyond it queue on the host. */
XB200_API int xb200_chain_capacity(xb200_ctx *c);
/* Device time (ms, CUDA events) from the first picture enqueued after the last reset to the latest completion among the pictures
 * fetched since: the span of a batch of pictures that ran concurrently on the library's own streams.  reset != 0 starts a new batch. */
XB200_API double xb200_chain_span_ms(xb200_ctx *c, int reset);
/* Debug builds only (-DXB200_CHAIN_DEBUG, otherwise XB200_ERR_UNSUPPORTED): progress words of the decision kernel, kept in host-mapped
 * memory so that they can be read while a kernel hangs.  Never blocks. */
XB200_API int xb200_chain_debug(xb200_ctx *c, int32_t out[64]);
/* Profiling builds only (-DXB200_CHAIN_PROF, otherwise XB200_ERR_UNSUPPORTED): SM cycles [0..31] and visit counts [32..63] per phase of
 * the decision kernel since the previous call (phase numbers: xb200_analyze.cuh / xb200_chain.cuh CU_PROF). */
XB200_API int xb200_chain_prof(xb200_ctx *c, uint64_t out[64]);

/* ---- Main profile, first operator: the two-stage 16-bit transforms -------------------------------------------------
 * Forward / inverse transform of a list of s16 blocks in place (row-major, w * h samples at element offset `off` of `blocks`):
 *   ats = 0: the "IQT" DCT-II of sps.tool_iqt -- xeve_trans with iqt_flag (src_main/xevem_tq.c:709-716, stages tx_pb2 .. tx_pb64
 *            :58-334) and xeve_itrans (src_main/xevem_itdq.c:551-557, stages itx_pb2 .. itx_pb64 :302-549); 2..64 samples per side;
 *   ats = 1: DST-VII / DCT-VIII chosen per direction by tridx (ats_intra_tridx: bit 1 horizontal, bit 0 vertical; 0 -> DST-VII,
 *            1 -> DCT-VIII) -- xeve_t_MxN_ats_intra (src_main/xevem_tq.c:684-702) and xeve_it_MxN_ats_intra
 *            (src_main/xevem_itdq.c:278-300); 4..32 samples per side.
 * Both stages round, shift and store 16 bits (forward: truncation, inverse: saturation) exactly as the reference's; blocks need not be
 * square.  The bit depth is the context's (xb200_seq::bit_depth).  `mem` applies to items and blocks. */
typedef struct {
    uint8_t log2_w, log2_h;
    uint8_t inverse;                /* 0 forward (residual -> coefficients), 1 inverse */
    uint8_t ats, tridx;
    uint8_t pad_[3];
    int64_t off;
} xb200_trm_item;
XB200_API int xb200_transform_main(xb200_ctx *c, const xb200_trm_item *items, int64_t n, int16_t *blocks, int64_t elems, int mem);

/* ---- timing aid for bench.py: device time (ms) of the kernels of the last call, measured with
 *      CUDA events on the library's own stream ------------------------------------------------- */
XB200_API double xb200_last_kernel_ms(const xb200_ctx *c);

#ifdef __cplusplus
}
#endif
#endif /* XEVE_B200_H_ */
