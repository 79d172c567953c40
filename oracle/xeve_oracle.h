/* xeve_oracle.h -- TEST INFRASTRUCTURE ONLY: prototypes of the scalar CPU restatement
 * (oracle/xeve_oracle.c).  Work-list records are the C-ABI's own (include/xeve_b200.h). */
#ifndef XEVE_ORACLE_H_
#define XEVE_ORACLE_H_
#include "../include/xeve_b200.h"

typedef struct {          /* one picture: pointers to the top-left sample of the active area */
    int16_t *y, *u, *v;
    int32_t  s_l, s_c, w_l, h_l, poc;
} xo_planes;

/* CU decision chain (xo_chain_picture): full coder state and the per-CTU record == RH_LCU_REC of oracle/ref_harness.c */
typedef struct { xb200_sbac s; uint16_t ipm[2], split, pad_; } xo_state;
typedef struct {          /* decision of one 4x4 unit of a CTU (XEVE_CU_DATA fields the entropy coder and the loop filter read) */
    uint8_t mode, log2;    /* 0 SKIP, 1 DIRECT, 2 INTER, 3 INTRA; log2 size of the leaf CU covering the unit */
    int8_t  ipm, refi[2];
    uint8_t mvp_idx[2], pad_;
    int16_t mv[2][2], mvd[2][2];
    int32_t nnz[3];
} xo_scu_rec;
typedef struct {
    int32_t  poc, slice_type, lcu_num, x_pel, y_pel, tile_qp, cur_pic;
    int32_t  num_refp[2], ref_pic[2][4], ref_poc[2][4], col_list_poc0;
    int32_t  max_cu_inter, min_cu_inter, max_cu_intra, min_cu_intra, cip;
    int32_t  qp[3];
    uint32_t lambda_mv;
    int32_t  max_search_range, parallel_rows; /* ctx->parallel_rows (threads, at most the CTU rows): CTU rows y, y + n, .. share a state chain */
    double   lambda[3], sqrt_lambda0, dist_chroma_weight[2];
    int64_t  col_off[2];
    xo_state state_in, state_out;
} xo_ctu_rec;

#define XO_API __attribute__((visibility("default")))
XO_API int     xo_sad(int w, int h, const int16_t *a, int sa, const int16_t *b, int sb, int bd);
XO_API int64_t xo_ssd(int w, int h, const int16_t *a, int sa, const int16_t *b, int sb, int bd);
XO_API void    xo_diff(int w, int h, const int16_t *a, int sa, const int16_t *b, int sb, int16_t *d, int sd);
XO_API int     xo_satd(int w, int h, const int16_t *a, int sa, const int16_t *b, int sb, int bd);
XO_API void xo_mc_luma(const int16_t *ref, int sr, int gx, int gy, int sel_x, int sel_y, int16_t *pred, int sp, int w,
                       int h, int bd);
XO_API void xo_mc_chroma(const int16_t *ref, int sr, int gx, int gy, int sel_x, int sel_y, int16_t *pred, int sp, int w,
                         int h, int bd);
XO_API void xo_mc(const xb200_seq *sq, const xo_planes *pl, const xb200_mc_item *it, int16_t *pred);
XO_API void xo_fwd_transform(int16_t *blk, int log2w, int log2h, int bd);
XO_API void xo_inv_transform(int16_t *blk, int log2w, int log2h, int bd);
XO_API void xo_iqt_fwd(int16_t *blk, int log2w, int log2h, int bd);
XO_API void xo_iqt_inv(int16_t *blk, int log2w, int log2h, int bd);
XO_API void xo_ats_matrix(int type, int log2n, int8_t *out);
XO_API void xo_ats_fwd(int16_t *blk, int log2w, int log2h, int bd, int tridx);
XO_API void xo_ats_inv(int16_t *blk, int log2w, int log2h, int bd, int tridx);
XO_API int  xo_quant_rdoq(int16_t *coef, int log2n, int qp, double d_lambda, int is_intra, int ch, int slice_type,
                          const xb200_rates *rt, int bd);
XO_API int  xo_quant_plain(int16_t *coef, int log2n, int qp, int slice_type, int bd);
XO_API void xo_tq(const xb200_seq *sq, const xb200_tq_item *it, const xb200_rates *rates, int16_t *planes, int nnz[3]);
XO_API void xo_itdq(const xb200_seq *sq, const xb200_tq_item *it, int16_t *planes, const int nnz[3]);
XO_API void xo_recon(const int16_t *resi, const int16_t *pred, int is_coef, int n, int16_t *rec, int bd);
XO_API void xo_me(const xb200_seq *sq, const xo_planes *pl, const int16_t *side, xb200_me_item *it);
XO_API void xo_residue(const xb200_seq *sq, const xo_planes *pl, const xb200_rates *rates, xb200_residue_item *it,
                       int16_t *coef, int16_t *rec);
XO_API void xo_bi_org(const xb200_seq *sq, const xo_planes *pl, const xb200_mc_item *it, int cur_pic, int16_t *org_bi);
XO_API void xo_bi_org_batch(const xb200_seq *sq, const xo_planes *pl, const xb200_mc_item *items, int64_t n,
                            const int32_t *cur_pic, const int64_t *off, int16_t *side);
XO_API void xo_rdo_bits(xb200_bits_item *it, xb200_sbac *states, const int16_t *coef);
XO_API void xo_rdo_bits_batch(xb200_bits_item *items, int64_t n, xb200_sbac *states, const int16_t *coef);
XO_API void xo_rdoq_rates(const xb200_sbac *st, int64_t n, xb200_rates *out);
XO_API void xo_analyze_cu(const xb200_seq *sq, const xo_planes *pl, const xb200_rates *rates, xb200_cu_item *cu, xb200_sbac *states,
                          int16_t *coef_out, int16_t *rec_out);
XO_API void xo_analyze_cu_batch(const xb200_seq *sq, const xo_planes *pl, const xb200_rates *rates, xb200_cu_item *items, int64_t n,
                                xb200_sbac *states, int16_t *coef, int16_t *rec);
XO_API void xo_analyze_intra(const xb200_seq *sq, const xo_planes *pl, const xb200_rates *rates, xb200_intra_item *it, xb200_sbac *states,
                             const int16_t *side, int16_t *coef_out, int16_t *rec_out);
XO_API void xo_analyze_intra_batch(const xb200_seq *sq, const xo_planes *pl, const xb200_rates *rates, xb200_intra_item *items,
                                   int64_t n, xb200_sbac *states, const int16_t *side, int16_t *coef, int16_t *rec);
XO_API void xo_intra_nbr(const int16_t *y, const int16_t *u, const int16_t *v, int s_l, int s_c, xb200_nbr_item *it, const uint32_t *map_scu,
                         const int8_t *map_ipm, int w_scu, int h_scu, int cip, int bd, int16_t *side);
XO_API void xo_intra_nbr_batch(const int16_t *y, const int16_t *u, const int16_t *v, int s_l, int s_c, xb200_nbr_item *items, int64_t n,
                               const uint32_t *map_scu, const int8_t *map_ipm, int w_scu, int h_scu, int cip, int bd, int16_t *side);
XO_API void xo_mvp(xb200_mvp_item *it, const xb200_mvp_pic *pp, const uint32_t *map_scu, const int16_t *map_mv, const int16_t *col_mv0,
                   const int16_t *col_mv1);
XO_API void xo_mvp_batch(xb200_mvp_item *items, int64_t n, const xb200_mvp_pic *pp, const uint32_t *map_scu, const int16_t *map_mv,
                         const int16_t *col_mv0, const int16_t *col_mv1);
typedef void (*xo_chain_cu_fn)(xb200_cu_item *cu, xb200_sbac *states, const xb200_rates *rates, int16_t *coef, int16_t *rec, int16_t *pred_y);
typedef void (*xo_chain_intra_fn)(xb200_intra_item *it, xb200_sbac *states, const xb200_rates *rates, const int16_t *side, int16_t *coef,
                                  int16_t *rec);
typedef void (*xo_chain_mvp_fn)(xb200_mvp_item *it, const xb200_mvp_pic *pp);
typedef void (*xo_chain_nbr_fn)(xb200_nbr_item *it, int16_t *side);
XO_API void xo_chain_set_callbacks(xo_chain_cu_fn cu_fn, xo_chain_intra_fn intra_fn);
XO_API void xo_chain_set_input_callbacks(xo_chain_mvp_fn mvp_fn, xo_chain_nbr_fn nbr_fn);
XO_API int  xo_sizeof_chain(int what);
XO_API void xo_chain_picture(const xb200_seq *sq, const xo_planes *pl, const xo_ctu_rec *pp, const int16_t *col_mv0, const int16_t *col_mv1,
                             xo_ctu_rec *out, double *ctu_cost, int16_t *rec_y, int16_t *rec_u, int16_t *rec_v, int s_l, int s_c,
                             uint32_t *map_scu, int8_t *map_ipm, int8_t *map_refi, int16_t *map_mv, xb200_df_cu *cus, int64_t cus_cap,
                             xb200_cu_item *cu_log, int64_t cu_cap, xb200_intra_item *intra_log, int64_t intra_cap, int64_t *n_out,
                             int ctu_limit, xo_scu_rec *scu_out, int16_t *coef_out);
XO_API void xo_hash_slots(const int16_t *buf, const int64_t *off, const int64_t *elems, int64_t n, uint64_t *out);
XO_API void xo_deblock(int16_t *y, int16_t *u, int16_t *v, int s_l, int s_c, int w, int h, const xb200_df_cu *cus, int64_t n,
                       const xb200_df_pic *pp, const uint32_t *map_scu, const int8_t *map_refi, const int16_t *map_mv, int bit_depth);
XO_API void xo_pad_plane(int16_t *buf, int stride, int w, int h, int pad);
XO_API void xo_me_batch(const xb200_seq *sq, const xo_planes *pl, const int16_t *side, xb200_me_item *items, int64_t n);
XO_API void xo_mc_batch(const xb200_seq *sq, const xo_planes *pl, const xb200_mc_item *items, int64_t n,
                        const int64_t *off, int16_t *pred);
XO_API void xo_tq_batch(const xb200_seq *sq, xb200_tq_item *items, int64_t n, const xb200_rates *rates, int16_t *coef,
                        int16_t *resi);
XO_API void xo_residue_batch(const xb200_seq *sq, const xo_planes *pl, const xb200_rates *rates,
                             xb200_residue_item *items, int64_t n, int16_t *coef, int16_t *rec);
#endif
