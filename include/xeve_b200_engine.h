/* xeve_b200_engine.h -- the picture-level entry points of include/xeve_b200.h as a table.
 *
 * The reference-side binding (integration/xeve_b200_dropin.c: the reference's own xeve_create / xeve_push / xeve_encode with the
 * decision pass of every picture handed to the device) calls the library through this table.  The default table holds the
 * xb200_* functions of libxeve_b200.so and nothing else ships; the table exists so that the host plumbing of the binding
 * (picture plans, hand-over order, tail re-planning) can be pinned on machines without a GPU by a TEST that installs a table
 * of CPU functions (oracle/engine_standin.c, test infrastructure) through xeve_b200_set_engine().  The product never installs
 * another table by itself: without an sm_100 device create() fails and so does xeve_create. */
#ifndef XEVE_B200_ENGINE_H
#define XEVE_B200_ENGINE_H
#include "xeve_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct xb200_engine {
    int  (*create)(xb200_ctx **out, int device, const xb200_seq *seq);                              /* xb200_create */
    void (*destroy)(xb200_ctx *c);                                                                  /* xb200_destroy */
    int  (*pic_create)(xb200_ctx *c, int padded, int32_t *handle);                                  /* xb200_pic_create */
    int  (*pic_destroy)(xb200_ctx *c, int32_t handle);                                              /* xb200_pic_destroy */
    int  (*pic_upload)(xb200_ctx *c, int32_t handle, const void *const planes[3], const int32_t stride_bytes[3], int in_bit_depth,
                       int mem);                                                                    /* xb200_pic_upload */
    int  (*pic_download)(xb200_ctx *c, int32_t handle, int with_padding, int16_t *const planes[3],
                         const int32_t stride_elems[3]);                                            /* xb200_pic_download */
    int  (*analyze_picture)(xb200_ctx *c, const xb200_picture *pp);                                 /* xb200_analyze_picture */
    int  (*picture_fetch)(xb200_ctx *c, int32_t rec_pic, xb200_scu_rec *scu, int16_t *coef, xb200_state *ctu_states, double *ctu_cost,
                          xb200_picture_stat *stat);                                                /* xb200_picture_fetch */
    int  (*picture_ready)(xb200_ctx *c, int32_t rec_pic);                                           /* xb200_picture_ready */
} xb200_engine;

/* Test seam of the drop-in library (see above).  NULL restores the default (CUDA) table.  Affects encoders created afterwards. */
XB200_API void xeve_b200_set_engine(const xb200_engine *e);

/* What the drop-in did for an encoder instance (id = the XEVE handle), for tests and benches: pictures decided on the device,
 * pictures whose speculative plan had to be redone at the end of the stream, CU analyses run, summed kernel time. */
typedef struct {
    int64_t pictures, replanned, n_inter, n_intra, deferred;   /* deferred: xeve_encode calls answered XEVE_OK_OUT_NOT_AVAILABLE while the device worked */
    double  chain_ms, filter_ms, wait_ms;   /* wait_ms: host time xeve_encode spent blocked on the device */
    int32_t device_path;                    /* 1: decisions on the device; 0: configuration outside the path, reference host code ran */
    int32_t pad_;
} xeve_b200_stats;
XB200_API int xeve_b200_get_stats(void *id, xeve_b200_stats *out);

#ifdef __cplusplus
}
#endif
#endif
