"""Python binding of the C ABI in include/xeve_b200.h (libxeve_b200.so, hand-written sm_100a CUDA).

This module is plumbing only: numpy record dtypes that mirror the C structs byte for byte, and a
thin ``Hotpath`` object around the opaque context.  There is no CPU implementation behind it --
if the shared library or a B200 is missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("XB200_LIB") or os.path.join(_HERE, "libxeve_b200.so")   # XB200_LIB: an alternative build (profiling hooks)

OK, ERR, ERR_INVALID_ARGUMENT, ERR_OUT_OF_MEMORY, ERR_UNSUPPORTED, ERR_UNEXPECTED = 0, -1, -101, -102, -104, -105
MEM_HOST, MEM_DEVICE = 0, 1
PAD_L, PAD_C = 144, 72

SEQ = np.dtype([
    ("w", "<i4"), ("h", "<i4"), ("bit_depth", "<i4"), ("me_level", "<i4"), ("hpel_cnt", "<i4"), ("qpel_cnt", "<i4"),
    ("me_complexity", "<i4"), ("min_clip", "<i4", (2,)), ("max_clip", "<i4", (2,)),
    ("merge_num", "<i4"), ("me_range", "<i4"), ("gop_size", "<i4"), ("rdoq", "<i4"), ("tool_iqt", "<i4"),
], align=True)

ME_ITEM = np.dtype([
    ("poc", "<i4"), ("cur_pic", "<i4"), ("ref_pic", "<i4"), ("ref_poc", "<i4"),
    ("x", "<i2"), ("y", "<i2"),
    ("log2_cuw", "u1"), ("log2_cuh", "u1"), ("lidx", "u1"), ("bi", "u1"),
    ("refi", "i1"), ("num_refp", "u1"),
    ("mvp", "<i2", (2,)), ("mv_in", "<i2", (2,)),
    ("lambda_mv", "<u4"), ("mot_bits_in", "<i4", (2,)),
    ("max_search_range", "<i4"), ("gop_size", "<i4"), ("org_bi_off", "<i4"),
    ("mv_out", "<i2", (2,)), ("cost", "<u4"), ("mot_bits_out", "<i4", (2,)),
], align=True)

MC_ITEM = np.dtype([
    ("poc", "<i4"), ("ref_pic", "<i4", (2,)), ("ref_poc", "<i4", (2,)),
    ("x", "<i2"), ("y", "<i2"), ("w", "<i2"), ("h", "<i2"),
    ("refi", "i1", (2,)), ("mv", "<i2", (2, 2)),
    ("out_hash", "<u8"),
], align=True)

RATES = np.dtype([
    ("cbf_all", "<i4", (2,)), ("cbf_luma", "<i4", (2,)), ("cbf_cb", "<i4", (2,)), ("cbf_cr", "<i4", (2,)),
    ("run", "<i4", (24, 2)), ("level", "<i4", (24, 2)), ("last", "<i4", (2, 2)),
], align=True)

TQ_ITEM = np.dtype([
    ("poc", "<i4"),
    ("log2_cuw", "u1"), ("log2_cuh", "u1"), ("slice_type", "u1"), ("is_intra", "u1"),
    ("run_stats", "u1"), ("qp", "u1", (3,)),
    ("rate_idx", "<i4"), ("in_off", "<i8"),
    ("lambda", "<f8", (3,)), ("nnz", "<i4", (3,)),
    ("out_hash", "<u8"),
], align=True)

RESIDUE_ITEM = np.dtype([
    ("mc", MC_ITEM),
    ("cur_pic", "<i4"),
    ("slice_type", "u1"), ("run_stats", "u1"), ("qp", "u1", (3,)), ("pad_", "u1", (3,)),
    ("rate_idx", "<i4"),
    ("lambda", "<f8", (3,)),
    ("out_off", "<i8"),
    ("nnz", "<i4", (3,)),
    ("dist_pred", "<i8", (3,)), ("dist_rec", "<i8", (3,)),
], align=True)

BLK_ITEM = np.dtype([
    ("pic1", "<i4"), ("pic2", "<i4"),
    ("x1", "<i2"), ("y1", "<i2"), ("x2", "<i2"), ("y2", "<i2"),
    ("plane1", "u1"), ("plane2", "u1"), ("log2w", "u1"), ("log2h", "u1"),
], align=True)

MVP_ITEM = np.dtype([
    ("x_scu", "<i2"), ("y_scu", "<i2"), ("log2_cuw", "u1"), ("log2_cuh", "u1"), ("lidx", "u1"), ("pad_", "u1"),
    ("avail", "<u2"), ("refi", "i1", (4,)), ("mvp", "<i2", (4, 2)), ("mv_dir", "<i2", (2, 2)),
], align=True)
MVP_PIC = np.dtype([("w_scu", "<i4"), ("h_scu", "<i4"), ("poc", "<i4"), ("ref_poc", "<i4", (2,)), ("col_list_poc0", "<i4")], align=True)

INTRA_ITEM = np.dtype([
    ("poc", "<i4"), ("cur_pic", "<i4"), ("x", "<i2"), ("y", "<i2"), ("log2_cuw", "u1"), ("log2_cuh", "u1"), ("slice_type", "u1"),
    ("ctx_skip", "u1"), ("ctx_pred_mode", "u1"), ("all_preds", "u1"), ("qp", "u1", (3,)), ("mpm", "u1", (5,)), ("pad0_", "u1", (2,)),
    ("inter_satd", "<u4"), ("rate_idx", "<i4"), ("state_in", "<i4"), ("state_out", "<i4"), ("cm_ipm_in", "<u2", (2,)),
    ("cm_ipm_out", "<u2", (2,)), ("lambda", "<f8", (3,)), ("sqrt_lambda0", "<f8"), ("dist_chroma_weight", "<f8", (2,)),
    ("nb_off", "<i8"), ("out_off", "<i8"), ("cost", "<f8"), ("dist_cu", "<i4"), ("ipm", "i1", (2,)), ("pad1_", "u1", (2,)),
    ("nnz", "<i4", (3,)), ("coef_hash", "<u8"), ("rec_hash", "<u8"),
], align=True)

TRM_ITEM = np.dtype([("log2_w", "u1"), ("log2_h", "u1"), ("inverse", "u1"), ("ats", "u1"), ("tridx", "u1"), ("pad_", "u1", (3,)),
                     ("off", "<i8")], align=True)
NBR_ITEM = np.dtype([("x", "<i2"), ("y", "<i2"), ("log2_cuw", "u1"), ("log2_cuh", "u1"), ("mpm", "u1", (5,)), ("pad_", "u1"),
                     ("avail", "<u2"), ("pad2_", "<u2"), ("nb_off", "<i8")], align=True)

DF_CU = np.dtype([("x", "<i2"), ("y", "<i2"), ("log2_cuw", "u1"), ("log2_cuh", "u1"), ("pad_", "u1", (2,))], align=True)
DF_PIC = np.dtype([("w_scu", "<i4"), ("h_scu", "<i4"), ("qp_u_offset", "<i4"), ("qp_v_offset", "<i4"),
                   ("chroma_qp", "<i4", (2, 70))], align=True)

SBAC = np.dtype([("range", "<u4"), ("m", "<u2", (68,))], align=True)
BITS_ITEM = np.dtype([
    ("kind", "u1"), ("slice_type", "u1"), ("log2_cuw", "u1"), ("log2_cuh", "u1"), ("pidx", "u1"), ("ch", "u1"),
    ("ctx_skip", "u1"), ("ctx_pred_mode", "u1"), ("refi", "i1", (2,)), ("mvp_idx", "u1", (2,)), ("num_refp", "u1", (2,)),
    ("all_preds", "u1"), ("pad_", "u1"), ("mvd", "<i2", (2, 2)), ("nnz", "<i4", (3,)), ("state_in", "<i4"),
    ("state_out", "<i4"), ("coef_off", "<i8"), ("bits", "<u4"), ("pad2_", "<u4"),
], align=True)
(CM_SKIP_FLAG, CM_PRED_MODE, CM_DIRECT, CM_INTER_DIR, CM_REFI, CM_MVP_IDX, CM_MVD, CM_CBF_ALL, CM_CBF_LUMA, CM_CBF_CB,
 CM_CBF_CR, CM_RUN, CM_LAST, CM_LEVEL, CM_COUNT) = (0, 2, 5, 6, 8, 10, 13, 14, 15, 16, 17, 18, 42, 44, 68)

CU_ITEM = np.dtype([
    ("poc", "<i4"), ("cur_pic", "<i4"), ("x", "<i2"), ("y", "<i2"), ("log2_cuw", "u1"), ("log2_cuh", "u1"), ("slice_type", "u1"),
    ("ctx_skip", "u1"), ("ctx_pred_mode", "u1"), ("all_preds", "u1"), ("num_refp", "u1", (2,)), ("qp", "u1", (3,)), ("pad0_", "u1"),
    ("max_search_range", "<i4"), ("ref_pic", "<i4", (2, 4)), ("ref_poc", "<i4", (2, 4)), ("lambda_mv", "<u4"), ("rate_idx", "<i4"),
    ("state_in", "<i4"), ("state_out", "<i4"), ("lambda", "<f8", (3,)), ("dist_chroma_weight", "<f8", (2,)),
    ("mvp", "<i2", (2, 4, 2)), ("refi_pred", "i1", (2, 4)), ("mv_dir", "<i2", (2, 2)), ("out_off", "<i8"), ("cost", "<f8"),
    ("best_idx", "u1"), ("pad1_", "u1"), ("refi", "i1", (2,)), ("mvp_idx", "u1", (2,)), ("mv", "<i2", (2, 2)), ("mvd", "<i2", (2, 2)),
    ("nnz", "<i4", (3,)), ("coef_hash", "<u8"), ("rec_hash", "<u8"), ("me_first", "<i4"), ("me_cnt", "<i4"),
], align=True)

# ---- picture-level decision pass (xb200_analyze_picture) ----
SCU_REC = np.dtype([("mode", "u1"), ("log2", "u1"), ("ipm", "i1"), ("refi", "i1", (2,)), ("mvp_idx", "u1", (2,)), ("pad_", "u1"),
                    ("mv", "<i2", (2, 2)), ("mvd", "<i2", (2, 2)), ("nnz", "<i4", (3,))], align=True)
STATE = np.dtype([("s", SBAC), ("ipm", "<u2", (2,)), ("split", "<u2"), ("pad_", "<u2")], align=True)
PICTURE = np.dtype([
    ("poc", "<i4"), ("slice_type", "<i4"), ("cur_pic", "<i4"), ("rec_pic", "<i4"), ("tile_qp", "<i4"),
    ("num_refp", "<i4", (2,)), ("ref_pic", "<i4", (2, 4)), ("ref_poc", "<i4", (2, 4)), ("col_list_poc0", "<i4"),
    ("max_cu_inter", "<i4"), ("min_cu_inter", "<i4"), ("max_cu_intra", "<i4"), ("min_cu_intra", "<i4"), ("cip", "<i4"),
    ("qp", "<i4", (3,)), ("lambda_mv", "<u4"), ("max_search_range", "<i4"), ("parallel_rows", "<i4"), ("deblock", "<i4"),
    ("unfiltered_pic", "<i4"), ("lambda", "<f8", (3,)), ("sqrt_lambda0", "<f8"), ("dist_chroma_weight", "<f8", (2,)),
    ("df", DF_PIC),
], align=True)
PICTURE_STAT = np.dtype([("n_inter", "<i8"), ("n_intra", "<i8"), ("chain_ms", "<f8"), ("filter_ms", "<f8")], align=True)

VP = C.c_void_p
_lib = None


class Xb200Error(RuntimeError):
    def __init__(self, code, what):
        super().__init__(f"{what} failed with status {code}")
        self.code = code


def load():
    """dlopen the in-tree library; raises if it was not built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        L.xb200_version.restype = C.c_char_p
        L.xb200_create.argtypes = [C.POINTER(VP), C.c_int, VP]
        L.xb200_destroy.argtypes = [VP]
        L.xb200_destroy.restype = None
        L.xb200_launch_count.argtypes = [VP]
        L.xb200_launch_count.restype = C.c_int64
        L.xb200_last_kernel_ms.argtypes = [VP]
        L.xb200_last_kernel_ms.restype = C.c_double
        L.xb200_pic_create.argtypes = [VP, C.c_int, C.POINTER(C.c_int32)]
        L.xb200_pic_destroy.argtypes = [VP, C.c_int32]
        L.xb200_pic_upload.argtypes = [VP, C.c_int32, VP, VP, C.c_int, C.c_int]
        L.xb200_pic_upload_s16.argtypes = [VP, C.c_int32, VP, VP, C.c_int]
        L.xb200_pic_download.argtypes = [VP, C.c_int32, C.c_int, VP, VP]
        for f in (L.xb200_sad, L.xb200_ssd, L.xb200_satd):
            f.argtypes = [VP, VP, C.c_int64, VP, C.c_int]
        L.xb200_me.argtypes = [VP, VP, C.c_int64, VP, C.c_int64, C.c_int]
        L.xb200_mc.argtypes = [VP, VP, C.c_int64, VP, VP, C.c_int64, C.c_int]
        L.xb200_bi_org.argtypes = [VP, VP, C.c_int64, VP, VP, VP, C.c_int64, C.c_int]
        L.xb200_fwd_dct_tc.argtypes = [VP, VP, VP, C.c_int64, C.c_int]
        L.xb200_mvp.argtypes = [VP, VP, C.c_int64, VP, VP, VP, VP, VP]
        L.xb200_tq.argtypes = [VP, VP, C.c_int64, VP, C.c_int64, VP, C.c_int64, C.c_int]
        L.xb200_itdq.argtypes = [VP, VP, C.c_int64, VP, C.c_int64, C.c_int]
        L.xb200_recon.argtypes = [VP, VP, C.c_int64, VP, VP, VP, C.c_int64, C.c_int]
        L.xb200_residue.argtypes = [VP, VP, C.c_int64, VP, C.c_int64, VP, VP, C.c_int64, C.c_int]
        L.xb200_rdo_bits.argtypes = [VP, VP, C.c_int64, VP, C.c_int64, VP, C.c_int64]
        L.xb200_rdoq_rates.argtypes = [VP, VP, C.c_int64, VP]
        L.xb200_analyze_cu.argtypes = [VP, VP, C.c_int64, VP, C.c_int64, VP, C.c_int64, VP, VP, C.c_int64]
        L.xb200_analyze_intra.argtypes = [VP, VP, C.c_int64, VP, C.c_int64, VP, C.c_int64, VP, C.c_int64, VP, VP, C.c_int64]
        L.xb200_intra_nbr.argtypes = [VP, C.c_int32, VP, C.c_int64, VP, VP, C.c_int, C.c_int, C.c_int, VP, C.c_int64]
        L.xb200_deblock.argtypes = [VP, C.c_int32, VP, C.c_int64, VP, VP, VP, VP, C.c_int, C.c_int]
        L.xb200_transform_main.argtypes = [VP, VP, C.c_int64, VP, C.c_int64, C.c_int]
        L.xb200_analyze_picture.argtypes = [VP, VP]
        L.xb200_picture_fetch.argtypes = [VP, C.c_int32, VP, VP, VP, VP, VP]
        L.xb200_picture_maps.argtypes = [VP, C.c_int32, VP, VP, VP, VP]
        L.xb200_picture_adopt.argtypes = [VP, C.c_int32, VP]
        L.xb200_picture_log_enable.argtypes = [VP, C.c_int64, C.c_int64]
        L.xb200_picture_log.argtypes = [VP, C.c_int32, VP, VP, VP]
        L.xb200_chain_capacity.argtypes = [VP]
        L.xb200_chain_prof.argtypes = [VP, VP]
        L.xb200_chain_debug.argtypes = [VP, VP]
        L.xb200_chain_span_ms.argtypes = [VP, C.c_int]
        L.xb200_chain_span_ms.restype = C.c_double
        _lib = L
    return _lib


EXPORTS = ["xb200_picture_ready", "xb200_create", "xb200_destroy", "xb200_version", "xb200_launch_count", "xb200_pic_create", "xb200_pic_destroy",
           "xb200_pic_upload", "xb200_pic_upload_s16", "xb200_pic_download", "xb200_sad", "xb200_ssd", "xb200_satd",
           "xb200_me", "xb200_mc", "xb200_bi_org", "xb200_fwd_dct_tc", "xb200_mvp", "xb200_tq", "xb200_itdq", "xb200_recon", "xb200_residue", "xb200_last_kernel_ms",
           "xb200_rdo_bits", "xb200_rdoq_rates", "xb200_analyze_cu", "xb200_deblock", "xb200_analyze_intra", "xb200_intra_nbr",
           "xb200_transform_main", "xb200_analyze_picture", "xb200_picture_fetch", "xb200_picture_maps", "xb200_picture_adopt",
           "xb200_picture_log_enable", "xb200_picture_log", "xb200_chain_capacity", "xb200_chain_prof", "xb200_chain_span_ms", "xb200_chain_debug"]


def _p(a):
    if a is None:
        return None
    if isinstance(a, int):
        return VP(a)
    return a.ctypes.data_as(VP)


def make_seq(w, h, preset="fast", bit_depth=10, **kw):
    """Sequence constants of a Baseline encode (reference presets, src_base/xeve_enc.c:2431-2531)."""
    s = np.zeros(1, SEQ)
    pre = {"fast": dict(me_range=32, hpel_cnt=2, merge_num=2), "medium": dict(me_range=64, hpel_cnt=4, merge_num=3)}[preset]
    s["w"], s["h"], s["bit_depth"] = w, h, bit_depth
    s["me_level"], s["hpel_cnt"], s["qpel_cnt"], s["me_complexity"] = 2, pre["hpel_cnt"], pre["hpel_cnt"], 1
    s["min_clip"], s["max_clip"] = [-127, -127], [w - 1, h - 1]
    s["merge_num"], s["me_range"], s["gop_size"], s["rdoq"], s["tool_iqt"] = pre["merge_num"], pre["me_range"], 16, 1, 0
    for k, v in kw.items():
        s[k] = v
    return s


class Hotpath:
    """One encoder instance's device context (xb200_ctx)."""

    def __init__(self, seq: np.ndarray, device: int = 0):
        self.L = load()
        self.seq = np.ascontiguousarray(seq).copy()
        h = VP()
        r = self.L.xb200_create(C.byref(h), device, _p(self.seq))
        if r != OK:
            raise Xb200Error(r, "xb200_create")
        self.h = h
        self.w, self.hgt = int(self.seq["w"][0]), int(self.seq["h"][0])

    def close(self):
        if getattr(self, "h", None):
            self.L.xb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, r, what):
        if r != OK:
            raise Xb200Error(r, what)

    @property
    def launches(self):
        return int(self.L.xb200_launch_count(self.h))

    @property
    def last_kernel_ms(self):
        return float(self.L.xb200_last_kernel_ms(self.h))

    # ---- pictures ----------------------------------------------------------------------------
    def pic_create(self, padded: bool) -> int:
        hd = C.c_int32(-1)
        self._ck(self.L.xb200_pic_create(self.h, int(padded), C.byref(hd)), "xb200_pic_create")
        return hd.value

    def pic_destroy(self, handle):
        self._ck(self.L.xb200_pic_destroy(self.h, handle), "xb200_pic_destroy")

    def pic_upload(self, handle, y, u, v, in_bit_depth):
        planes = (VP * 3)(_p(y), _p(u), _p(v))
        strides = (C.c_int32 * 3)(y.strides[0], u.strides[0], v.strides[0])
        self._ck(self.L.xb200_pic_upload(self.h, handle, planes, strides, in_bit_depth, MEM_HOST), "xb200_pic_upload")

    def pic_upload_s16(self, handle, y, u, v):
        planes = (VP * 3)(_p(y), _p(u), _p(v))
        strides = (C.c_int32 * 3)(y.strides[0] // 2, u.strides[0] // 2, v.strides[0] // 2)
        self._ck(self.L.xb200_pic_upload_s16(self.h, handle, planes, strides, MEM_HOST), "xb200_pic_upload_s16")

    def pic_download(self, handle, with_padding: bool):
        pl, pc = (PAD_L, PAD_C) if with_padding else (0, 0)
        y = np.empty((self.hgt + 2 * pl, self.w + 2 * pl), np.int16)
        u = np.empty((self.hgt // 2 + 2 * pc, self.w // 2 + 2 * pc), np.int16)
        v = np.empty_like(u)
        planes = (VP * 3)(_p(y), _p(u), _p(v))
        strides = (C.c_int32 * 3)(y.shape[1], u.shape[1], v.shape[1])
        self._ck(self.L.xb200_pic_download(self.h, handle, int(with_padding), planes, strides), "xb200_pic_download")
        return y, u, v

    # ---- probes ------------------------------------------------------------------------------
    def _probe(self, fn, items, dtype, what):
        items = np.ascontiguousarray(items, BLK_ITEM)
        out = np.zeros(len(items), dtype)
        self._ck(fn(self.h, _p(items), len(items), _p(out), MEM_HOST), what)
        return out

    def sad(self, items):
        return self._probe(self.L.xb200_sad, items, np.int32, "xb200_sad")

    def ssd(self, items):
        return self._probe(self.L.xb200_ssd, items, np.int64, "xb200_ssd")

    def satd(self, items):
        return self._probe(self.L.xb200_satd, items, np.int32, "xb200_satd")

    # ---- operators (host buffers) -----------------------------------------------------------
    def me(self, items, side=None):
        items = np.ascontiguousarray(items, ME_ITEM).copy()
        n_side = 0 if side is None else len(side)
        self._ck(self.L.xb200_me(self.h, _p(items), len(items), _p(side), n_side, MEM_HOST), "xb200_me")
        return items

    def mc(self, items, pred_off, total):
        items = np.ascontiguousarray(items, MC_ITEM)
        pred_off = np.ascontiguousarray(pred_off, np.int64)
        pred = np.zeros(total, np.int16)
        self._ck(self.L.xb200_mc(self.h, _p(items), len(items), _p(pred_off), _p(pred), total, MEM_HOST), "xb200_mc")
        return pred

    def bi_org(self, items, cur_pic, off, total):
        items = np.ascontiguousarray(items, MC_ITEM)
        cur_pic = np.ascontiguousarray(cur_pic, np.int32)
        off = np.ascontiguousarray(off, np.int64)
        side = np.zeros(total, np.int16)
        self._ck(self.L.xb200_bi_org(self.h, _p(items), len(items), _p(cur_pic), _p(off), _p(side), total, MEM_HOST), "xb200_bi_org")
        return side

    def fwd_dct_tc(self, blocks, log2n):
        blocks = np.ascontiguousarray(blocks, np.int16)
        out = np.zeros_like(blocks)
        n = blocks.size >> (2 * log2n)
        self._ck(self.L.xb200_fwd_dct_tc(self.h, _p(blocks), _p(out), n, log2n), "xb200_fwd_dct_tc")
        return out

    def mvp(self, items, pic, map_scu, map_mv, col0, col1):
        items = np.ascontiguousarray(items, MVP_ITEM).copy()
        self._ck(self.L.xb200_mvp(self.h, _p(items), len(items), _p(np.ascontiguousarray(pic, MVP_PIC)),
                                  _p(np.ascontiguousarray(map_scu, np.uint32)), _p(np.ascontiguousarray(map_mv, np.int16)),
                                  _p(np.ascontiguousarray(col0, np.int16)), _p(np.ascontiguousarray(col1, np.int16))), "xb200_mvp")
        return items

    def analyze_intra(self, items, rates, states, side, elems, want_rec=True):
        """pintra_analyze_cu over a CU list -> (items with results, states with the state_out slots written, coef, rec)"""
        items = np.ascontiguousarray(items, INTRA_ITEM).copy()
        rates = np.ascontiguousarray(rates, RATES)
        states = np.ascontiguousarray(states, SBAC).copy()
        side = np.ascontiguousarray(side, np.int16)
        coef = np.zeros(elems, np.int16)
        rec = np.zeros(elems, np.int16) if want_rec else None
        self._ck(self.L.xb200_analyze_intra(self.h, _p(items), len(items), _p(rates), len(rates), _p(states), len(states), _p(side),
                                            len(side), _p(coef), _p(rec), elems), "xb200_analyze_intra")
        return items, states, coef, rec

    def intra_nbr(self, handle, items, map_scu, map_ipm, w_scu, h_scu, cip, side_elems):
        """xeve_get_avail_intra + xeve_get_nbr + xeve_get_mpm of a CU list from device picture `handle` -> (items, side)"""
        items = np.ascontiguousarray(items, NBR_ITEM).copy()
        side = np.zeros(side_elems, np.int16)
        self._ck(self.L.xb200_intra_nbr(self.h, handle, _p(items), len(items), _p(np.ascontiguousarray(map_scu, np.uint32)),
                                        _p(np.ascontiguousarray(map_ipm, np.int8)), w_scu, h_scu, int(cip), _p(side), side_elems),
                 "xb200_intra_nbr")
        return items, side

    def transform_main(self, items, blocks):
        """Main-profile two-stage 16-bit transforms (IQT DCT-II / ATS), forward or inverse per item, on a copy of `blocks`"""
        items = np.ascontiguousarray(items, TRM_ITEM)
        blocks = np.ascontiguousarray(blocks, np.int16).copy()
        self._ck(self.L.xb200_transform_main(self.h, _p(items), len(items), _p(blocks), len(blocks), MEM_HOST), "xb200_transform_main")
        return blocks

    def deblock(self, handle, cus, pp, map_scu, map_refi, map_mv, expand=True):
        """xeve_loop_filter (+ xeve_picbuf_expand) on the device picture `handle`, in place"""
        cus = np.ascontiguousarray(cus, DF_CU)
        pp = np.ascontiguousarray(pp, DF_PIC).reshape(1)
        self._ck(self.L.xb200_deblock(self.h, handle, _p(cus), len(cus), _p(pp), _p(np.ascontiguousarray(map_scu, np.uint32)),
                                      _p(np.ascontiguousarray(map_refi, np.int8)), _p(np.ascontiguousarray(map_mv, np.int16)),
                                      int(expand), MEM_HOST), "xb200_deblock")

    def rdo_bits(self, items, states, coef=None):
        """-> (items with .bits, states with the state_out slots written)"""
        items = np.ascontiguousarray(items, BITS_ITEM).copy()
        states = np.ascontiguousarray(states, SBAC).copy()
        coef = np.zeros(1, np.int16) if coef is None else np.ascontiguousarray(coef, np.int16)
        self._ck(self.L.xb200_rdo_bits(self.h, _p(items), len(items), _p(states), len(states), _p(coef), len(coef)), "xb200_rdo_bits")
        return items, states

    def rdoq_rates(self, states):
        states = np.ascontiguousarray(states, SBAC)
        out = np.zeros(len(states), RATES)
        self._ck(self.L.xb200_rdoq_rates(self.h, _p(states), len(states), _p(out)), "xb200_rdoq_rates")
        return out

    def analyze_cu(self, items, rates, states, elems, want_rec=True):
        """xeve_pinter_analyze_cu over a CU list -> (items with results, states with s_next_best slots, coef, rec)"""
        items = np.ascontiguousarray(items, CU_ITEM).copy()
        rates = np.ascontiguousarray(rates, RATES)
        states = np.ascontiguousarray(states, SBAC).copy()
        coef = np.zeros(elems, np.int16)
        rec = np.zeros(elems, np.int16) if want_rec else None
        self._ck(self.L.xb200_analyze_cu(self.h, _p(items), len(items), _p(rates), len(rates), _p(states), len(states), _p(coef),
                                         _p(rec), elems), "xb200_analyze_cu")
        return items, states, coef, rec

    def tq(self, items, rates, coef):
        items = np.ascontiguousarray(items, TQ_ITEM).copy()
        rates = np.ascontiguousarray(rates, RATES)
        coef = np.ascontiguousarray(coef, np.int16).copy()
        self._ck(self.L.xb200_tq(self.h, _p(items), len(items), _p(rates), len(rates), _p(coef), len(coef), MEM_HOST), "xb200_tq")
        return items, coef

    def itdq(self, items, coef):
        items = np.ascontiguousarray(items, TQ_ITEM)
        coef = np.ascontiguousarray(coef, np.int16).copy()
        self._ck(self.L.xb200_itdq(self.h, _p(items), len(items), _p(coef), len(coef), MEM_HOST), "xb200_itdq")
        return coef

    def recon(self, items, resi, pred):
        items = np.ascontiguousarray(items, TQ_ITEM)
        rec = np.zeros(len(resi), np.int16)
        self._ck(self.L.xb200_recon(self.h, _p(items), len(items), _p(resi), _p(pred), _p(rec), len(resi), MEM_HOST), "xb200_recon")
        return rec

    def residue(self, items, rates, elems, want_rec=True):
        items = np.ascontiguousarray(items, RESIDUE_ITEM).copy()
        rates = np.ascontiguousarray(rates, RATES)
        coef = np.zeros(elems, np.int16)
        rec = np.zeros(elems, np.int16) if want_rec else None
        self._ck(self.L.xb200_residue(self.h, _p(items), len(items), _p(rates), len(rates), _p(coef), _p(rec), elems, MEM_HOST),
                 "xb200_residue")
        return items, coef, rec

    # ---- the decision pass of a whole picture (persistent kernel, one CTA per coder-state chain) ----------------------------
    @property
    def n_lcu(self):
        return ((self.w + 63) // 64) * ((self.hgt + 63) // 64)

    @property
    def f_scu(self):
        return ((self.w + 3) // 4) * ((self.hgt + 3) // 4)

    def chain_capacity(self):
        r = self.L.xb200_chain_capacity(self.h)
        if r < 0:
            raise Xb200Error(r, "xb200_chain_capacity")
        return r

    def chain_span_ms(self, reset=False):
        """device time from the first enqueue after the last reset to the latest completion among the pictures fetched since"""
        return float(self.L.xb200_chain_span_ms(self.h, int(reset)))

    def chain_debug(self):
        """progress words of the decision kernel (debug builds only, else None); never blocks"""
        out = np.zeros(64, np.int32)
        return out if self.L.xb200_chain_debug(self.h, _p(out)) == OK else None

    def chain_prof(self):
        """(cycles[32], counts[32]) per phase of the decision kernel -- profiling builds only, else None"""
        out = np.zeros(64, np.uint64)
        r = self.L.xb200_chain_prof(self.h, _p(out))
        return None if r != OK else (out[:32].copy(), out[32:].copy())

    def analyze_picture(self, pic):
        """Enqueue one picture (PICTURE record); returns at once, xb200 orders it after its reference pictures."""
        pic = np.ascontiguousarray(pic, PICTURE).reshape(1)
        self._ck(self.L.xb200_analyze_picture(self.h, _p(pic)), "xb200_analyze_picture")

    def picture_fetch(self, rec_pic, want_states=True):
        """Wait for picture rec_pic -> dict(scu [n_lcu, 256], coef [n_lcu, 6144], states [n_lcu, 2], cost [n_lcu], stat)"""
        n = self.n_lcu
        scu, coef = np.zeros((n, 256), SCU_REC), np.zeros((n, 6144), np.int16)
        st = np.zeros((n, 2), STATE) if want_states else None
        cost, stat = np.zeros(n, np.float64), np.zeros(1, PICTURE_STAT)
        self._ck(self.L.xb200_picture_fetch(self.h, rec_pic, _p(scu), _p(coef), _p(st), _p(cost), _p(stat)), "xb200_picture_fetch")
        return dict(scu=scu, coef=coef, states=st, cost=cost, stat=stat[0])

    def picture_wait(self, rec_pic):
        """Wait for picture rec_pic without copying its records (they are released) -> its PICTURE_STAT"""
        stat = np.zeros(1, PICTURE_STAT)
        self._ck(self.L.xb200_picture_fetch(self.h, rec_pic, None, None, None, None, _p(stat)), "xb200_picture_fetch")
        return stat[0]

    def picture_maps(self, rec_pic):
        f = self.f_scu
        m = dict(map_scu=np.zeros(f, np.uint32), map_ipm=np.zeros(f, np.int8), map_refi=np.zeros((f, 2), np.int8),
                 map_mv=np.zeros((f, 2, 2), np.int16))
        self._ck(self.L.xb200_picture_maps(self.h, rec_pic, _p(m["map_scu"]), _p(m["map_ipm"]), _p(m["map_refi"]), _p(m["map_mv"])),
                 "xb200_picture_maps")
        return m

    def picture_adopt(self, pic, map_mv):
        map_mv = np.ascontiguousarray(map_mv, np.int16)
        assert map_mv.size == self.f_scu * 4
        self._ck(self.L.xb200_picture_adopt(self.h, pic, _p(map_mv)), "xb200_picture_adopt")

    def picture_log_enable(self, cap_cu, cap_intra):
        self._ck(self.L.xb200_picture_log_enable(self.h, cap_cu, cap_intra), "xb200_picture_log_enable")
        self._log_caps = (cap_cu, cap_intra)

    def picture_log(self, rec_pic):
        """(inter CU records, intra CU records) of picture rec_pic in call order -- before picture_fetch releases them"""
        cu, it, n = np.zeros(self._log_caps[0], CU_ITEM), np.zeros(self._log_caps[1], INTRA_ITEM), np.zeros(2, np.int64)
        self._ck(self.L.xb200_picture_log(self.h, rec_pic, _p(cu), _p(it), _p(n)), "xb200_picture_log")
        return cu[:int(n[0])], it[:int(n[1])]
