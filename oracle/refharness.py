"""ctypes binding of oracle/_ref/libref_harness.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It drives the UNMODIFIED reference (built by oracle/Makefile.ref from the
sources under /root/reference, outputs only into oracle/_ref/) through oracle/ref_harness.c:
kernel probes, call tracing of a real encode, and multi-threaded replay of work lists.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libref_harness.so")

NUM_CTX_CC_RUN = 24
NUM_CTX_CC_LEVEL = 24
NUM_CTX_CC_LAST = 2

ME_REC = np.dtype([
    ("poc", "<i4"), ("cur_pic", "<i4"), ("ref_pic", "<i4"), ("ref_poc", "<i4"),
    ("x", "<i2"), ("y", "<i2"),
    ("log2_cuw", "u1"), ("log2_cuh", "u1"), ("lidx", "u1"), ("bi", "u1"),
    ("refi", "i1"), ("num_refp", "u1"),
    ("mvp", "<i2", (2,)), ("mv_in", "<i2", (2,)),
    ("lambda_mv", "<u4"), ("mot_bits_in", "<i4", (2,)),
    ("max_search_range", "<i4"), ("gop_size", "<i4"), ("org_bi_off", "<i4"),
    ("mv_out", "<i2", (2,)), ("cost", "<u4"), ("mot_bits_out", "<i4", (2,)),
], align=True)

MC_REC = np.dtype([
    ("poc", "<i4"), ("ref_pic", "<i4", (2,)), ("ref_poc", "<i4", (2,)),
    ("x", "<i2"), ("y", "<i2"), ("w", "<i2"), ("h", "<i2"),
    ("refi", "i1", (2,)), ("mv", "<i2", (2, 2)),
    ("out_hash", "<u8"),
], align=True)

TQ_REC = np.dtype([
    ("poc", "<i4"),
    ("log2_cuw", "u1"), ("log2_cuh", "u1"), ("slice_type", "u1"), ("is_intra", "u1"),
    ("run_stats", "u1"), ("qp", "u1", (3,)),
    ("rate_idx", "<i4"), ("in_off", "<i8"),
    ("lambda", "<f8", (3,)), ("nnz", "<i4", (3,)),
    ("out_hash", "<u8"),
], align=True)

RATES = np.dtype([
    ("cbf_all", "<i4", (2,)), ("cbf_luma", "<i4", (2,)), ("cbf_cb", "<i4", (2,)), ("cbf_cr", "<i4", (2,)),
    ("run", "<i4", (NUM_CTX_CC_RUN, 2)), ("level", "<i4", (NUM_CTX_CC_LEVEL, 2)), ("last", "<i4", (NUM_CTX_CC_LAST, 2)),
], align=True)

PIC = np.dtype([
    ("poc", "<i4"), ("kind", "<i4"),
    ("w_l", "<i4"), ("h_l", "<i4"), ("w_c", "<i4"), ("h_c", "<i4"),
    ("s_l", "<i4"), ("s_c", "<i4"), ("pad_l", "<i4"), ("pad_c", "<i4"),
    ("off_y", "<i8"), ("off_u", "<i8"), ("off_v", "<i8"),
], align=True)

CONST = np.dtype([
    ("w", "<i4"), ("h", "<i4"), ("bit_depth", "<i4"), ("me_level", "<i4"), ("hpel_cnt", "<i4"), ("qpel_cnt", "<i4"),
    ("me_complexity", "<i4"), ("min_clip", "<i4", (2,)), ("max_clip", "<i4", (2,)),
    ("merge_num", "<i4"), ("me_range", "<i4"), ("gop_size", "<i4"), ("rdoq", "<i4"), ("tool_iqt", "<i4"),
], align=True)


class PLANES(C.Structure):
    _fields_ = [("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p),
                ("s_l", C.c_int32), ("s_c", C.c_int32), ("w_l", C.c_int32), ("h_l", C.c_int32), ("poc", C.c_int32)]


def available() -> bool:
    return os.path.exists(_LIB_PATH)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref is not built (run: make -C oracle -f Makefile.ref; needs /root/reference)")
        L = C.CDLL(_LIB_PATH)
        L.rh_encode_clip.restype = C.c_double
        L.rh_encode_clip.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        L.rh_trace_get.restype = C.c_int64
        L.rh_trace_get.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        L.rh_trace_const.argtypes = [C.c_void_p]
        L.rh_sizeof.restype = C.c_int
        L.rh_sad.restype = C.c_int
        L.rh_sad.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.rh_ssd.restype = C.c_int64
        L.rh_ssd.argtypes = L.rh_sad.argtypes
        L.rh_diff.restype = None
        L.rh_diff.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.rh_satd.restype = C.c_int
        L.rh_satd.argtypes = L.rh_sad.argtypes
        for f in (L.rh_mc_l, L.rh_mc_c):
            f.restype = None
            f.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        for f in (L.rh_fwd_transform, L.rh_inv_transform):
            f.restype = None
            f.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.rh_recon.restype = None
        L.rh_recon.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.rh_average.restype = None
        L.rh_average.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.rh_table.restype = C.c_void_p
        L.rh_table.argtypes = [C.c_int, C.POINTER(C.c_int)]
        L.rh_replay_me.restype = C.c_double
        L.rh_replay_me.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.rh_replay_mc.restype = C.c_double
        L.rh_replay_mc.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.rh_replay_tq.restype = C.c_double
        L.rh_replay_tq.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_int]
        for i, dt in enumerate((ME_REC, MC_REC, TQ_REC, RATES, PIC)):
            assert L.rh_sizeof(i) == dt.itemsize, (i, L.rh_sizeof(i), dt.itemsize)
        assert L.rh_sizeof(5) == C.sizeof(PLANES)
        assert L.rh_sizeof(6) == CONST.itemsize
        _lib = L
    return _lib


CU_REC = np.dtype([
    ("poc", "<i4"), ("cur_pic", "<i4"), ("x", "<i2"), ("y", "<i2"), ("log2_cuw", "u1"), ("log2_cuh", "u1"), ("slice_type", "u1"),
    ("ctx_skip", "u1"), ("ctx_pred_mode", "u1"), ("all_preds", "u1"), ("num_refp", "u1", (2,)), ("qp", "u1", (3,)), ("pad0_", "u1"),
    ("max_search_range", "<i4"), ("ref_pic", "<i4", (2, 4)), ("ref_poc", "<i4", (2, 4)), ("lambda_mv", "<u4"), ("rate_idx", "<i4"), ("state_in", "<i4"),
    ("state_out", "<i4"), ("lambda", "<f8", (3,)), ("dist_chroma_weight", "<f8", (2,)), ("mvp", "<i2", (2, 4, 2)),
    ("refi_pred", "i1", (2, 4)), ("mv_dir", "<i2", (2, 2)), ("out_off", "<i8"), ("cost", "<f8"), ("best_idx", "u1"), ("pad1_", "u1"),
    ("refi", "i1", (2,)), ("mvp_idx", "u1", (2,)), ("mv", "<i2", (2, 2)), ("mvd", "<i2", (2, 2)), ("nnz", "<i4", (3,)),
    ("coef_hash", "<u8"), ("rec_hash", "<u8"), ("me_first", "<i4"), ("me_cnt", "<i4"),
], align=True)
TRACE_CU, TRACE_CU_TIME = 8, 16


def cu_time():
    """(seconds inside the reference's xeve_pinter_analyze_cu, calls) of the last encode_clip(trace_mask=TRACE_CU_TIME)"""
    L = lib()
    L.rh_cu_time.restype = C.c_double
    n = C.c_int64(0)
    sec = L.rh_cu_time(C.byref(n))
    return sec, n.value
SBAC = np.dtype([("range", "<u4"), ("m", "<u2", (68,))], align=True)
BITS_REC = np.dtype([
    ("kind", "u1"), ("slice_type", "u1"), ("log2_cuw", "u1"), ("log2_cuh", "u1"), ("pidx", "u1"), ("ch", "u1"),
    ("ctx_skip", "u1"), ("ctx_pred_mode", "u1"), ("refi", "i1", (2,)), ("mvp_idx", "u1", (2,)), ("num_refp", "u1", (2,)),
    ("all_preds", "u1"), ("pad_", "u1"), ("mvd", "<i2", (2, 2)), ("nnz", "<i4", (3,)), ("state_in", "<i4"),
    ("state_out", "<i4"), ("coef_off", "<i8"), ("bits", "<u4"), ("pad2_", "<u4"),
], align=True)


def rdo_bits(items, states, coef):
    """The reference's xeve_rdo_bit_cnt_* (src_base/xeve_mode.c:57-302) on a scratch core."""
    L = lib()
    assert L.rh_sizeof_bits() == BITS_REC.itemsize
    items = np.ascontiguousarray(items, BITS_REC).copy()
    states = np.ascontiguousarray(states, SBAC).copy()
    coef = np.ascontiguousarray(coef, np.int16)
    L.rh_rdo_bits.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    assert L.rh_rdo_bits(_p(items), len(items), _p(states), _p(coef)) == 0
    return items, states


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


TRACE_ME, TRACE_MC, TRACE_TQ = 1, 2, 4
TRACE_DF = 32
TRACE_INTRA, TRACE_INTRA_TIME = 64, 128
TRACE_LCU = 256
TRACE_INJECT = 512
TRACE_PLAN = 1024    # the reference's control plane run dry: picture-level parameters of every picture, trivial decisions
TRACE_NO_LF = 2048   # ctx->fn_loop_filter does nothing (the deblocked picture lives on the device)
TRACE_ISOLATED = 4096  # threads = 1 only: private harness state for this call, so several encodes can run on several host threads


def intra_time():
    """(seconds inside the reference's pintra_analyze_cu, calls) of the last encode_clip(trace_mask=TRACE_INTRA_TIME)"""
    L = lib()
    L.rh_intra_time.restype = C.c_double
    n = C.c_int64(0)
    sec = L.rh_intra_time(C.byref(n))
    return sec, n.value


INTRA_REC = np.dtype([
    ("poc", "<i4"), ("cur_pic", "<i4"), ("x", "<i2"), ("y", "<i2"), ("log2_cuw", "u1"), ("log2_cuh", "u1"), ("slice_type", "u1"),
    ("ctx_skip", "u1"), ("ctx_pred_mode", "u1"), ("all_preds", "u1"), ("qp", "u1", (3,)), ("mpm", "u1", (5,)), ("pad0_", "u1", (2,)),
    ("inter_satd", "<u4"), ("rate_idx", "<i4"), ("state_in", "<i4"), ("state_out", "<i4"), ("cm_ipm_in", "<u2", (2,)),
    ("cm_ipm_out", "<u2", (2,)), ("lambda", "<f8", (3,)), ("sqrt_lambda0", "<f8"), ("dist_chroma_weight", "<f8", (2,)),
    ("nb_off", "<i8"), ("out_off", "<i8"), ("cost", "<f8"), ("dist_cu", "<i4"), ("ipm", "i1", (2,)), ("pad1_", "u1", (2,)),
    ("nnz", "<i4", (3,)), ("coef_hash", "<u8"), ("rec_hash", "<u8"),
], align=True)

# deblocking: layouts == xb200_df_cu / xb200_df_pic of include/xeve_b200.h
DF_CU = np.dtype([("x", "<i2"), ("y", "<i2"), ("log2_cuw", "u1"), ("log2_cuh", "u1"), ("pad_", "u1", (2,))], align=True)
DF_PIC = np.dtype([("w_scu", "<i4"), ("h_scu", "<i4"), ("qp_u_offset", "<i4"), ("qp_v_offset", "<i4"),
                   ("chroma_qp", "<i4", (2, 70))], align=True)
DF_REC = np.dtype([("poc", "<i4"), ("pre_pic", "<i4"), ("post_pic", "<i4"), ("on", "<i4"), ("cu_first", "<i8"), ("cu_cnt", "<i8"),
                   ("maps_off", "<i8"), ("pp", DF_PIC)], align=True)


# CTU decision trace (ctx->fn_mode_analyze_lcu): layouts == RH_STATE / RH_LCU_REC of ref_harness.c == xo_state / xo_chain_pic + states
STATE = np.dtype([("range", "<u4"), ("m", "<u2", (68,)), ("ipm", "<u2", (2,)), ("split", "<u2"), ("pad_", "<u2")], align=True)
LCU_REC = np.dtype([("poc", "<i4"), ("slice_type", "<i4"), ("lcu_num", "<i4"), ("x_pel", "<i4"), ("y_pel", "<i4"), ("tile_qp", "<i4"),
                    ("cur_pic", "<i4"), ("num_refp", "<i4", (2,)), ("ref_pic", "<i4", (2, 4)), ("ref_poc", "<i4", (2, 4)),
                    ("col_list_poc0", "<i4"), ("max_cu_inter", "<i4"), ("min_cu_inter", "<i4"), ("max_cu_intra", "<i4"),
                    ("min_cu_intra", "<i4"), ("cip", "<i4"), ("qp", "<i4", (3,)), ("lambda_mv", "<u4"), ("max_search_range", "<i4"),
                    ("parallel_rows", "<i4"), ("lambda", "<f8", (3,)), ("sqrt_lambda0", "<f8"), ("dist_chroma_weight", "<f8", (2,)),
                    ("col_off", "<i8", (2,)), ("state_in", STATE), ("state_out", STATE)], align=True)


NBR_REC = np.dtype([("x", "<i2"), ("y", "<i2"), ("log2_cuw", "u1"), ("log2_cuh", "u1"), ("mpm", "u1", (5,)), ("pad_", "u1"),
                    ("avail", "<u2"), ("pad2_", "<u2"), ("nb_off", "<i8")], align=True)


def intra_nbr(planes, items, map_scu, map_ipm, w_scu, h_scu, cip, side_elems, bit_depth=10):
    """The reference's xeve_get_avail_intra / xeve_get_nbr / xeve_get_mpm over a CU list -> (items, side)"""
    L = lib()
    assert L.rh_sizeof_nbr() == NBR_REC.itemsize
    L.rh_intra_nbr.restype = None
    L.rh_intra_nbr.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    pl = PLANES()
    keep = [np.ascontiguousarray(a, np.int16) for a in planes]
    pl.y, pl.u, pl.v = (a.ctypes.data for a in keep)
    pl.s_l, pl.s_c, pl.w_l, pl.h_l = keep[0].shape[1], keep[1].shape[1], keep[0].shape[1], keep[0].shape[0]
    items = np.ascontiguousarray(items, NBR_REC).copy()
    side = np.zeros(side_elems, np.int16)
    L.rh_intra_nbr(C.addressof(pl), _p(items), len(items), _p(np.ascontiguousarray(map_scu, np.uint32)),
                   _p(np.ascontiguousarray(map_ipm, np.int8)), w_scu, h_scu, int(cip), bit_depth, _p(side))
    return items, side


def df_maps(maps, off, f):
    """(map_scu u32[f], map_refi s8[f,2], map_mv s16[f,2,2]) of one traced picture (copies)."""
    b = maps[off:off + 14 * f]
    return (np.frombuffer(b[:4 * f].tobytes(), "<u4").copy(), np.frombuffer(b[4 * f:6 * f].tobytes(), "i1").reshape(f, 2).copy(),
            np.frombuffer(b[6 * f:14 * f].tobytes(), "<i2").reshape(f, 2, 2).copy())


def deblock(planes, cus, pp, map_scu, map_refi, map_mv, bit_depth=10):
    """The reference's xeve_deblock_cu_ver / _hor over a CU list (vertical edges of the whole picture, then horizontal),
    on copies of the (Y, U, V) active-area arrays; returns (filtered planes, seconds)."""
    L = lib()
    L.rh_deblock.restype = C.c_double
    L.rh_deblock.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    assert L.rh_sizeof_df(1) == DF_CU.itemsize and L.rh_sizeof_df(2) == DF_PIC.itemsize
    out = [np.ascontiguousarray(a, np.int16).copy() for a in planes]
    pl = PLANES()
    pl.y, pl.u, pl.v = (a.ctypes.data for a in out)
    pl.s_l, pl.s_c, pl.w_l, pl.h_l, pl.poc = out[0].shape[1], out[1].shape[1], out[0].shape[1], out[0].shape[0], 0
    cus = np.ascontiguousarray(cus, DF_CU)
    pp = np.ascontiguousarray(pp, DF_PIC).reshape(1)
    sec = L.rh_deblock(C.addressof(pl), _p(cus), len(cus), _p(pp), _p(np.ascontiguousarray(map_scu, np.uint32)),
                       _p(np.ascontiguousarray(map_refi, np.int8)), _p(np.ascontiguousarray(map_mv, np.int16)), bit_depth)
    return out, sec
PRESET = {"default": 0, "fast": 1, "medium": 2, "slow": 3, "placebo": 4}


class Trace:
    """Work lists recorded from one real reference encode (copied out of the harness)."""

    def __init__(self, me, mc, tq, rates, pics, samp, const, enc_seconds, bitstream):
        self.me, self.mc, self.tq, self.rates, self.pics, self.samp = me, mc, tq, rates, pics, samp
        self.const, self.enc_seconds, self.bitstream = const, enc_seconds, bitstream

    # -- picture helpers -------------------------------------------------------------
    def plane_views(self, idx):
        """(Y, U, V) full-buffer 2-D views (including padding) of picture-table entry idx."""
        p = self.pics[idx]
        out = []
        for off, s, hh, pad in ((p["off_y"], p["s_l"], p["h_l"], p["pad_l"]),
                                (p["off_u"], p["s_c"], p["h_c"], p["pad_c"]),
                                (p["off_v"], p["s_c"], p["h_c"], p["pad_c"])):
            rows = int(hh + 2 * pad)
            out.append(self.samp[int(off):int(off) + rows * int(s)].reshape(rows, int(s)))
        return out

    def planes_struct(self):
        """ctypes PLANES array addressing the active areas inside self.samp (keep self alive)."""
        arr = (PLANES * len(self.pics))()
        base = self.samp.ctypes.data
        for i, p in enumerate(self.pics):
            pl, pc = int(p["pad_l"]), int(p["pad_c"])
            arr[i].y = base + 2 * (int(p["off_y"]) + pl * int(p["s_l"]) + pl)
            arr[i].u = base + 2 * (int(p["off_u"]) + pc * int(p["s_c"]) + pc)
            arr[i].v = base + 2 * (int(p["off_v"]) + pc * int(p["s_c"]) + pc)
            arr[i].s_l, arr[i].s_c = int(p["s_l"]), int(p["s_c"])
            arr[i].w_l, arr[i].h_l, arr[i].poc = int(p["w_l"]), int(p["h_l"]), int(p["poc"])
        return arr


def encode_clip(yuv: np.ndarray, nframes, w, h, in_depth=8, preset="fast", qp=-1, threads=1, bframes=-1,
                extra="", trace_mask=0, pic_lo=0, pic_hi=1 << 30, want_bitstream=True) -> Trace:
    """Run the reference encoder over an in-memory I420 clip, optionally tracing hot-path calls."""
    L = lib()
    yuv = np.ascontiguousarray(yuv)
    cap = 32 << 20
    bs = np.empty(cap, np.uint8) if want_bitstream else None
    n = C.c_int64(0)
    sec = L.rh_encode_clip(_p(yuv), nframes, w, h, in_depth, PRESET[preset], qp, threads, bframes,
                           extra.encode(), trace_mask, pic_lo, pic_hi, _p(bs), cap, C.byref(n))
    if sec < 0:
        raise RuntimeError(f"reference encode failed ({sec})")
    if trace_mask & TRACE_ISOLATED:      # nothing was recorded in the shared state
        return Trace(None, None, None, None, None, None, None, sec, bs[: n.value].copy() if want_bitstream else None)

    def grab(what, dt):
        ptr = C.c_void_p()
        cnt = L.rh_trace_get(what, C.byref(ptr))
        if cnt == 0:
            return np.empty(0, dt)
        buf = (C.c_char * (cnt * dt.itemsize)).from_address(ptr.value)
        return np.frombuffer(buf, dtype=dt, count=cnt).copy()

    me, mc, tq = grab(0, ME_REC), grab(1, MC_REC), grab(2, TQ_REC)
    rates, pics, samp = grab(3, RATES), grab(4, PIC), grab(5, np.dtype("<i2"))
    sbac = grab(6, SBAC)
    assert L.rh_sizeof(7) == CU_REC.itemsize, (L.rh_sizeof(7), CU_REC.itemsize)
    cu, cu_sbac = grab(7, CU_REC), grab(8, SBAC)
    cst = np.zeros(1, CONST)
    L.rh_trace_const(_p(cst))
    tr = Trace(me, mc, tq, rates, pics, samp, cst, sec, bs[: n.value].copy() if want_bitstream else None)
    tr.sbac, tr.cu, tr.cu_sbac = sbac, cu, cu_sbac
    assert L.rh_sizeof_df(0) == DF_REC.itemsize, (L.rh_sizeof_df(0), DF_REC.itemsize)
    assert L.rh_sizeof_intra() == INTRA_REC.itemsize, (L.rh_sizeof_intra(), INTRA_REC.itemsize)
    tr.intra = grab(12, INTRA_REC)
    tr.df, tr.df_cu, tr.df_maps = grab(9, DF_REC), grab(10, DF_CU), grab(11, np.dtype("u1"))
    assert L.rh_sizeof_lcu() == LCU_REC.itemsize, (L.rh_sizeof_lcu(), LCU_REC.itemsize)
    tr.lcu = grab(13, LCU_REC)
    return tr


class INJECT_PIC(C.Structure):
    _fields_ = [("poc", C.c_int32), ("s_l", C.c_int32), ("s_c", C.c_int32), ("pad_", C.c_int32), ("scu", C.c_void_p), ("coef", C.c_void_p),
                ("rec_y", C.c_void_p), ("rec_u", C.c_void_p), ("rec_v", C.c_void_p)]


def encode_clip_injected(yuv, nframes, w, h, decisions, **kw):
    """Reference encode whose mode decision (ctx->fn_mode_analyze_lcu) is REPLACED by externally supplied decisions: `decisions` =
    per picture dict(poc, scu [n_ctu, 256] SCU records, coef [n_ctu, 6144] s16, rec (Y, U, V) before deblocking).  Everything
    else -- entropy coding, bitstream writing, loop filter, reference picture management -- is the unmodified reference.
    Returns (trace, CTUs injected, inter analyses run by the reference, intra analyses run by the reference)."""
    L = lib()
    assert L.rh_sizeof_inject(1) == C.sizeof(INJECT_PIC)
    arr = (INJECT_PIC * len(decisions))()
    keep = []
    for i, d in enumerate(decisions):
        scu, coef = np.ascontiguousarray(d["scu"]), np.ascontiguousarray(d["coef"], np.int16)
        rec = [np.ascontiguousarray(a, np.int16) for a in d["rec"]]
        assert scu.dtype.itemsize == L.rh_sizeof_inject(0)
        keep.append((scu, coef, rec))
        arr[i].poc, arr[i].s_l, arr[i].s_c = int(d["poc"]), rec[0].shape[1], rec[1].shape[1]
        arr[i].scu, arr[i].coef = scu.ctypes.data, coef.ctypes.data
        arr[i].rec_y, arr[i].rec_u, arr[i].rec_v = [a.ctypes.data for a in rec]
    L.rh_inject.restype = None
    L.rh_inject.argtypes = [C.c_void_p, C.c_int]
    L.rh_inject_count.restype = C.c_int64
    L.rh_inject(C.addressof(arr), len(decisions))
    try:
        tr = encode_clip(yuv, nframes, w, h, trace_mask=TRACE_INJECT | TRACE_CU_TIME | TRACE_INTRA_TIME, **kw)
        n = int(L.rh_inject_count())
    finally:
        L.rh_inject(None, 0)
    return tr, n, cu_time()[1], intra_time()[1]


def plan_clip(nframes, w, h, in_depth=8, preset="fast", qp=-1, threads=1, bframes=-1, extra=""):
    """The picture-level plan of a clip from the reference's OWN control plane run dry (RH_T_PLAN): per picture in coding order the
    LCU_REC of CTU 0 (slice type, POC, QPs, lambdas, reference POCs, CU size limits, search range, parallel rows) and the loop-filter
    parameters -- the inputs of xb200_analyze_picture.  With constant QP none of them depends on a decision, so the dry run (every CTU
    gets 8x8 SKIP / intra DC units) yields what a real encode of `nframes` frames computes; takes milliseconds per picture.
    Returns (seq constants, [dict(pp, df_pp, deblock)])."""
    bps = 2 if in_depth > 8 else 1
    yuv = np.zeros(nframes * w * h * 3 // 2 * bps, np.uint8)
    if in_depth <= 8:
        yuv[:] = 128
    tr = encode_clip(yuv, nframes, w, h, in_depth=in_depth, preset=preset, qp=qp, threads=threads, bframes=bframes, extra=extra,
                     trace_mask=TRACE_PLAN, want_bitstream=False)
    assert len(tr.lcu) == len(tr.df) == nframes, (len(tr.lcu), len(tr.df), nframes)
    return tr.const, [dict(pp=tr.lcu[i].copy(), df_pp=tr.df[i]["pp"].copy(), deblock=int(tr.df[i]["on"])) for i in range(nframes)]


_FETCH_CB = C.CFUNCTYPE(C.c_int, C.c_int, C.POINTER(INJECT_PIC))


def encode_clip_lazy(yuv, nframes, w, h, fetch, no_loop_filter=True, label_threads=1, **kw):
    """Reference encode whose mode decision is replaced by records handed over picture by picture: fetch(poc) -> dict(scu, coef[, rec])
    is called when the reference starts coding picture `poc` (ctx->fn_mode_analyze_frame) and may block until that picture is decided.
    no_loop_filter: the reference's own loop filter is skipped (the reconstruction stays with the decision engine; the bitstream does
    not depend on it).  label_threads: the `threads` value the decisions were made with -- the reference records it in the parameter SEI of
    the first access unit, so the single-threaded host pass labels the stream like the reference run it reproduces.
    Returns (trace with the bitstream, CTUs injected)."""
    L = lib()
    keep, err = [], []

    def cb(poc, out):
        try:
            d = fetch(int(poc))
            scu, coef = np.ascontiguousarray(d["scu"]), np.ascontiguousarray(d["coef"], np.int16)
            assert scu.dtype.itemsize == L.rh_sizeof_inject(0)
            keep[:] = [(scu, coef, d.get("rec"))]
            o = out.contents
            o.poc, o.scu, o.coef = int(poc), scu.ctypes.data, coef.ctypes.data
            o.rec_y = o.rec_u = o.rec_v = None
            if d.get("rec") is not None:
                rec = [np.ascontiguousarray(a, np.int16) for a in d["rec"]]
                keep.append(rec)
                o.s_l, o.s_c = rec[0].shape[1], rec[1].shape[1]
                o.rec_y, o.rec_u, o.rec_v = [a.ctypes.data for a in rec]
            return 0
        except Exception as e:  # noqa: BLE001 -- must not propagate through the C frame
            err.append(e)
            return -1
    cfn = _FETCH_CB(cb)
    L.rh_inject_lazy.restype = None
    L.rh_inject_lazy.argtypes = [C.c_void_p]
    L.rh_inject_lazy_count.restype = C.c_int64
    L.rh_inject_lazy(C.cast(cfn, C.c_void_p))
    L.rh_label_threads.restype = None
    L.rh_label_threads.argtypes = [C.c_int]
    L.rh_label_threads(int(label_threads))
    kw = dict(kw)
    kw["threads"] = 1              # the bitstream is written from the records alone; hooks must run in the calling thread
    try:
        tr = encode_clip(yuv, nframes, w, h, trace_mask=TRACE_INJECT | TRACE_ISOLATED | (TRACE_NO_LF if no_loop_filter else 0), **kw)
        n = int(L.rh_inject_lazy_count())
    finally:
        L.rh_inject_lazy(None)
        L.rh_label_threads(0)
    if err:
        raise err[0]
    return tr, n


def replay_me(tr: Trace, recs: np.ndarray | None = None, nthreads=1):
    """Reference pinter_me_epzs over a work list; returns (records with outputs filled, seconds)."""
    L = lib()
    recs = tr.me if recs is None else np.ascontiguousarray(recs)
    out = np.empty_like(recs)
    planes = tr.planes_struct()
    sec = L.rh_replay_me(_p(tr.const), C.addressof(planes), _p(tr.samp), _p(recs), _p(out), len(recs), nthreads)
    return out, sec


def mc_offsets(recs):
    sz = recs["w"].astype(np.int64) * recs["h"].astype(np.int64) * 3 // 2
    off = np.zeros(len(recs), np.int64)
    np.cumsum(sz[:-1], out=off[1:])
    return off, int(sz.sum())


def replay_mc(tr: Trace, recs=None, nthreads=1, want_pred=True):
    L = lib()
    recs = tr.mc if recs is None else np.ascontiguousarray(recs)
    off, total = mc_offsets(recs)
    pred = np.empty(total, np.int16) if want_pred else None
    hsh = np.empty(len(recs), np.uint64)
    planes = tr.planes_struct()
    sec = L.rh_replay_mc(_p(tr.const), C.addressof(planes), _p(recs), len(recs), _p(pred), _p(off), _p(hsh), nthreads)
    return pred, off, hsh, sec


def replay_tq(tr: Trace, recs=None, nthreads=1, want_itdq=True):
    """Reference xeve_sub_block_tq (+ xeve_itdq) over a work list.

    Returns (coef, nnz[n,3], resi or None, seconds); coef/resi share tr.samp's element offsets.
    """
    L = lib()
    recs = tr.tq if recs is None else np.ascontiguousarray(recs)
    coef = np.zeros_like(tr.samp)
    resi = np.zeros_like(tr.samp) if want_itdq else None
    nnz = np.zeros((len(recs), 3), np.int32)
    sec = L.rh_replay_tq(_p(tr.const), _p(tr.samp), _p(recs), _p(tr.rates), len(recs), _p(coef), _p(nnz), _p(resi),
                         nthreads)
    return coef, nnz, resi, sec


RES_REC = np.dtype([
    ("mc", MC_REC), ("cur_pic", "<i4"),
    ("slice_type", "u1"), ("run_stats", "u1"), ("qp", "u1", (3,)), ("pad_", "u1", (3,)),
    ("rate_idx", "<i4"), ("lambda", "<f8", (3,)), ("out_off", "<i8"),
    ("nnz", "<i4", (3,)), ("dist_pred", "<i8", (3,)), ("dist_rec", "<i8", (3,)),
], align=True)


def replay_residue(const, planes_struct, rates, items, elems, nthreads=1):
    """Reference xeve_mc -> diff -> SSD -> fn_tq -> fn_itdp -> recon -> SSD over residue items.

    planes_struct: ctypes PLANES array (picture-table order); returns (items, coef, rec, seconds).
    """
    L = lib()
    L.rh_replay_residue.restype = C.c_double
    L.rh_replay_residue.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    assert L.rh_sizeof_res() == RES_REC.itemsize
    items = np.ascontiguousarray(items).copy()
    coef = np.zeros(elems, np.int16)
    rec = np.zeros(elems, np.int16)
    sec = L.rh_replay_residue(_p(const), C.addressof(planes_struct), _p(np.ascontiguousarray(rates)), _p(items), len(items),
                              _p(coef), _p(rec), nthreads)
    return items, coef, rec, sec


def replay_me_raw(const, planes_struct, side, recs, nthreads=1):
    """rh_replay_me on caller-provided pictures (ctypes PLANES array) instead of a Trace."""
    L = lib()
    recs = np.ascontiguousarray(recs)
    out = np.empty_like(recs)
    sec = L.rh_replay_me(_p(const), C.addressof(planes_struct), _p(side), _p(recs), _p(out), len(recs), nthreads)
    return out, sec


def replay_mc_raw(const, planes_struct, recs, nthreads=1):
    L = lib()
    recs = np.ascontiguousarray(recs)
    off, total = mc_offsets(recs)
    pred = np.empty(total, np.int16)
    sec = L.rh_replay_mc(_p(const), C.addressof(planes_struct), _p(recs), len(recs), _p(pred), _p(off), None, nthreads)
    return pred, off, sec


def table(which: int, dtype) -> np.ndarray:
    L = lib()
    n = C.c_int(0)
    ptr = L.rh_table(which, C.byref(n))
    buf = (C.c_char * n.value).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()
