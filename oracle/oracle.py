"""ctypes binding of oracle/liboracle.so (the scalar C restatement) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")


class PLANES(C.Structure):
    _fields_ = [("y", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p),
                ("s_l", C.c_int32), ("s_c", C.c_int32), ("w_l", C.c_int32), ("h_l", C.c_int32), ("poc", C.c_int32)]


_lib = None
VP = C.c_void_p
I = C.c_int


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, os.path.join(_HERE, "liboracle.so")])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.xo_sad.restype = I
        L.xo_sad.argtypes = [I, I, VP, I, VP, I, I]
        L.xo_ssd.restype = C.c_int64
        L.xo_ssd.argtypes = [I, I, VP, I, VP, I, I]
        L.xo_diff.restype = None
        L.xo_diff.argtypes = [I, I, VP, I, VP, I, VP, I]
        L.xo_satd.restype = I
        L.xo_satd.argtypes = [I, I, VP, I, VP, I, I]
        for f in (L.xo_mc_luma, L.xo_mc_chroma):
            f.restype = None
            f.argtypes = [VP, I, I, I, I, I, VP, I, I, I, I]
        for f in (L.xo_fwd_transform, L.xo_inv_transform):
            f.restype = None
            f.argtypes = [VP, I, I, I]
        L.xo_recon.restype = None
        L.xo_recon.argtypes = [VP, VP, I, I, VP, I]
        L.xo_pad_plane.restype = None
        L.xo_pad_plane.argtypes = [VP, I, I, I, I]
        L.xo_me_batch.restype = None
        L.xo_me_batch.argtypes = [VP, VP, VP, VP, C.c_int64]
        L.xo_mc_batch.restype = None
        L.xo_mc_batch.argtypes = [VP, VP, VP, C.c_int64, VP, VP]
        L.xo_analyze_cu_batch.restype = None
        L.xo_analyze_cu_batch.argtypes = [VP, VP, VP, VP, C.c_int64, VP, VP, VP]
        L.xo_rdo_bits_batch.restype = None
        L.xo_rdo_bits_batch.argtypes = [VP, C.c_int64, VP, VP]
        L.xo_rdoq_rates.restype = None
        L.xo_rdoq_rates.argtypes = [VP, C.c_int64, VP]
        L.xo_bi_org_batch.restype = None
        L.xo_bi_org_batch.argtypes = [VP, VP, VP, C.c_int64, VP, VP, VP]
        L.xo_tq_batch.restype = None
        L.xo_tq_batch.argtypes = [VP, VP, C.c_int64, VP, VP, VP]
        L.xo_residue_batch.restype = None
        L.xo_residue_batch.argtypes = [VP, VP, VP, VP, C.c_int64, VP, VP]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(VP) if a is not None else None


def me_batch(seq, planes, side, items):
    out = items.copy()
    lib().xo_me_batch(_p(seq), C.addressof(planes), _p(side), _p(out), len(out))
    return out


def mc_batch(seq, planes, items, off, total):
    pred = np.zeros(total, np.int16)
    lib().xo_mc_batch(_p(seq), C.addressof(planes), _p(np.ascontiguousarray(items)), len(items), _p(off), _p(pred))
    return pred


def bi_org_batch(seq, planes, items, cur_pic, off, total):
    side = np.zeros(total, np.int16)
    lib().xo_bi_org_batch(_p(seq), C.addressof(planes), _p(np.ascontiguousarray(items)), len(items),
                          _p(np.ascontiguousarray(cur_pic, np.int32)), _p(np.ascontiguousarray(off, np.int64)), _p(side))
    return side


def tq_batch(seq, items, rates, coef_in, want_itdq=True):
    items = items.copy()
    coef = coef_in.copy()
    resi = np.zeros_like(coef) if want_itdq else None
    lib().xo_tq_batch(_p(seq), _p(items), len(items), _p(rates), _p(coef), _p(resi))
    return items, coef, resi


def residue_batch(seq, planes, rates, items, elems):
    items = items.copy()
    coef = np.zeros(elems, np.int16)
    rec = np.zeros(elems, np.int16)
    lib().xo_residue_batch(_p(seq), C.addressof(planes), _p(rates), _p(items), len(items), _p(coef), _p(rec))
    return items, coef, rec


def rdo_bits_batch(items, states, coef):
    items, states = items.copy(), states.copy()
    coef = np.ascontiguousarray(coef, np.int16)
    lib().xo_rdo_bits_batch(_p(items), len(items), _p(states), _p(coef))
    return items, states


def rdoq_rates(states, rates_dtype):
    states = np.ascontiguousarray(states)
    out = np.zeros(len(states), rates_dtype)
    lib().xo_rdoq_rates(_p(states), len(states), _p(out))
    return out


def analyze_cu_batch(seq, planes, rates, items, states, elems):
    items, states = items.copy(), states.copy()
    coef = np.zeros(elems, np.int16)
    rec = np.zeros(elems, np.int16)
    lib().xo_analyze_cu_batch(_p(seq), C.addressof(planes), _p(np.ascontiguousarray(rates)), _p(items), len(items), _p(states),
                              _p(coef), _p(rec))
    return items, states, coef, rec


def hash_slots(buf, off, elems):
    """FNV-1a of buf[off[i] : off[i] + elems[i]] (s16) per item -- the hash the harness stores for in-situ outputs."""
    off = np.ascontiguousarray(off, np.int64)
    elems = np.ascontiguousarray(elems, np.int64)
    out = np.zeros(len(off), np.uint64)
    L = lib()
    L.xo_hash_slots.restype = None
    L.xo_hash_slots.argtypes = [VP, VP, VP, C.c_int64, VP]
    L.xo_hash_slots(_p(np.ascontiguousarray(buf, np.int16)), _p(off), _p(elems), len(off), _p(out))
    return out


def deblock(planes, cus, pp, map_scu, map_refi, map_mv, bit_depth=10):
    """xo_deblock on copies of the (Y, U, V) active-area arrays; returns the filtered planes."""
    L = lib()
    L.xo_deblock.restype = None
    L.xo_deblock.argtypes = [VP, VP, VP, I, I, I, I, VP, C.c_int64, VP, VP, VP, VP, I]
    out = [np.ascontiguousarray(a, np.int16).copy() for a in planes]
    cus, pp = np.ascontiguousarray(cus), np.ascontiguousarray(pp).reshape(1)
    L.xo_deblock(_p(out[0]), _p(out[1]), _p(out[2]), out[0].shape[1], out[1].shape[1], out[0].shape[1], out[0].shape[0], _p(cus),
                 len(cus), _p(pp), _p(np.ascontiguousarray(map_scu, np.uint32)), _p(np.ascontiguousarray(map_refi, np.int8)),
                 _p(np.ascontiguousarray(map_mv, np.int16)), bit_depth)
    return out


def analyze_intra_batch(seq, planes, rates, items, states, side, elems):
    items, states = items.copy(), states.copy()
    coef = np.zeros(elems, np.int16)
    rec = np.zeros(elems, np.int16)
    L = lib()
    L.xo_analyze_intra_batch.restype = None
    L.xo_analyze_intra_batch.argtypes = [VP, VP, VP, VP, C.c_int64, VP, VP, VP, VP]
    L.xo_analyze_intra_batch(_p(seq), C.addressof(planes), _p(np.ascontiguousarray(rates)), _p(items), len(items), _p(states),
                             _p(np.ascontiguousarray(side, np.int16)), _p(coef), _p(rec))
    return items, states, coef, rec


def intra_nbr(planes, items, map_scu, map_ipm, w_scu, h_scu, cip, side_elems, bit_depth=10):
    L = lib()
    L.xo_intra_nbr_batch.restype = None
    L.xo_intra_nbr_batch.argtypes = [VP, VP, VP, I, I, VP, C.c_int64, VP, VP, I, I, I, I, VP]
    keep = [np.ascontiguousarray(a, np.int16) for a in planes]
    items = items.copy()
    side = np.zeros(side_elems, np.int16)
    L.xo_intra_nbr_batch(_p(keep[0]), _p(keep[1]), _p(keep[2]), keep[0].shape[1], keep[1].shape[1], _p(items), len(items),
                         _p(np.ascontiguousarray(map_scu, np.uint32)), _p(np.ascontiguousarray(map_ipm, np.int8)), w_scu, h_scu, int(cip),
                         bit_depth, _p(side))
    return items, side


def mvp_batch(items, pic, map_scu, map_mv, col0, col1):
    L = lib()
    L.xo_mvp_batch.restype = None
    L.xo_mvp_batch.argtypes = [VP, C.c_int64, VP, VP, VP, VP, VP]
    items = items.copy()
    L.xo_mvp_batch(_p(items), len(items), _p(np.ascontiguousarray(pic)), _p(np.ascontiguousarray(map_scu, np.uint32)),
                   _p(np.ascontiguousarray(map_mv, np.int16)), _p(np.ascontiguousarray(col0, np.int16)), _p(np.ascontiguousarray(col1, np.int16)))
    return items


chain_current = None
SCU_REC = np.dtype([("mode", "u1"), ("log2", "u1"), ("ipm", "i1"), ("refi", "i1", (2,)), ("mvp_idx", "u1", (2,)), ("pad_", "u1"),
                    ("mv", "<i2", (2, 2)), ("mvd", "<i2", (2, 2)), ("nnz", "<i4", (3,))], align=True)


def chain_picture(seq, planes, pp, col0, col1, dtypes, ctu_limit=0):
    """xo_chain_picture: the CU decision chain of one picture (mode_coding_tree over every CTU).  pp: one CTU record (the
    harness's LCU_REC layout) carrying the picture-level inputs; dtypes = (LCU_REC, DF_CU, CU_ITEM, INTRA_ITEM).  Returns a
    dict: ctu (records with state_in / state_out), cost, rec (Y, U, V), map_scu, map_ipm, map_refi, map_mv, cus (leaf CUs in
    coding order), cu_log / intra_log (every inter / intra analysis in call order)."""
    lcu_dt, dfcu_dt, cu_dt, intra_dt = dtypes
    sq = np.ascontiguousarray(seq).reshape(-1)[0]
    w, h = int(sq["w"]), int(sq["h"])
    w_scu, h_scu = (w + 3) // 4, (h + 3) // 4
    f, n_ctu = w_scu * h_scu, ((w + 63) // 64) * ((h + 63) // 64)
    out = np.zeros(n_ctu, lcu_dt)
    cost = np.zeros(n_ctu, np.float64)
    rec = [np.zeros((h, w), np.int16), np.zeros((h // 2, w // 2), np.int16), np.zeros((h // 2, w // 2), np.int16)]
    map_scu, map_ipm = np.zeros(f, np.uint32), np.zeros(f, np.int8)
    map_refi, map_mv = np.zeros((f, 2), np.int8), np.zeros((f, 2, 2), np.int16)
    cap = f * 2
    cus, cu_log, intra_log = np.zeros(f, dfcu_dt), np.zeros(cap, cu_dt), np.zeros(cap, intra_dt)
    n_out = np.zeros(3, np.int64)
    pp = np.ascontiguousarray(pp).reshape(-1)[:1].copy()
    col0 = None if col0 is None else np.ascontiguousarray(col0, np.int16)
    col1 = None if col1 is None else np.ascontiguousarray(col1, np.int16)
    global chain_current   # what a stand-in for xo_mvp / xo_intra_nbr needs to see: the maps and the picture as they stand right now
    chain_current = dict(map_scu=map_scu, map_ipm=map_ipm, map_mv=map_mv, rec=rec, col0=col0, col1=col1 if col1 is not None else col0,
                         w_scu=w_scu, h_scu=h_scu, cip=int(pp["cip"][0]))
    L = lib()
    assert L.xo_sizeof_chain(0) == lcu_dt.itemsize, (L.xo_sizeof_chain(0), lcu_dt.itemsize)
    L.xo_chain_picture.restype = None
    assert L.xo_sizeof_chain(2) == SCU_REC.itemsize
    scu, coef = np.zeros((n_ctu, 256), SCU_REC), np.zeros((n_ctu, 6144), np.int16)
    L.xo_chain_picture.argtypes = [VP] * 10 + [I, I] + [VP] * 5 + [C.c_int64, VP, C.c_int64, VP, C.c_int64, VP, I, VP, VP]
    L.xo_chain_picture(_p(np.ascontiguousarray(seq)), C.addressof(planes), _p(pp), _p(col0), _p(col1), _p(out), _p(cost), _p(rec[0]),
                       _p(rec[1]), _p(rec[2]), w, w // 2, _p(map_scu), _p(map_ipm), _p(map_refi), _p(map_mv), _p(cus), len(cus),
                       _p(cu_log), cap, _p(intra_log), cap, _p(n_out), ctu_limit, _p(scu), _p(coef))
    assert n_out[1] <= cap and n_out[2] <= cap
    return dict(ctu=out, cost=cost, rec=rec, scu=scu, coef=coef, map_scu=map_scu, map_ipm=map_ipm, map_refi=map_refi, map_mv=map_mv,
                cus=cus[:int(n_out[0])], cu_log=cu_log[:int(n_out[1])], intra_log=intra_log[:int(n_out[2])])
