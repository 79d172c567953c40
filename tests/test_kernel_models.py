"""Index-arithmetic models of kernels that have not run on hardware yet.

k_transform_main (xeve_b200/csrc/xb200_main.cu) was written after the round's GPU budget was spent.  Until its GPU test
(tests/test_zz_gpu_main_profile.py) has run, this is the evidence for its index mapping: a line-by-line Python transcription of the
kernel -- the same staging of the matrices as padded int8 rows, the same thread -> element mappings of the four stages, the same
shifts, truncation and saturation -- executed "one thread after the other" per stage and compared with the oracle that is pinned to the
Main-profile reference.  It says nothing about races or performance."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import transform_main_on_device as tm  # noqa: E402
from oracle import oracle as xo  # noqa: E402
from xeve_b200 import api  # noqa: E402


def _tables():
    """the 64-point DCT-II matrix from the generator the library uses (xb200_tables.h, compiled as C) and the eight ATS matrices"""
    import subprocess
    import tempfile
    d = tempfile.mkdtemp()
    src = os.path.join(d, "g.c")
    open(src, "w").write('#include <stdio.h>\n#include <stdint.h>\n#include "xeve_b200/csrc/xb200_tables.h"\n'
                         'int main(void){ static int8_t t[4096]; xb200_gen_tm64(t); fwrite(t,1,4096,stdout); return 0; }\n')
    subprocess.check_call(["gcc", "-O1", "-I", ROOT, "-o", os.path.join(d, "g"), src, "-lm"])
    tm64 = np.frombuffer(subprocess.check_output([os.path.join(d, "g")]), np.int8).astype(np.int64)
    L = xo.lib()
    ats = np.zeros((2, 4, 1024), np.int64)
    for ty in range(2):
        for l2 in range(2, 6):
            m = np.zeros(1 << (2 * l2), np.int8)
            L.xo_ats_matrix(ty, l2, m.ctypes.data_as(C.c_void_p))
            ats[ty, l2 - 2, :1 << (2 * l2)] = m
    return tm64, ats


def _model(it, blk, bd, tm64, ats):
    lw, lh = int(it["log2_w"]), int(it["log2_h"])
    w, h = 1 << lw, 1 << lh
    n, pw, ph = w * h, w + 4, h + 4

    def trm_matrix(log2n, dct8):
        nn = 1 << log2n
        if it["ats"]:
            return dict(src=ats[0 if dct8 else 1, log2n - 2], stride=nn, kstep=1, kmax=nn)
        return dict(src=tm64, stride=64, kstep=64 >> log2n, kmax=32 if nn == 64 else nn)

    def stage_matrix(t, nn):                       # trm_stage_matrix: rows of nn + 4 bytes, only kmax rows exist
        dst = [None] * (64 * 68)
        for e in range(t["kmax"] * nn):
            k = e // nn
            x = e - k * nn
            dst[k * (nn + 4) + x] = int(t["src"][(k * t["kstep"]) * t["stride"] + x])
        return dst
    s16 = lambda v: ((int(v) + 32768) & 0xffff) - 32768          # noqa: E731  (int16_t) store
    sat = lambda v: max(-32768, min(32767, int(v)))               # noqa: E731
    mw, mh = trm_matrix(lw, int(it["tridx"]) >> 1), trm_matrix(lh, int(it["tridx"]) & 1)
    s_mw, s_mh = stage_matrix(mw, w), stage_matrix(mh, h)
    s_a, s_b = [int(v) for v in blk], [0] * n
    if not it["inverse"]:
        sh1, sh2 = lw - 1 + bd - 8, lh + 6
        add1, add2 = (1 << (sh1 - 1)) if sh1 else 0, 1 << (sh2 - 1)
        for e in range(n):
            k, j, acc = e & (w - 1), e >> lw, 0
            if k < mw["kmax"]:
                acc = (sum(s_mw[k * pw + x] * s_a[j * w + x] for x in range(w)) + add1) >> sh1
            s_b[e] = s16(acc)
        out = [0] * n
        for e in range(n):
            j, k, acc = e & (w - 1), e >> lw, 0
            if k < mh["kmax"]:
                acc = (sum(s_mh[k * ph + y] * s_b[y * w + j] for y in range(h)) + add2) >> sh2
            out[e] = s16(acc)
        return out
    sh2 = 12 - (bd - 8)
    for e in range(n):
        j, y = e & (w - 1), e >> lw
        s_b[e] = sat((sum(s_mh[k * ph + y] * s_a[k * w + j] for k in range(mh["kmax"])) + 64) >> 7)
    out = [0] * n
    for e in range(n):
        x, y = e & (w - 1), e >> lw
        out[e] = sat((sum(s_mw[k * pw + x] * s_b[y * w + k] for k in range(mw["kmax"])) + (1 << (sh2 - 1))) >> sh2)
    return out


def test_transform_main_index_model_matches_oracle():
    tm64, ats = _tables()
    rng = np.random.default_rng(3)
    checked = 0
    for bd in (8, 10):
        items, blocks = tm.work_list(bd, rng, 1)
        exp = tm.expected(items, blocks, bd)
        for it in items:
            bits = int(it["log2_w"]) + int(it["log2_h"])
            big = 6 in (int(it["log2_w"]), int(it["log2_h"]))
            if bits > 8 and not (big and bits <= 10 and bd == 10):      # every code path, the large shapes once (pure-Python loops)
                continue
            n, off = 1 << bits, int(it["off"])
            assert _model(it, blocks[off:off + n], bd, tm64, ats) == list(exp[off:off + n]), (bd, it)
            checked += 1
    assert checked > 200
    assert api.TRM_ITEM.itemsize == 16
