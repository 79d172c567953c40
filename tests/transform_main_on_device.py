"""xb200_transform_main (Main-profile two-stage 16-bit transforms) against the oracle (xo_iqt_* / xo_ats_*, themselves
pinned against the Main-profile reference in tests/test_oracle.py).  Run by tests/test_zz_gpu_main_profile.py in a subprocess."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as xo  # noqa: E402
from xeve_b200 import api  # noqa: E402

p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731


def work_list(bd, rng, reps):
    items, chunks, off = [], [], 0
    for ats in (0, 1):
        lo, hi = (2, 5) if ats else (1, 6)
        for lw in range(lo, hi + 1):
            for lh in range(max(lo, lw - 2), min(hi, lw + 2) + 1):
                for tridx in (range(4) if ats else (0,)):
                    for inverse in (0, 1):
                        for _ in range(reps):
                            n = 1 << (lw + lh)
                            amp = (1 << bd) - 1 if not inverse else 2000
                            blk = rng.integers(-amp, amp + 1, n).astype(np.int16)
                            if inverse:
                                blk[rng.random(n) < 0.7] = 0
                                if rng.random() < 0.2:
                                    blk[:] = rng.integers(-32768, 32768, n)          # saturating inputs
                            items.append((lw, lh, inverse, ats, tridx, (0, 0, 0), off))
                            chunks.append(blk)
                            off += n
    return np.array(items, api.TRM_ITEM), np.concatenate(chunks)


def expected(items, blocks, bd):
    L = xo.lib()
    out = blocks.copy()
    for it in items:
        n, off = 1 << (int(it["log2_w"]) + int(it["log2_h"])), int(it["off"])
        blk = np.ascontiguousarray(out[off:off + n])
        fn = {(0, 0): L.xo_iqt_fwd, (0, 1): L.xo_iqt_inv, (1, 0): L.xo_ats_fwd, (1, 1): L.xo_ats_inv}[(int(it["ats"]), int(it["inverse"]))]
        args = (p(blk), int(it["log2_w"]), int(it["log2_h"]), bd) + ((int(it["tridx"]),) if it["ats"] else ())
        fn(*args)
        out[off:off + n] = blk
    return out


def main():
    rng = np.random.default_rng(11)
    total = 0
    for bd, (w, h) in ((10, (352, 288)), (8, (352, 288))):
        seq = api.make_seq(w, h)
        seq["bit_depth"] = bd
        hp = api.Hotpath(seq)
        items, blocks = work_list(bd, rng, 6)
        perm = rng.permutation(len(items))                     # blocks need not be in offset order
        got = hp.transform_main(items[perm], blocks)
        exp = expected(items, blocks, bd)
        assert np.array_equal(got, exp), ("mismatch", bd, int(np.flatnonzero(got != exp)[0]))
        assert hp.launches >= 1
        total += len(items)
        bad = items[:3].copy()
        bad["log2_w"][1] = 7
        try:
            hp.transform_main(bad, blocks)
            raise SystemExit("a 128-point item was accepted")
        except api.Xb200Error as e:
            assert e.code == api.ERR_INVALID_ARGUMENT
        assert np.array_equal(hp.transform_main(items[:5], blocks), expected(items[:5], blocks, bd))   # the context still works
        assert len(hp.transform_main(np.zeros(0, api.TRM_ITEM), np.zeros(0, np.int16))) == 0
        hp.close()
    print(f"{total} blocks (IQT DCT-II 2..64, ATS DST-VII / DCT-VIII 4..32, forward and inverse, 8 and 10 bit) equal the oracle")
    print("TRANSFORM_MAIN_OK")


if __name__ == "__main__":
    main()
