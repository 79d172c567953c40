"""Generates tests/golden/rate_golden.npz from the compiled reference (oracle/_ref, this container).

Run:  python tests/golden/make_rate_golden.py
  bits / states : what the reference's xeve_rdo_bit_cnt_* + xeve_get_bit_number (src_base/xeve_mode.c:39-302) return
                  for the seeded work list tests/ratedata.work() (the list itself is regenerated from the seed)
  sbac / rates  : coder states of a traced qcif encode and the RDOQ rate tables xeve_rdoq_bit_est derived from them
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ratedata  # noqa: E402
import tracedata  # noqa: E402
from oracle import refharness as rh  # noqa: E402

it, st, coef = ratedata.work()
a, sa = rh.rdo_bits(it, st, coef)
td = tracedata.live_trace("cif", frames=20, pic_lo=1, pic_hi=2, mask=4, **tracedata.QCIF)
tr = td.live
pick = np.linspace(0, len(tr.sbac) - 1, 300).astype(int)
path = os.path.join(ROOT, "tests", "golden", "rate_golden.npz")
np.savez_compressed(path, bits=a["bits"], states=sa, sbac=tr.sbac[pick], rates=tr.rates[pick])
print("wrote", path, os.path.getsize(path) // 1024, "KiB")
