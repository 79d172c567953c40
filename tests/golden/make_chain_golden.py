"""Regenerates tests/golden/chain_golden.npz (needs oracle/_ref, i.e. this container): the first three pictures in coding order
(POC 0 intra, POC 16 and POC 8 bi-predicted) of a 128x96 clip encoded by the UNMODIFIED reference (fast preset, default GOP), with
what its decision pass produced: per-CTU coder states, frame maps, leaf CUs and the deblocked pictures.  The fixture holds the
original pictures and picture-level parameters only as inputs -- tests re-encode the three pictures with the oracle alone."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tracedata  # noqa: E402

seq, pics = tracedata.live_chain("cif", 17, "fast", "", w=128, h=96, squares=[(24, 16, 28, 3, 2)])
pics = pics[:3]
assert [int(p["pp"]["poc"]) for p in pics] == [0, 16, 8]
tracedata.chain_sequence(seq, pics)        # the oracle reproduces them (asserts inside)
out = dict(n=len(pics), seq=seq)
for i, p in enumerate(pics):
    e = p["expect"]
    out.update({f"pp{i}": np.array(p["pp"]).reshape(1), f"df_pp{i}": np.array(p["df_pp"]).reshape(1), f"state_in{i}": e["state_in"],
                f"state_out{i}": e["state_out"], f"map_scu{i}": e["map_scu"], f"map_refi{i}": e["map_refi"], f"map_mv{i}": e["map_mv"],
                f"cus{i}": e["cus"]})
    for k, a, b in zip("yuv", p["org"], e["post"]):
        out[f"org_{k}{i}"], out[f"post_{k}{i}"] = a, b
np.savez_compressed(tracedata.CHAIN_GOLDEN, **out)
print(os.path.getsize(tracedata.CHAIN_GOLDEN), "bytes")
