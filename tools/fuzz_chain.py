"""Randomized sweep of the decision-chain oracle against the compiled reference (needs oracle/_ref, i.e. this container).

  python tools/fuzz_chain.py chain  SEED SECONDS   # xo_chain_picture -> deblock -> pad over whole random sequences == the reference's
  python tools/fuzz_chain.py inject SEED SECONDS   # ... and the decisions injected into the unmodified reference: byte-identical bitstream

Random picture sizes (multiples of 8), QP 8..51, presets fast / medium, 0..15 B pictures, P slices, two references, plain quantiser,
quarter-pel search, four skip candidates, 8- / 10-bit input, threads 1..3 (chain mode).  Round 1: 574 + 200 configurations, no
difference."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tracedata  # noqa: E402


def main():
    mode, seed, seconds = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
    rng = np.random.default_rng(seed)
    t_end, ok, fails = time.time() + seconds, 0, 0
    while time.time() < t_end:
        w, h = int(rng.integers(16, 48)) * 8, int(rng.integers(8, 40)) * 8          # at least two CTUs wide (the
        qp, preset = int(rng.integers(8, 52)), ["fast", "medium"][int(rng.integers(0, 2))]   # reference races on one-CTU-wide pictures)
        frames = int(rng.integers(3, 9))
        extra = [f"qp={qp}", f"bframes={[0, 1, 3, 7, 15][int(rng.integers(0, 5))]}"]
        for prob, opt in ((0.3, "inter_slice_type=1"), (0.3, "ref=2;me_ref_num=2"), (0.2, "rdoq=0"), (0.2, "me_sub=3;me_sub_pos=8"),
                          (0.2, "merge_num=4")):
            if rng.random() < prob:
                extra.append(opt)
        threads = int(rng.integers(1, 4)) if mode == "chain" else 1
        clip = ["cif", "2160p10"][int(rng.integers(0, 2))]
        sq = [(int(rng.integers(8, 33)), int(rng.integers(0, max(1, w - 40))), int(rng.integers(0, max(1, h - 40))),
               int(rng.integers(-6, 7)), int(rng.integers(-4, 5)))]
        ov = dict(w=w, h=h, squares=sq, pan=(int(rng.integers(-8, 9)), int(rng.integers(-4, 5))), seed=int(rng.integers(0, 1000)))
        ex = ";".join(extra)
        try:
            if mode == "chain":
                tracedata.chain_sequence(*tracedata.live_chain(clip, frames, preset, ex, threads=threads, **ov))
            else:
                a, b, _, calls, _ = tracedata.chain_inject_roundtrip(clip, frames, preset, ex, **ov)
                assert calls == 0 and np.array_equal(a, b), "bitstream differs"
            ok += 1
        except Exception as e:  # noqa: BLE001
            fails += 1
            print("FAIL", clip, frames, preset, ex, threads, ov, repr(e)[:300], flush=True)
    print("configurations ok", ok, "failed", fails, flush=True)


if __name__ == "__main__":
    main()
