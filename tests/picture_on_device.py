"""The decision pass of whole pictures on the device (xb200_analyze_picture; run by tests/test_gpu_picture.py in a subprocess, or by hand).

For every picture ONE enqueue runs the persistent chain kernel (mode_analyze_lcu / mode_coding_tree / mode_coding_unit with every inter
and intra CU analysis, one CTA per coder-state chain), the loop filter and the border expansion; the host only uploads the original
picture and reads the records back.  Checked against what the unmodified reference left behind for the same pictures: coder states
before / after every CTU, frame maps, leaf CUs, the picture before and after deblocking; then the records are injected into the
unmodified reference (its own entropy coder writes the bitstream): the bitstream must be byte-identical.

  python tests/picture_on_device.py                 # fixture (3 pictures) + live QCIF default GOP, threads 1 and 2
  --oracle      also compare EVERY CU analysis (inputs and results, in call order) with the oracle chain: first difference is reported
  --more        further live configurations (10-bit medium, low QP, plain quantiser, CIF)
  --fixture-only
  --config clip:preset:frames:threads[:extra][,...]   live runs of full-size clips (e.g. 1080p:fast:3:8)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tracedata  # noqa: E402
from tracedata import rh  # noqa: E402


def run(seq, pics, compare_oracle=False, threads=None, overlap=False, check=True):
    from xeve_b200 import api
    hp = api.Hotpath(seq)
    if os.environ.get("XB200_WATCHDOG"):       # debug builds: report where the kernel is if it does not finish in time
        import threading

        def dog():
            time.sleep(float(os.environ["XB200_WATCHDOG"]))
            print("WATCHDOG: kernel progress words", hp.chain_debug(), flush=True)
            os._exit(3)
        threading.Thread(target=dog, daemon=True).start()
    enc = tracedata.DevicePictureEncoder(seq, hp, check=check, compare_oracle=compare_oracle, threads=threads)
    t0 = time.time()
    try:
        if overlap:                      # enqueue everything, then collect: the library orders the pictures by their references
            for pc in pics:
                enc.enqueue(pc)
            out = enc.collect()
        else:
            out = [enc.encode(pc) for pc in pics]
    except Exception:
        print("kernel progress words (debug builds):", hp.chain_debug(), flush=True)
        raise
    sec = time.time() - t0
    launches = hp.launches
    prof = hp.chain_prof()
    if prof is not None:
        names = {0: "load", 1: "skip", 3: "ME uni", 4: "mvp check", 6: "bi loop", 7: "pre-final", 8: "final copy-out", 9: "intra (all)",
                 10: "tree bookkeeping", 12: "RDO: predict", 13: "RDO: transforms", 14: "RDO: cbf coder", 18: "[inter 8x8 total]",
                 19: "[inter 16x16 total]", 20: "[inter 32x32 total]", 21: "[inter 64x64 total]", 22: "ME: window staging",
                 23: "ME: first diamond", 24: "ME: refinement runs", 25: "ME: sub-pel"}
        cyc, cnt = prof
        tot = sum(int(cyc[k]) for k in names if k < 18 or k > 21) or 1
        for k, nm in names.items():
            if cnt[k]:
                print(f"    prof {nm:22s} {100 * int(cyc[k]) / tot:5.1f}%  {int(cyc[k]) / 1.9e3 / max(int(cnt[k]), 1):9.1f} us/visit  x{int(cnt[k])}")
    hp.close()
    return out, launches, sec


def summary(out):
    n_cu = sum(int(r["stat"]["n_inter"]) + int(r["stat"]["n_intra"]) for r in out)
    ms = sum(float(r["stat"]["chain_ms"]) for r in out)
    return n_cu, ms


QCIF = {k: v for k, v in tracedata.QCIF.items() if k != "n"}
DEFAULT = [("cif", "fast", 6, "", 1, QCIF), ("cif", "fast", 6, "", 2, QCIF)]
MORE = [("2160p10", "medium", 5, "", 1, dict(w=256, h=192, squares=[(48, 60, 40, 5, 2)], pan=(6, 2))),   # 10-bit input, preset medium
        ("cif", "fast", 5, "rdoq=0;qp=27", 1, QCIF),                                                      # plain quantiser, lower QP
        ("cif", "fast", 5, "qp=22", 3, QCIF),                                                             # low QP, three row chains
        ("cif", "fast", 20, "", 1, QCIF)]                                                                 # the whole default GOP + 3


def main():
    oracle = "--oracle" in sys.argv
    if "--config" not in sys.argv:
        seq, pics = tracedata.chain_golden()
        out, launches, sec = run(seq, pics, compare_oracle=oracle)
        n_cu, ms = summary(out)
        print(f"fixture: {len(out)} pictures, {n_cu} CU analyses in {launches} kernel launches, chain kernels {ms:.1f} ms "
              f"({1e3 * ms / max(n_cu, 1):.0f} us per CU decision): states, maps, leaf CUs and pictures equal the reference's", flush=True)
    if "--fixture-only" in sys.argv or not rh.available():
        print("PICTURE_ON_DEVICE_OK")
        return
    configs = MORE if "--more" in sys.argv else DEFAULT
    if "--config" in sys.argv:           # --config clip:preset:frames:threads[:extra]  (full-size clips of xeve_b200/clips.py)
        configs = []
        for a in sys.argv[sys.argv.index("--config") + 1].split(","):
            f = a.split(":")
            configs.append((f[0], f[1], int(f[2]), f[4] if len(f) > 4 else "", int(f[3]), {}))
    for name, preset, frames, extra, threads, override in configs:
        c, yuv = tracedata.clip_yuv(name, frames, **override)
        tr = rh.encode_clip(yuv, frames, c.w, c.h, in_depth=c.depth, preset=preset, extra=extra, threads=threads,
                            trace_mask=rh.TRACE_LCU | rh.TRACE_DF, pic_lo=0, pic_hi=1 << 20)
        seq, pics = tracedata.chain_inputs_from_trace(tr)
        out, launches, sec = run(seq, pics, compare_oracle=oracle and threads == 1, overlap="XB200_LIB" not in os.environ)
        n_cu, ms = summary(out)
        dec = [dict(poc=r["poc"], scu=r["scu"], coef=r["coef"], rec=r["rec"]) for r in out]
        tr2, n_ctu, ncu, nintra = rh.encode_clip_injected(yuv, frames, c.w, c.h, dec, in_depth=c.depth, preset=preset, extra=extra, threads=threads)
        assert ncu == 0 and nintra == 0 and n_ctu == sum(len(r["scu"]) for r in out)
        assert len(tr.bitstream) > 1000 and np.array_equal(tr.bitstream, tr2.bitstream), "bitstream differs"
        print(f"live {name} {c.w}x{c.h} {preset} {extra or 'default'} threads={threads}: {len(out)} pictures, {n_cu} CU analyses, {launches} kernel "
              f"launches, chain kernels {ms:.1f} ms ({1e3 * ms / max(n_cu, 1):.0f} us per CU decision per chain), wall {sec:.2f} s: "
              f"injected into the unmodified reference -> byte-identical bitstream ({len(tr.bitstream)} bytes)", flush=True)
    print("PICTURE_ON_DEVICE_OK")


if __name__ == "__main__":
    main()
