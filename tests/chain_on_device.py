"""The CUDA library inside the real decision chain (run by tests/test_zz_gpu_chain.py in a subprocess, or by hand).

Every xeve_pinter_analyze_cu / pintra_analyze_cu of every CU the tree visits, the winner's prediction and the loop filter + border
expansion of every picture are computed by libxeve_b200.so on the device (xb200_analyze_cu, xb200_analyze_intra, xb200_mc,
xb200_deblock; reference pictures stay device-resident); the tree bookkeeping around them is the oracle's (test infrastructure as the
DRIVER, not as the thing under test).  Checked: coder states per CTU, frame maps, leaf CUs and the deblocked pictures equal the
reference's for the committed three-picture fixture and, where oracle/_ref is present, for a live 6-picture encode whose decisions are
then injected into the unmodified reference: the bitstream must be byte-identical.

  python tests/chain_on_device.py             # needs a GPU
  python tests/chain_on_device.py --stand-in  # same plumbing with a CPU stand-in for the device context (no GPU; plumbing check)
  --all-inputs   also the inputs of the analyses come from the device: xb200_mvp (MV predictor candidates, temporal direct MVs) and
                 xb200_intra_nbr (availability, reference samples, MPM list; the picture under reconstruction is uploaded per call)
  --more         further configurations (10-bit medium, P slices, plain quantiser) instead of the default pair
  --dag          (under torchrun, one rank per GPU; with --stand-in: gloo on CPU) one stream over several ranks: picture-DAG waves,
                 reference pictures broadcast after each wave (NCCL), every rank's pictures checked against the reference's
"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tracedata  # noqa: E402
from tracedata import rh, xo  # noqa: E402


class StandIn:
    """CPU stand-in with the Hotpath methods the chain uses (oracle functions on host copies of the pictures)."""

    def __init__(self, seq):
        self.seq, self.pics, self.planes = seq, {}, (xo.PLANES * 4096)()

    def pic_create(self, padded):
        h = len(self.pics)
        self.pics[h] = dict(padded=padded)
        return h

    def _bind(self, h, bufs, pads):
        self.pics[h]["bufs"] = bufs
        self.planes[h].y, self.planes[h].u, self.planes[h].v = [b.ctypes.data + 2 * (pd * b.shape[1] + pd) for b, pd in zip(bufs, pads)]
        self.planes[h].s_l, self.planes[h].s_c = bufs[0].shape[1], bufs[1].shape[1]

    def _bind_padded(self, h, planes):      # what the library does for a padded picture: borders replicated on upload / after deblocking
        bufs = []
        for a, pad in zip(planes, (144, 72, 72)):
            hh, ww = a.shape
            buf = np.zeros((hh + 2 * pad, ww + 2 * pad), np.int16)
            buf[pad:pad + hh, pad:pad + ww] = a
            xo.lib().xo_pad_plane(buf.ctypes.data_as(C.c_void_p), buf.shape[1], ww, hh, pad)
            bufs.append(buf)
        self._bind(h, bufs, (144, 72, 72))

    def pic_upload_s16(self, h, y, u, v):
        self.pics[h]["act"] = [np.array(a, np.int16) for a in (y, u, v)]
        if self.pics[h]["padded"]:
            self._bind_padded(h, self.pics[h]["act"])
        else:
            self._bind(h, self.pics[h]["act"], (0, 0, 0))

    def deblock(self, h, cus, pp, map_scu, map_refi, map_mv, expand=True):
        post = xo.deblock(self.pics[h]["act"], cus, pp, map_scu, map_refi, map_mv, bit_depth=int(np.asarray(self.seq).reshape(-1)[0]["bit_depth"]))
        self.pics[h]["act"] = post
        self._bind_padded(h, post)

    def pic_download(self, h, with_padding):
        return self.pics[h]["act"]

    def analyze_cu(self, items, rates, states, elems):
        return xo.analyze_cu_batch(self.seq, self.planes, rates, items, states, elems)

    def mc(self, items, off, total):
        return xo.mc_batch(self.seq, self.planes, items, off, total)

    def analyze_intra(self, items, rates, states, side, elems):
        return xo.analyze_intra_batch(self.seq, self.planes, rates, items, states, side, elems)

    def mvp(self, items, pic, map_scu, map_mv, col0, col1):
        return xo.mvp_batch(items, pic, map_scu, map_mv, col0, col1)

    def intra_nbr(self, h, items, map_scu, map_ipm, w_scu, h_scu, cip, side_elems):
        return xo.intra_nbr(self.pics[h]["act"], items, map_scu, map_ipm, w_scu, h_scu, cip, side_elems,
                            bit_depth=int(np.asarray(self.seq).reshape(-1)[0]["bit_depth"]))

    def close(self):
        pass


def run(seq, pics, stand_in, all_inputs=False):
    if stand_in:
        hp = StandIn(seq)
    else:
        from xeve_b200 import api
        hp = api.Hotpath(seq)
    t0 = time.time()
    extra = {}
    if all_inputs:
        scratch = hp.pic_create(padded=True)    # the configuration xb200_intra_nbr is tested with (test_gpu_parity.py)

        def upload(rec):
            hp.pic_upload_s16(scratch, *rec)
            return scratch
        extra = dict(mvp=hp.mvp, intra_nbr=hp.intra_nbr, upload=upload)
    kernel_ms = [0.0]

    def timed(fn):                      # device time of the operator's kernels (CUDA events on the library's stream), summed over the calls
        def call(*args, **kw):
            r = fn(*args, **kw)
            kernel_ms[0] += float(getattr(hp, "last_kernel_ms", 0.0))
            return r
        return call
    with tracedata.chain_with(timed(hp.analyze_cu), timed(hp.mc), timed(hp.analyze_intra), **extra) as cw:
        out = tracedata.chain_sequence(seq, pics, check=True, hp=hp)
    n_calls = (cw.n_cu, cw.n_intra)
    if not stand_in:
        print(f"  device kernel time inside the {cw.n_cu * 2 + cw.n_intra} analysis / prediction calls: {kernel_ms[0] / 1e3:.2f} s")
    assert cw.n_cu == sum(len(r["cu_log"]) for r in out) and cw.n_intra == sum(len(r["intra_log"]) for r in out)
    assert not all_inputs or (cw.n_nbr == cw.n_intra and cw.n_mvp >= cw.n_cu)
    launches = 0 if stand_in else hp.launches
    hp.close()
    return out, n_calls, launches, time.time() - t0


QCIF = {k: v for k, v in tracedata.QCIF.items() if k != "n"}
DEFAULT = [("cif", "fast", 6, "", QCIF)]
MORE = [("2160p10", "medium", 5, "", dict(w=256, h=192, squares=[(48, 60, 40, 5, 2)], pan=(6, 2))),     # 10-bit input, preset medium
        ("cif", "fast", 6, "bframes=0;inter_slice_type=1", QCIF),                                       # low delay, P slices
        ("cif", "fast", 5, "rdoq=0;qp=27", QCIF)]                                                       # plain quantiser, lower QP


def refs_of(pc):
    pp = pc["pp"]
    return sorted({int(pp["ref_poc"][l][k]) for l in range(2) for k in range(4) if int(pp["ref_pic"][l][k]) >= 0})


def run_dag(rank, world, dist, stand_in, frames=17, device="cpu"):
    """One stream over several ranks: the pictures of each wave of the picture DAG round-robin over the ranks
    (xeve_b200.dist.picture_plan), every reference picture broadcast once after its wave (broadcast_picture); each rank decides its
    pictures with its own device context (or the CPU stand-in) inside the decision chain and checks them against the reference's.
    Returns (pictures decided here, pictures received)."""
    from xeve_b200 import dist as xd
    seq, pics = tracedata.live_chain(frames=frames)                  # every rank traces the same deterministic reference encode
    by_poc = {int(pc["pp"]["poc"]): pc for pc in pics}
    waves, owner, exchanged = xd.picture_plan([(poc, refs_of(pc)) for poc, pc in by_poc.items()], world)
    if stand_in:
        hp = StandIn(seq)
    else:
        from xeve_b200 import api
        hp = api.Hotpath(seq, device=int(str(device).split(":")[1]) if ":" in str(device) else 0)
    enc = tracedata.ChainEncoder(seq, check=True, hp=hp)
    mine, received = [], []
    with tracedata.chain_with(hp.analyze_cu, hp.mc, hp.analyze_intra):
        for wave in waves:
            for poc in wave:
                if owner[poc] == rank:
                    enc.encode(by_poc[poc])
                    mine.append(poc)
            for poc in wave:                                          # the exchange step of the wave
                if poc not in exchanged:
                    continue
                shape = [a.shape for a in by_poc[poc]["org"]]
                f = ((shape[0][1] + 3) // 4) * ((shape[0][0] + 3) // 4)
                if owner[poc] == rank:
                    planes, mv = enc.done[poc]["post"], enc.done[poc]["map_mv"]
                else:
                    planes, mv = [np.zeros(sh, np.int16) for sh in shape], np.zeros((f, 2, 2), np.int16)
                planes, mv = xd.broadcast_picture(planes, np.asarray(mv).reshape(f, 2, 2), owner[poc], dist, device=device)
                if owner[poc] != rank:
                    enc.adopt(poc, planes, mv)
                    received.append(poc)
    hp.close()
    return mine, received, waves, sorted(exchanged)


def dag_main():
    """torchrun entry: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/chain_on_device.py --dag"""
    import torch
    import torch.distributed as dist
    stand_in = "--stand-in" in sys.argv
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    if stand_in:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        device = "cpu"
    else:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
        device = f"cuda:{local}"
    t0 = time.time()
    mine, received, waves, exchanged = run_dag(rank, world, dist, stand_in, device=device)
    dist.barrier()
    print(f"rank {rank}/{world}: decided {sorted(mine)}, received {sorted(received)}, {time.time() - t0:.1f} s", flush=True)
    if rank == 0:
        print(f"waves {waves}; reference pictures that travel: {exchanged}\nCHAIN_DAG_OK", flush=True)
    dist.destroy_process_group()


def main():
    if "--dag" in sys.argv:
        return dag_main()
    stand_in, all_inputs = "--stand-in" in sys.argv, "--all-inputs" in sys.argv
    if "--more" not in sys.argv:
        seq, pics = tracedata.chain_golden()
        out, calls, launches, sec = run(seq, pics, stand_in, all_inputs)
        print(f"fixture: {len(out)} pictures, {calls[0]} inter + {calls[1]} intra CU analyses, {launches} kernel launches, {sec:.1f} s: "
              "states, maps, leaf CUs and deblocked pictures equal the reference's")
    for name, preset, frames, extra, override in (MORE if "--more" in sys.argv else DEFAULT) if rh.available() and "--fixture-only" not in sys.argv else []:
        c, yuv = tracedata.clip_yuv(name, frames, **override)
        tr = rh.encode_clip(yuv, frames, c.w, c.h, in_depth=c.depth, preset=preset, extra=extra, trace_mask=rh.TRACE_LCU | rh.TRACE_DF,
                            pic_lo=0, pic_hi=1 << 20)
        seq, pics = tracedata.chain_inputs_from_trace(tr)
        out, calls, launches, sec = run(seq, pics, stand_in, all_inputs)
        dec = [dict(poc=r["poc"], scu=r["scu"], coef=r["coef"], rec=r["rec"]) for r in out]
        tr2, n_ctu, ncu, nintra = rh.encode_clip_injected(yuv, frames, c.w, c.h, dec, in_depth=c.depth, preset=preset, extra=extra)
        assert ncu == 0 and nintra == 0 and n_ctu == sum(len(r["ctu"]) for r in out)
        assert len(tr.bitstream) > 1000 and np.array_equal(tr.bitstream, tr2.bitstream)
        print(f"live {name} {preset} {extra or 'default'}: {len(out)} pictures, {calls[0]} inter + {calls[1]} intra CU analyses, {launches} kernel "
              f"launches, {sec:.1f} s: injected into the unmodified reference -> byte-identical bitstream ({len(tr.bitstream)} bytes)")
    print("CHAIN_ON_DEVICE_OK")


if __name__ == "__main__":
    main()
