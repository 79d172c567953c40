#!/usr/bin/env python
"""bench.py -- pictures ENCODED per second: whole streams through the device decision pass, bitstream checked inside the run.

A "step" = S independent streams x F pictures of the workload's synthetic clip (default: 1080p 8-bit, Baseline profile, preset fast,
default hierarchical-B GOP, CQP 32), each stream encoded from its original pictures to the MPEG-5 EVC bitstream:

  device : xb200_analyze_picture per picture -- ONE persistent kernel runs the CTU loop, the quad-tree mode decision with every inter /
           intra CU analysis, then the loop filter and the border expansion; pictures of all streams are enqueued at once and ordered by
           events on their reference pictures (the picture DAG), no host round trip per CU or CTU;
  host   : the reference's own control plane (picture plan: slice types, QPs, lambdas, reference lists) and its own entropy coder /
           bitstream writer, reached through the compiled hook harness (oracle/_ref/libref_harness.so: ctx->fn_mode_analyze_frame hands
           each decided picture over, ctx->fn_mode_analyze_lcu copies the CTU records).  north_star keeps both on the host.  No decision
           is made there: the harness's counters of reference inter / intra analyses must stay 0.

Parity mode: `--threads T` = the reference's `threads` parameter.  Each picture is decided as T coder-state chains (CTU rows y, y + T, ..)
exactly like the reference's worker threads, so the bitstream equals `xeveb_app -m T` byte for byte; T = 1 is the single-thread
bitstream.  The md5 of every stream's bitstream is compared with the unmodified reference's inside the run.

  value : pictures/s over all streams with the original pictures already resident in HBM (decision pass + loop filter only)
  e2e   : the same streams from HOST buffers to the bitstream: H2D of every original picture, D2H of every picture's records, entropy
          coding by the reference's host code (one host thread per stream), md5 check
  --impl reference : the unmodified reference (oracle/_ref) on all usable host cores: floor(cores / T) concurrent instances with T
          threads each, same clip, same frames, "frames / wall time inside xeve_encode" per instance like app/xeve_app.c:1397-1402
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {   # name -> (clip of xeve_b200/clips.py, preset)
    "1080p": ("1080p", "fast"),            # BASELINE.json configs[1]
    "2160p10": ("2160p10", "fast"),        # north_star target: 3840x2160 10-bit preset fast
    "2160p10-medium": ("2160p10", "medium"),  # BASELINE.json configs[2]
    "2160p": ("2160p", "fast"),            # configs[4]: 8-bit 2160p streams
    "cif": ("cif", "fast"),                # configs[0]
}
# compulsory HBM traffic of one inter picture (ME + MC/diff/recon + TQ/ITDQ), bytes per luma sample
ALG_BYTES_PER_SAMPLE = (16.5e6 + 31.1e6 + 24.9e6) / (1920 * 1080)


def usable_cores():
    """Host threads the reference can really run on: the affinity mask capped by the cgroup CPU quota (the GPU boxes expose 128 logical
    CPUs under a 16-CPU quota; oversubscribing the quota makes the reference slower, not faster)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = min(n, max(1, int(float(q) / float(p) + 0.5)))
    except (OSError, ValueError):
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, (q + p // 2) // p))
        except (OSError, ValueError):
            pass
    return n


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.check_output(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                              timeout=5).decode().strip()
                self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


# ---- workload ---------------------------------------------------------------------------------------------------------------------------
def make_clip(args):
    from xeve_b200.clips import Clip
    name, preset = WORKLOADS[args.workload]
    c = Clip(name)
    frames = [c.frame(n) for n in range(args.frames)]
    yuv = np.concatenate([np.concatenate([p.reshape(-1).view(np.uint8) for p in f]) for f in frames])
    return c, preset, frames, yuv


def workload_string(args, c, preset):
    return (f"{c.w}x{c.h} {c.depth}-bit 4:2:0, Baseline profile, preset {preset}, CQP 32, default hierarchical-B GOP16, {args.streams} stream(s) x "
            f"{args.frames} pictures of the synthetic '{WORKLOADS[args.workload][0]}' clip per step, parity mode threads={args.threads} "
            f"(bitstream == reference -m {args.threads})")


# ---- reference arm: the unmodified reference on the host cores ----------------------------------------------------------------------------
def _ref_worker(job):
    path, nframes, w, h, depth, preset, threads = job
    sys.path.insert(0, ROOT)
    from oracle import refharness as rh
    yuv = np.fromfile(path, np.uint8)
    tr = rh.encode_clip(yuv, nframes, w, h, in_depth=depth, preset=preset, threads=threads)
    return tr.enc_seconds, hashlib.md5(tr.bitstream.tobytes()).hexdigest(), len(tr.bitstream)


def reference_step(path, args, c, preset, n_inst):
    """n_inst concurrent instances of the unmodified reference, `threads` threads each, one stream each -> (wall s, [(sec, md5, bytes)])"""
    import multiprocessing as mp
    jobs = [(path, args.frames, c.w, c.h, c.depth, preset, args.threads)] * n_inst
    t0 = time.perf_counter()
    if n_inst == 1:
        res = [_ref_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(n_inst) as pool:
            res = pool.map(_ref_worker, jobs)
    return time.perf_counter() - t0, res


def run_reference(args, steps, warmup, clip=None, quiet=False):
    from oracle import refharness as rh
    if not rh.available():
        return {"impl": "reference", "unavailable": "oracle/_ref is not built"}
    c, preset, frames, yuv = clip or make_clip(args)
    cores = usable_cores()
    n_inst = max(1, min(cores // max(args.threads, 1), 16))
    path = f"/dev/shm/xb200_bench_{os.getpid()}.yuv"
    yuv.tofile(path)
    try:
        times, md5s = [], set()
        for it in range(warmup + steps):
            wall, res = reference_step(path, args, c, preset, n_inst)
            # each instance's fps is frames / time inside xeve_encode; concurrent instances: the step's throughput is over the slowest
            enc = max(r[0] for r in res)
            md5s |= {r[1] for r in res}
            if it >= warmup:
                times.append(enc)
    finally:
        os.unlink(path)
    sec = float(np.mean(times))
    value = n_inst * args.frames / sec
    sample = (f"{n_inst} concurrent instance(s) x {args.threads} threads, {args.frames} pictures each per step (bounded sample of the "
              f"{args.streams}-stream workload: the reference's streams are independent), {steps} step(s) after {warmup} warm-up; "
              f"time = inside xeve_encode, slowest instance")
    cb = {"value": round(value, 3), "unit": "pictures/s", "cores": n_inst * args.threads, "kind": "reference", "sample": sample,
          "usable_cores": cores, "md5": sorted(md5s)[0], "md5_unique": len(md5s) == 1}
    return {"impl": "reference", "metric": "encoded pictures/s", "value": round(value, 3), "unit": "pictures/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": round(1e3 * sec, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "s16", "data": "synthetic", "config": {"workload": workload_string(args, c, preset)},
            "cpu_baseline": cb, "e2e": {"value": round(value, 3), "unit": "pictures/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


# ---- device arm -----------------------------------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, dist, clip, ref_md5):
    import torch
    from oracle import refharness as rh          # host side only: the reference's control plane (plan) and entropy coder
    from xeve_b200 import api
    from xeve_b200.encoder import ClipEncoder
    dev = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(dev)
    c, preset, frames, yuv = clip
    S, F, T = args.streams, args.frames, args.threads
    cst, plan = rh.plan_clip(F, c.w, c.h, in_depth=c.depth, preset=preset, threads=T)
    seq = np.zeros(1, api.SEQ)
    for k in api.SEQ.names:
        seq[k] = cst[k]
    if dist is not None:                          # the "trivial NCCL broadcast of headers": every rank configures its context identically
        t = torch.from_numpy(seq.view(np.uint8).copy()).cuda()
        dist.broadcast(t, src=0)
        seq = t.cpu().numpy().view(api.SEQ).copy()
    hp = api.Hotpath(seq, device=dev)
    capacity = hp.chain_capacity()
    encs = [ClipEncoder(seq, plan, hp=hp, threads=T) for _ in range(S)]
    pocs = [int(p["pp"]["poc"]) for p in plan]

    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    host_frames = {n: tuple(pin(p) for p in frames[n]) for n in range(F)}     # pinned host copies of the original pictures
    h2d = S * sum(p.nbytes for f in host_frames.values() for p in f)
    d2h = S * F * hp.n_lcu * (256 * api.SCU_REC.itemsize + 6144 * 2)

    def upload_all():
        for e in encs:
            e.upload(host_frames, c.depth)

    def enqueue_all():
        for e in encs:
            e.reset()
        for k in range(F):                         # coding order, the streams interleaved: every stream advances with the others
            for e in encs:
                e.enqueue(k)

    chain_ms, n_cu = [], [0, 0]

    def step_device():
        """originals resident in HBM -> every picture decided, filtered, border-expanded (records stay on the device)"""
        enqueue_all()
        for e in encs:
            for poc in pocs:
                st = e.wait(poc)
                chain_ms.append(float(st["chain_ms"]))
                n_cu[0] += int(st["n_inter"]); n_cu[1] += int(st["n_intra"])

    bitstreams = [None] * S

    def step_e2e():
        """host buffers -> bitstream: H2D of the originals, device decision pass, D2H of the records, host entropy coding"""
        upload_all()
        th = threading.Thread(target=enqueue_all, daemon=True)
        th.start()
        errs = []

        def entropy(i):
            try:
                tr, n = rh.encode_clip_lazy(yuv, F, c.w, c.h, encs[i].fetch, in_depth=c.depth, preset=preset, label_threads=T)
                assert n == F * hp.n_lcu
                bitstreams[i] = tr.bitstream
            except Exception as e:  # noqa: BLE001
                errs.append(e)
        ts = [threading.Thread(target=entropy, args=(i,), daemon=True) for i in range(S)]
        for t_ in ts:
            t_.start()
        for t_ in ts:
            t_.join()
        th.join()
        if errs:
            raise errs[0]

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    upload_all()
    for _ in range(args.warmup):
        step_device()
    chain_ms.clear(); n_cu[0] = n_cu[1] = 0
    sampler = ClockSampler(dev)
    sampler.start()
    launches0 = hp.launches
    hp.chain_span_ms(reset=True)
    barrier()
    t0 = time.perf_counter()
    spans = []
    for _ in range(args.steps):
        step_device()
        spans.append(hp.chain_span_ms(reset=True))
    barrier()
    sec = time.perf_counter() - t0
    launches = hp.launches - launches0
    # e2e: one warm-up, then the timed steps
    step_e2e()
    barrier()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    sec_e2e = time.perf_counter() - t1
    sampler.stop_flag = True
    sampler.join()
    md5s = sorted({hashlib.md5(b.tobytes()).hexdigest() for b in bitstreams})
    ok = ref_md5 is not None and md5s == [ref_md5]
    if dist is not None:
        tt = torch.tensor([sec, sec_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec, sec_e2e = float(tt[0]), float(tt[1])
        okt = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        ok = bool(int(okt[0]))
    pics = world * S * F * args.steps
    value, e2e = pics / sec, pics / sec_e2e
    # roofline of the dominant kernel (k_chain): algorithmic bytes of one picture / mean launch duration measured live with CUDA events
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg = ALG_BYTES_PER_SAMPLE * c.w * c.h
    mean_ms = float(np.mean(chain_ms)) if chain_ms else 0.0
    achieved = alg / (mean_ms * 1e-3) / 1e9 if mean_ms else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("k_chain_" + args.workload)
    except (OSError, ValueError):
        pass
    out = {
        "metric": "encoded pictures/s", "value": round(value, 3), "unit": "pictures/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(1e3 * sec / args.steps, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "s16", "data": "synthetic",
        "config": {"workload": workload_string(args, c, preset), "streams_per_gpu": S, "pictures_per_stream": F, "threads": T,
                   "l2": f"inputs larger than L2 ({h2d / 2**20:.0f} MiB of original pictures per step)",
                   "host_side": "reference control plane + entropy coder through oracle/_ref/libref_harness.so (decisions: 0 on the host)"},
        "e2e": {"value": round(e2e, 3), "unit": "pictures/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": round(1e3 * sec_e2e / args.steps, 2), "bitstream_md5": md5s, "reference_md5": ref_md5, "bitstream_matches_reference": ok,
                "bitstream_bytes_per_stream": int(len(bitstreams[0]))},
        "gpu_launches": int(launches),
        "device_span_ms_per_step": round(float(np.mean(spans)), 2),
        "chain": {"capacity_chains": capacity, "chains_per_picture": min(T, (c.h + 63) // 64), "kernel_ms_per_picture": round(mean_ms, 2),
                  "cu_analyses_per_step": int((n_cu[0] + n_cu[1]) / max(args.steps, 1)),
                  "us_per_cu_decision_per_chain": round(1e3 * float(np.sum(chain_ms)) * min(T, (c.h + 63) // 64) / max(n_cu[0] + n_cu[1], 1), 1)},
        "roofline": {"bound": "hbm", "achieved": round(achieved, 4), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 8),
                     "traffic": traffic, "kernel": "k_chain (decision pass of one picture: ME + MC + TQ + RDO + tree, latency bound)",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                     "algorithmic_bytes_per_launch": int(alg)},
        "clocks": sampler.summary(),
    }
    if not ok:
        out["invalid"] = "bitstream md5 differs from the reference's"
    for e in encs:
        e.close()
    hp.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--streams", type=int, default=4, help="independent streams per GPU per step")
    ap.add_argument("--frames", type=int, default=17, help="pictures per stream (17 = the intra picture + one GOP of 16)")
    ap.add_argument("--threads", type=int, default=8, help="parity mode: the reference's `threads` (coder-state chains per picture)")
    args = ap.parse_args()
    # exactly ONE line on stdout: everything libraries print (NCCL's version banner, torchrun notices) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    if args.impl == "reference":
        if rank == 0:
            emit(run_reference(args, args.steps, min(args.warmup, 1)))
        return
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
    clip = make_clip(args)
    ref = None
    if rank == 0:      # the reference's bitstream of this clip (and, at N = 1, the CPU baseline): one step, no warm-up
        ref = run_reference(args, 1, 0, clip=clip)
    md5 = [ref["cpu_baseline"]["md5"] if ref and "cpu_baseline" in ref else None]
    if dist is not None:
        dist.broadcast_object_list(md5, src=0)
    out = run_b200(args, rank, world, dist, clip, md5[0])
    if rank == 0:
        if ref and "cpu_baseline" in ref:
            out["cpu_baseline"] = ref["cpu_baseline"]
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
