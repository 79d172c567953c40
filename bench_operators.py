#!/usr/bin/env python
"""bench_operators.py -- SECONDARY benchmark (round 1's bench.py): throughput of the batched work-list operators on one B picture.
Nothing is encoded here; the headline benchmark is bench.py (pictures ENCODED per second through xb200_analyze_picture).

Hot-path throughput of the B200 library on the BASELINE.json workload.

A "step" = one pass of the inter-search + transform hot path over ONE 1080p B picture
(Baseline profile, preset fast): the frame-wide work lists of xeve_b200/worklist.py
(2 uni-directional searches, the bi-prediction search, and the L0/L1/BI residue candidates of
every CU of the 64/32/16/8 quad-tree) run as frame-wide grids.  Metric: frames per second.

  value : whole-job fps with all inputs resident in HBM (C ABI called with XB200_MEM_DEVICE)
  e2e   : same calls with HOST (pinned) buffers -- picture upload, work lists in, every result out
  --impl reference : the reference's own CPU functions (oracle/_ref, all host threads) over a
                     bounded sample of the same work lists
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CLIP, W, H = "1080p", 1920, 1080
POC, REF_POCS = 8, (0, 16)
PAN = (3, 1)
SAMPLE_ROWS = 4  # CTU rows of the bounded CPU sample
PRESET = "fast"


def usable_cores():
    """Host threads the reference arm can really run on: the affinity mask capped by the cgroup CPU quota (the GPU boxes expose
    128 logical CPUs under a 16-CPU quota; oversubscribing the quota makes the reference slower, not faster)."""
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


def select_workload(name):
    """1080p fast (BASELINE.json configs[1], the default) or 2160p 10-bit medium (configs[2])."""
    global CLIP, W, H, PRESET
    if name == "2160p":
        CLIP, W, H, PRESET = "2160p10", 3840, 2160, "medium"


def frames_for_bench():
    from xeve_b200.clips import Clip
    c = Clip(CLIP)
    return c, {n: c.frame(n) for n in (REF_POCS[0], POC, REF_POCS[1])}


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


def algorithmic_bytes(fw):
    """SURVEY.md 8(d) compulsory HBM traffic of each stage for the units it processes (bytes).
    ME: original luma once + each searched padded reference luma once + one result record per item.
    residue: per item 3/2*N^2 samples x (org read + R reference reads + coef write); the per-candidate
    reconstruction is consumed on chip (SSD) and not stored, as in the reference (xeve_pinter.c:2006-2038)."""
    w, h = fw.w, fw.h
    ref_luma = 2 * (w + 288) * (h + 288)
    n_me = 2 * fw.n_cu
    me_uni = 2 * w * h + 2 * ref_luma + n_me * 16
    me_bi = 2 * fw.side_elems + 2 * ref_luma + fw.n_cu * 16
    bi_org = 2 * w * h + ref_luma + 2 * fw.side_elems
    area = (1 << (2 * fw.l2.astype(np.int64))) * 3 // 2 * 2  # bytes of one Y+U+V block
    residue = int((area * (1 + 1 + 1)).sum() * 2 + (area * (1 + 2 + 1)).sum())  # L0, L1 (R=1) + BI (R=2); coef out, rec not stored
    return dict(me_uni=me_uni, bi_org=bi_org, me_bi=me_bi, residue=residue)


def run_b200(args, rank, world, dist):
    import torch
    from xeve_b200 import api
    from xeve_b200.worklist import FrameWork

    dev = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(dev)
    seq = api.make_seq(W, H, PRESET)
    if world > 1:  # the only exchange the path has: rank 0 broadcasts the sequence header
        from xeve_b200 import dist as xd
        seq = xd.broadcast_seq(seq, dist, device="cuda")
    hp = api.Hotpath(seq, device=dev)
    L, ctx = hp.L, hp.h
    clip, fr = frames_for_bench()
    def pin(a):
        a = np.ascontiguousarray(a)
        if a.dtype == np.uint16:
            a = a.view(np.int16)  # torch has no pinned uint16; the bytes are what matters
        return torch.from_numpy(a).pin_memory()
    refs = []
    for poc in REF_POCS:
        hd = hp.pic_create(padded=True)
        hp.pic_upload(hd, *fr[poc], clip.depth)
        refs.append(hd)
    cur = hp.pic_create(padded=False)
    cur_planes = [pin(p) for p in fr[POC]]
    cur_np = [p.numpy() for p in cur_planes]
    hp.pic_upload(cur, *cur_np, clip.depth)

    fw = FrameWork(W, H, POC, REF_POCS, PAN, cur, refs, seed=rank, me_range=int(seq["me_range"][0]))
    cnt = fw.counts()
    # ---- warm-up pass through the host-buffer API; also builds the dependent work lists ----------------
    me_uni = hp.me(fw.me_uni)
    bi_mc, me_bi_in = fw.build_bi(me_uni)
    side = hp.bi_org(bi_mc, fw.bi_cur, fw.side_off, fw.side_elems)
    me_bi = hp.me(me_bi_in, side)
    res_in = fw.build_residue(me_bi)
    res_out, coef, _ = hp.residue(res_in, fw.rates, fw.res_elems, want_rec=False)
    checksum = int(res_out["dist_rec"].sum() % (1 << 31))

    # ---- device-resident buffers -----------------------------------------------------------------------------
    dv = lambda a: torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).cuda()
    d_me_uni, d_bi_mc, d_bi_cur, d_side_off = dv(fw.me_uni), dv(bi_mc), dv(fw.bi_cur), dv(fw.side_off)
    d_me_bi, d_res, d_rates = dv(me_bi_in), dv(res_in), dv(fw.rates)
    d_side = torch.zeros(fw.side_elems, dtype=torch.int16, device="cuda")
    d_coef = torch.zeros(fw.res_elems, dtype=torch.int16, device="cuda")
    P = lambda t: C.c_void_p(t.data_ptr())
    stage_ms = {k: 0.0 for k in ("me_uni", "bi_org", "me_bi", "residue")}

    def step_device(accumulate):
        for name, call in (
            ("me_uni", lambda: L.xb200_me(ctx, P(d_me_uni), cnt["me_uni"], None, 0, api.MEM_DEVICE)),
            ("bi_org", lambda: L.xb200_bi_org(ctx, P(d_bi_mc), cnt["bi_org"], P(d_bi_cur), P(d_side_off), P(d_side), fw.side_elems, api.MEM_DEVICE)),
            ("me_bi", lambda: L.xb200_me(ctx, P(d_me_bi), cnt["me_bi"], P(d_side), fw.side_elems, api.MEM_DEVICE)),
            ("residue", lambda: L.xb200_residue(ctx, P(d_res), cnt["residue"], P(d_rates), 1, P(d_coef), None, fw.res_elems, api.MEM_DEVICE)),
        ):
            r = call()
            if r != 0:
                raise RuntimeError(f"{name} failed: {r}")
            if accumulate:
                stage_ms[name] += hp.last_kernel_ms

    sampler = ClockSampler(dev)
    sampler.start()  # samples clocks / throttle reasons from the warm-up through the timed regions
    for _ in range(args.warmup):
        step_device(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = hp.launches
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    t_dev = 0.0
    for _ in range(args.steps):
        flush.fill_(1)  # evict the 126 MB L2 between timed steps (outside the timed region)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step_device(True)
        torch.cuda.synchronize()
        t_dev += time.perf_counter() - t0
    launches = hp.launches - launches0
    # results of the device-resident run equal the host-buffer run
    got = np.frombuffer(d_res.cpu().numpy().tobytes(), api.RESIDUE_ITEM)
    assert int(got["dist_rec"].sum() % (1 << 31)) == checksum, "device-resident run disagrees with host-buffer run"

    # ---- e2e: the same calls with host (pinned) buffers, picture upload included ----------------------------
    h_me_uni, h_me_bi, h_res = pin(fw.me_uni.view(np.uint8)), pin(me_bi_in.view(np.uint8)), pin(res_in.view(np.uint8))
    h_side = pin(np.zeros(fw.side_elems, np.int16))
    h_coef = pin(np.zeros(fw.res_elems, np.int16))
    HP = lambda t: C.c_void_p(t.data_ptr())
    planes = (C.c_void_p * 3)(*[p.data_ptr() for p in cur_planes])
    bps = 2 if clip.depth > 8 else 1
    strides = (C.c_int32 * 3)(W * bps, W // 2 * bps, W // 2 * bps)

    def step_host():
        rr = [L.xb200_pic_upload(ctx, cur, planes, strides, clip.depth, api.MEM_HOST),
              L.xb200_me(ctx, HP(h_me_uni), cnt["me_uni"], None, 0, api.MEM_HOST),
              L.xb200_bi_org(ctx, bi_mc.ctypes.data_as(C.c_void_p), cnt["bi_org"], fw.bi_cur.ctypes.data_as(C.c_void_p),
                             fw.side_off.ctypes.data_as(C.c_void_p), HP(h_side), fw.side_elems, api.MEM_HOST),
              L.xb200_me(ctx, HP(h_me_bi), cnt["me_bi"], HP(h_side), fw.side_elems, api.MEM_HOST),
              L.xb200_residue(ctx, HP(h_res), cnt["residue"], fw.rates.ctypes.data_as(C.c_void_p), 1, HP(h_coef), None,
                              fw.res_elems, api.MEM_HOST)]
        if any(rr):
            raise RuntimeError(f"e2e step failed: {rr}")

    for _ in range(max(1, args.warmup // 2)):
        step_host()
    if world > 1:
        dist.barrier()
    e2e_steps = max(2, args.steps // 2)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    # ---- the fused per-CU decision over the same picture (xb200_analyze_cu, host-buffer API) ------------------------
    cu_items = fw.build_cu(hp.rdoq_rates, max_search_range=int(seq["me_range"][0]), seed=rank)
    h_cu, h_st = pin(cu_items.view(np.uint8)), pin(fw.cu_states.view(np.uint8))
    h_cu_coef, h_cu_rec = pin(np.zeros(fw.cu_elems, np.int16)), pin(np.zeros(fw.cu_elems, np.int16))
    cu_rates_p = fw.cu_rates.ctypes.data_as(C.c_void_p)

    def step_cu():
        r = L.xb200_analyze_cu(ctx, HP(h_cu), len(cu_items), cu_rates_p, len(fw.cu_rates), HP(h_st), len(fw.cu_states), HP(h_cu_coef),
                               HP(h_cu_rec), fw.cu_elems)
        if r:
            raise RuntimeError(f"analyze_cu failed: {r}")
    step_cu()  # warm-up
    cu_out = np.frombuffer(h_cu.numpy().tobytes(), api.CU_ITEM)
    t_cu_k, t0 = 0.0, time.perf_counter()
    cu_steps = max(2, args.steps // 4)
    for _ in range(cu_steps):
        step_cu()
        t_cu_k += hp.last_kernel_ms
    t_cu = (time.perf_counter() - t0) / cu_steps
    analyze = {"cus_per_frame": int(len(cu_items)), "kernel_ms_per_frame": round(t_cu_k / cu_steps, 3),
               "host_api_ms_per_frame": round(t_cu * 1e3, 3), "frames_per_s_kernel": round(1e3 / (t_cu_k / cu_steps), 2),
               "frames_per_s_host_api": round(1.0 / t_cu, 2),
               "h2d_bytes": int(cu_items.nbytes + fw.cu_states.nbytes + fw.cu_rates.nbytes),
               "d2h_bytes": int(cu_items.nbytes + fw.cu_states.nbytes + 4 * fw.cu_elems),
               "best_mode_hist": np.bincount(cu_out["best_idx"], minlength=5).tolist(),
               "note": "whole xeve_pinter_analyze_cu per CU on the device (skip/direct/L0/L1/BI + cbf RDO + CABAC bit counts)"}
    # ---- intra analysis of every CU of the 32/16/8/4 quad-tree of the same picture (xb200_analyze_intra) --------
    from xeve_b200.clips import to_internal10
    from xeve_b200.worklist import synth_intra
    in_items, in_states, in_rates, in_side, in_elems = synth_intra(W, H, [to_internal10(p, clip.depth) for p in fr[POC]], cur, hp.rdoq_rates,
                                                                     seed=rank)
    h_in, h_inst, h_inside = pin(in_items.view(np.uint8)), pin(in_states.view(np.uint8)), pin(in_side)
    h_in_coef, h_in_rec = pin(np.zeros(in_elems, np.int16)), pin(np.zeros(in_elems, np.int16))
    in_rates_p = in_rates.ctypes.data_as(C.c_void_p)

    def step_intra():
        r = L.xb200_analyze_intra(ctx, HP(h_in), len(in_items), in_rates_p, len(in_rates), HP(h_inst), len(in_states), HP(h_inside),
                                  len(in_side), HP(h_in_coef), HP(h_in_rec), in_elems)
        if r:
            raise RuntimeError(f"analyze_intra failed: {r}")
    step_intra()
    in_out = np.frombuffer(h_in.numpy().tobytes(), api.INTRA_ITEM)
    t_in_k, t0 = 0.0, time.perf_counter()
    in_steps = max(2, args.steps // 4)
    for _ in range(in_steps):
        step_intra()
        t_in_k += hp.last_kernel_ms
    t_in = (time.perf_counter() - t0) / in_steps
    intra = {"cus_per_frame": int(len(in_items)), "cu_sizes": {str(1 << k): int(v) for k, v in enumerate(np.bincount(in_items["log2_cuw"])) if v},
             "kernel_ms_per_frame": round(t_in_k / in_steps, 3), "host_api_ms_per_frame": round(t_in * 1e3, 3),
             "cus_per_s_kernel": round(len(in_items) / (t_in_k / in_steps * 1e-3), 1),
             "h2d_bytes": int(in_items.nbytes + in_states.nbytes + in_rates.nbytes + in_side.nbytes),
             "d2h_bytes": int(in_items.nbytes + in_states.nbytes + 4 * in_elems),
             "mode_hist": np.bincount(in_out["ipm"][:, 0], minlength=5).tolist(),
             "note": "whole pintra_analyze_cu per CU on the device (5 predictors, SATD ranking, luma + chroma RDO with RDOQ and CABAC "
                     "bit counts); frame-parallel mode: reference samples from the original picture"}
    # ---- in-loop deblocking + border expansion of one reconstructed picture (xb200_deblock) ------------
    from xeve_b200.worklist import synth_deblock
    df = synth_deblock(W, H, seed=rank)
    df_pic = hp.pic_create(padded=True)
    df_pre = [np.ascontiguousarray(a) for a in df["pre"]]
    df_args = [np.ascontiguousarray(df["cus"], api.DF_CU), np.ascontiguousarray(df["pp"], api.DF_PIC).reshape(1),
               np.ascontiguousarray(df["map_scu"], np.uint32), np.ascontiguousarray(df["map_refi"], np.int8),
               np.ascontiguousarray(df["map_mv"], np.int16)]
    df_dev = [dv(a) for a in (df_args[0], df_args[2], df_args[3], df_args[4])]
    t_df_k = t_df_h = 0.0
    df_steps = max(3, args.steps // 2)
    for it in range(df_steps + 1):
        hp.pic_upload_s16(df_pic, *df_pre)     # the unfiltered reconstruction (untimed: it is produced on the device)
        flush.fill_(1)
        torch.cuda.synchronize()
        r = L.xb200_deblock(ctx, df_pic, P(df_dev[0]), len(df_args[0]), df_args[1].ctypes.data_as(C.c_void_p), P(df_dev[1]), P(df_dev[2]),
                            P(df_dev[3]), 1, api.MEM_DEVICE)
        if r:
            raise RuntimeError(f"deblock failed: {r}")
        if it:
            t_df_k += hp.last_kernel_ms
        hp.pic_upload_s16(df_pic, *df_pre)
        t0 = time.perf_counter()
        hp.deblock(df_pic, *df_args)           # host-buffer API: CU list + frame maps go in with the call
        if it:
            t_df_h += time.perf_counter() - t0
    f_scu = (W // 4) * (H // 4)
    df_alg = 2 * (W * H * 3 // 2 * 2) + 14 * f_scu + df_args[0].nbytes + 2 * f_scu \
        + 2 * ((W + 288) * (H + 288) - W * H) + 4 * ((W // 2 + 144) * (H // 2 + 144) - W * H // 4)
    deblock = {"cus_per_frame": int(len(df_args[0])), "kernel_ms_per_frame": round(t_df_k / df_steps, 4),
               "host_api_ms_per_frame": round(t_df_h / df_steps * 1e3, 3), "launches_per_frame": 4,
               "algorithmic_bytes": int(df_alg), "achieved_gbs": round(df_alg / (t_df_k / df_steps * 1e-3) / 1e9, 1),
               "note": "mark edges + vertical-edge pass + horizontal-edge pass + one border-expansion grid, L2 flushed before the call; "
                       "random quad-tree down to 4x4, 10 % intra CUs"}
    hp.pic_destroy(df_pic)
    sampler.stop_flag = True
    frame_bytes = W * H * 3 // 2 * bps
    h2d = frame_bytes + fw.me_uni.nbytes + bi_mc.nbytes + fw.bi_cur.nbytes + fw.side_off.nbytes + me_bi_in.nbytes + 2 * fw.side_elems \
        + res_in.nbytes + fw.rates.nbytes
    # coefficient planes travel compacted (only planes with a non-zero level): count what really crosses the bus
    wsq = res_out["mc"]["w"].astype(np.int64) ** 2
    coef_bytes = int(2 * (np.stack([wsq, wsq // 4, wsq // 4], 1) * (res_out["nnz"] != 0)).sum()) + 16 * int((res_out["nnz"] != 0).sum())
    d2h = fw.me_uni.nbytes + 2 * fw.side_elems + me_bi_in.nbytes + res_in.nbytes + coef_bytes

    if world > 1:
        t_dev, t_e2e = xd.max_over_ranks([t_dev, t_e2e], dist, device="cuda")
    if rank != 0:
        hp.close()
        return None
    ms_step = t_dev / args.steps * 1e3
    value = world * args.steps / t_dev
    e2e_val = world * e2e_steps / t_e2e
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg = algorithmic_bytes(fw)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
        traffic = json.load(open(tpath)).get(f"{CLIP}", {})
    per_stage = {k: v / args.steps for k, v in stage_ms.items()}
    dom = max(per_stage, key=per_stage.get)
    achieved = alg[dom] / (per_stage[dom] * 1e-3) / 1e9
    out = {
        "metric": f"hot-path operator pictures/s ({CLIP}, preset {PRESET}): ME + bi-org + residue work lists of one B picture, frozen dependent lists, nothing is encoded", "value": round(value, 3),
        "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s16 samples, s32/s64 arithmetic", "data": "synthetic",
        "config": {"workload": f"{W}x{H} B picture (POC 8 <- POC 0/16), Baseline preset {PRESET}, full 64/32/16/8 quad-tree: "
                               f"{cnt['me_uni']} uni ME + {cnt['me_bi']} bi ME + {cnt['bi_org']} bi_org + {cnt['residue']} residue items per frame",
                   "clip": f"seeded synthetic {W}x{H} {clip.depth}-bit (xeve_b200/clips.py)", "mvp": "synthetic (true motion + jitter)",
                   "parallelism": f"{world} independent picture streams (one per GPU), header broadcast only",
                   "l2": "L2 flushed (256 MB write) before every timed step" },
        "e2e": {"value": round(e2e_val, 3), "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps},
        "gpu_launches": int(launches),
        "kernel_ms_per_step": {k: round(v, 3) for k, v in per_stage.items()},
        "analyze_cu": analyze,
        "intra": intra,
        "deblock": deblock,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 5), "traffic": (traffic or {}).get(dom),
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                     "algorithmic_bytes_per_launch": int(alg[dom]),
                     "note": "integer ALU / shared-memory bound, not HBM bound (ncu: issue slots 48-60 % active, DRAM 0.4-0.5 % of peak for k_me<3..6>; "
                             "profiles/r01_final_ncu_summary.txt): see DESIGN.md section 5"},
        "clocks": sampler.summary(),
    }
    hp.close()
    return out


def run_reference(args, sample_rows=SAMPLE_ROWS, steps=None, quiet=False):
    """The reference's own CPU implementation of the path (oracle/_ref replay, all host threads) on a
    bounded sample: the CUs of the first `sample_rows` CTU rows of the same picture."""
    from oracle import refharness as rh
    from xeve_b200 import api
    from xeve_b200.clips import to_internal10
    from xeve_b200.worklist import FrameWork
    from tests_support import padded_planes_struct  # noqa: F401  (defined below, registered in sys.modules)

    if not rh.available():
        return {"impl": "reference", "unavailable": "oracle/_ref (compiled reference) is not present on this machine"}
    clip, fr = frames_for_bench()
    seq = api.make_seq(W, H, PRESET)
    keep, planes = padded_planes_struct(fr, [REF_POCS[0], REF_POCS[1], POC], clip.depth)
    mr = int(seq["me_range"][0])
    fw = FrameWork(W, H, POC, REF_POCS, PAN, 2, [0, 1], rows=sample_rows, me_range=mr)
    cores = usable_cores()
    frac = fw.n_cu / FrameWork(W, H, POC, REF_POCS, PAN, 2, [0, 1], me_range=mr).n_cu
    rows_total = (H + 63) // 64
    steps = steps or args.steps
    times = []
    for it in range(0 if args.warmup else 1, steps + 1):   # iteration 0 is the (untimed) warm-up pass
        t0 = time.perf_counter()
        me_uni, s1 = rh.replay_me_raw(seq, planes, None, fw.me_uni.astype(rh.ME_REC), cores)
        bi_mc, me_bi_in = fw.build_bi(me_uni.astype(api.ME_ITEM))
        pred, off, s2 = rh.replay_mc_raw(seq, planes, bi_mc.astype(rh.MC_REC), cores)
        side = np.zeros(fw.side_elems, np.int16)  # get_org_bi (untimed bookkeeping, trivial next to the search)
        cy = to_internal10(fr[POC][0], clip.depth)
        for i in range(fw.n_cu):
            s = 1 << int(fw.l2[i])
            blk = cy[fw.y[i]:fw.y[i] + s, fw.x[i]:fw.x[i] + s].astype(np.int32) * 2 - pred[off[i]:off[i] + s * s].reshape(s, s)
            side[fw.side_off[i]:fw.side_off[i] + s * s] = blk.reshape(-1)
        me_bi, s3 = rh.replay_me_raw(seq, planes, side, me_bi_in.astype(rh.ME_REC), cores)
        res_in = fw.build_residue(me_bi.astype(api.ME_ITEM))
        _, _, _, s4 = rh.replay_residue(seq, planes, fw.rates, res_in.astype(rh.RES_REC), fw.res_elems, cores)
        if it > 0:
            times.append(s1 + s2 + s3 + s4)
        _ = time.perf_counter() - t0
    sec = float(np.mean(times)) / frac  # scaled to a whole picture
    value = 1.0 / sec
    out = {"impl": "reference", "metric": f"hot-path operator pictures/s ({CLIP}, preset {PRESET}): ME + bi-org + residue work lists of one B picture, frozen dependent lists, nothing is encoded",
           "value": round(value, 4), "unit": "frames/s", "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
           "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "s16 samples, s32/s64 arithmetic", "data": "synthetic",
           "config": {"workload": f"same {W}x{H} B-picture work lists as the b200 arm", "sample": f"first {sample_rows} of {rows_total} CTU rows "
                      f"({fw.n_cu} CUs = {frac:.3f} of the picture), time scaled to the whole picture"},
           "cpu_baseline": {"value": round(value, 4), "unit": "frames/s", "cores": cores, "kind": "reference",
                            "sample": f"{sample_rows}/{rows_total} CTU rows, reference AVX2 functions replayed on {cores} host threads "
                                      f"(usable CPUs: affinity {len(os.sched_getaffinity(0))}, cgroup quota applied)"},
           "e2e": {"value": round(value, 4), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    return out


def reference_cu_rate(frames=3):
    """Reference side of the analyze_cu line: a real (single-thread, parity-configuration) encode of the first pictures of
    the same clip by the compiled reference with a stop-watch around its own xeve_pinter_analyze_cu (oracle/ref_harness.c,
    trace mask TRACE_CU_TIME) -> CUs per second per host core."""
    from oracle import refharness as rh
    from xeve_b200.clips import Clip
    if not rh.available():
        return {"unavailable": "oracle/_ref not present"}
    c = Clip(CLIP)
    dt = np.uint8 if c.depth == 8 else np.dtype("<u2")
    yuv = np.frombuffer(b"".join(c.frame_bytes(i) for i in range(frames)), dt)
    rh.encode_clip(yuv, frames, c.w, c.h, in_depth=c.depth, preset=PRESET, trace_mask=rh.TRACE_CU_TIME, pic_lo=1, pic_hi=1 << 30,
                   want_bitstream=False)
    sec, calls = rh.cu_time()
    cores = usable_cores()
    rate = calls / sec if sec > 0 else 0.0
    return {"cus_per_s_per_core": round(rate, 1), "calls": int(calls), "seconds": round(sec, 3), "cores": cores,
            "cus_per_s_all_cores_ideal": round(rate * cores, 1), "kind": "reference",
            "sample": f"xeve_pinter_analyze_cu inside a real 1-thread encode of {frames} {c.w}x{c.h} pictures (inter pictures only)"}


def reference_intra_rate(frames=1):
    """Reference side of the intra line: a real single-thread encode of the first (intra) picture of the same clip with a
    stop-watch around the reference's own pintra_analyze_cu (oracle/ref_harness.c, TRACE_INTRA_TIME) -> CUs / s per host core."""
    from oracle import refharness as rh
    from xeve_b200.clips import Clip
    if not rh.available():
        return {"unavailable": "oracle/_ref not present"}
    c = Clip(CLIP)
    dt = np.uint8 if c.depth == 8 else np.dtype("<u2")
    yuv = np.frombuffer(b"".join(c.frame_bytes(i) for i in range(frames)), dt)
    rh.encode_clip(yuv, frames, c.w, c.h, in_depth=c.depth, preset=PRESET, trace_mask=rh.TRACE_INTRA_TIME, pic_lo=0, pic_hi=1 << 30,
                   want_bitstream=False)
    sec, calls = rh.intra_time()
    rate = calls / sec if sec > 0 else 0.0
    return {"cus_per_s_per_core": round(rate, 1), "calls": int(calls), "seconds": round(sec, 3), "cores": usable_cores(),
            "cus_per_s_all_cores_ideal": round(rate * usable_cores(), 1), "kind": "reference",
            "sample": f"pintra_analyze_cu inside a real 1-thread encode of the intra picture of the {c.w}x{c.h} clip"}


def reference_deblock():
    """The reference's own edge filters (xeve_deblock_cu_ver / _hor, one thread -- xeve_loop_filter is single-threaded with one
    tile) over the same synthetic picture as the b200 arm's deblock line."""
    from oracle import refharness as rh
    from xeve_b200.worklist import synth_deblock
    if not rh.available():
        return {"unavailable": "oracle/_ref not present"}
    df = synth_deblock(W, H, seed=0)
    secs = [rh.deblock(df["pre"], df["cus"].astype(rh.DF_CU), df["pp"], df["map_scu"], df["map_refi"], df["map_mv"])[1] for _ in range(3)]
    return {"ms_per_frame": round(min(secs) * 1e3, 3), "cores": 1, "kind": "reference",
            "sample": "whole picture, best of 3 (filter only; the reference's border expansion is not included)"}


# ---- helper module (kept here so bench.py is self-contained) -----------------------------------------------------
import types  # noqa: E402

_ts = types.ModuleType("tests_support")


def _padded_planes_struct(fr, pocs, depth=8):
    """Internal-depth, edge-padded copies of frames + a ctypes PLANES array for the reference harness."""
    from oracle import refharness as rh
    from xeve_b200.clips import to_internal10
    keep, arr = [], (rh.PLANES * len(pocs))()
    for i, poc in enumerate(pocs):
        bufs = []
        for k, p in enumerate(fr[poc]):
            pad = 144 if k == 0 else 72
            bufs.append(np.ascontiguousarray(np.pad(to_internal10(p, depth), pad, mode="edge")))
        keep.append(bufs)
        arr[i].y = bufs[0].ctypes.data + 2 * (144 * bufs[0].shape[1] + 144)
        arr[i].u = bufs[1].ctypes.data + 2 * (72 * bufs[1].shape[1] + 72)
        arr[i].v = bufs[2].ctypes.data + 2 * (72 * bufs[2].shape[1] + 72)
        arr[i].s_l, arr[i].s_c, arr[i].w_l, arr[i].h_l, arr[i].poc = bufs[0].shape[1], bufs[1].shape[1], W, H, poc
    return keep, arr


_ts.padded_planes_struct = _padded_planes_struct
sys.modules["tests_support"] = _ts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="1080p", choices=["1080p", "2160p"], help="1080p fast (default, BASELINE configs[1]) or 2160p 10-bit medium")
    args = ap.parse_args()
    # exactly ONE line on stdout: everything libraries print (NCCL's version banner, torchrun notices) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()
    select_workload(args.workload)
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    if args.impl == "reference":
        if rank == 0:
            a = argparse.Namespace(**vars(args))
            a.steps = min(args.steps, 3)
            emit(run_reference(a))
        return
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import torch
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
    out = run_b200(args, rank, world, dist)
    if rank == 0:
        if world == 1:
            a = argparse.Namespace(**vars(args))
            a.steps, a.warmup = 1, 1
            ref = run_reference(a)
            out["cpu_baseline"] = ref.get("cpu_baseline", {"unavailable": ref.get("unavailable")})
            out["deblock"]["cpu_reference"] = reference_deblock()
            iref = reference_intra_rate()
            out["intra"]["cpu_reference"] = iref
            if iref.get("cus_per_s_per_core"):
                out["intra"]["host_cores_equivalent"] = round(out["intra"]["cus_per_s_kernel"] / iref["cus_per_s_per_core"], 1)
            cur = reference_cu_rate()
            out["analyze_cu"]["cpu_reference"] = cur
            if cur.get("cus_per_s_per_core"):
                gpu_rate = out["analyze_cu"]["cus_per_frame"] * out["analyze_cu"]["frames_per_s_kernel"]
                out["analyze_cu"]["cus_per_s_kernel"] = round(gpu_rate, 1)
                out["analyze_cu"]["host_cores_equivalent"] = round(gpu_rate / cur["cus_per_s_per_core"], 1)
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
