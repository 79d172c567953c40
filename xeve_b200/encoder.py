"""Clip-level driver of the device decision pass: the host side of "one frame's CTUs run as one grid" for a whole sequence.

Given the picture-level plan of a clip (slice types, POCs, QPs, lambdas, reference lists per picture in coding order -- what the
encoder's control plane computes; with constant QP it does not depend on any decision) a ClipEncoder uploads the original pictures,
enqueues EVERY picture at once (xb200_analyze_picture: the library orders them by events on their reference pictures, so the pictures
of one wave of the picture DAG run concurrently) and hands each picture's records (per-CTU unit records + coefficient planes, the
contents of the reference's ctx->map_cu_data[]) to the entropy coder in coding order as they complete.  The entropy coder and the
bitstream writer stay on the host (north_star); this module never touches them -- it only produces what they read.

Plumbing only: no arithmetic of the path happens here.
"""
from __future__ import annotations

import threading
import time

import numpy as np

from . import api

PLAN_FIELDS = ("poc", "slice_type", "tile_qp", "num_refp", "ref_poc", "col_list_poc0", "max_cu_inter", "min_cu_inter", "max_cu_intra",
               "min_cu_intra", "cip", "qp", "lambda_mv", "max_search_range", "lambda", "sqrt_lambda0", "dist_chroma_weight")


def picture_record(pp, df_pp, cur_pic, rec_pic, ref_handles, deblock=1, threads=None, unfiltered=-1):
    """xb200_picture from the picture-level fields of one plan entry (`pp`: any record with PLAN_FIELDS + ref_pic validity markers +
    parallel_rows) and the loop-filter parameters (`df_pp`: DF_PIC); ref_handles: POC -> device handle of the reconstructed picture."""
    pp = np.asarray(pp).reshape(-1)[0]
    p = np.zeros(1, api.PICTURE)
    for k in PLAN_FIELDS:
        p[k] = pp[k]
    p["parallel_rows"] = int(pp["parallel_rows"]) if threads is None else threads
    p["cur_pic"], p["rec_pic"], p["unfiltered_pic"], p["deblock"] = cur_pic, rec_pic, unfiltered, deblock
    p["ref_pic"] = -1
    for l in range(2):
        for k in range(4):
            if int(pp["ref_pic"][l][k]) >= 0:
                p["ref_pic"][0][l][k] = ref_handles[int(pp["ref_poc"][l][k])]
    p["df"] = np.asarray(df_pp).reshape(-1)[0]
    return p


class ClipEncoder:
    """One stream on one device context.

    enc = ClipEncoder(seq, plan, device=0, threads=None)
    enc.upload(frames, in_depth)        # frames[poc] = (y, u, v) planes of the caller (u8 or u16); H2D + depth conversion on the device
    enc.start()                         # enqueue every picture of the plan (returns at once unless the device is full)
    rec = enc.fetch(poc)                # blocks until that picture is decided: dict(scu, coef, stat)
    enc.close()
    """

    def __init__(self, seq, plan, device=0, threads=None, hp=None):
        self.hp = hp if hp is not None else api.Hotpath(seq, device=device)
        self.own = hp is None
        self.plan, self.threads = plan, threads
        self.h_org, self.h_rec = {}, {}
        self.enqueued = {int(np.asarray(p["pp"]).reshape(-1)[0]["poc"]): threading.Event() for p in plan}
        self.err = None
        self.t_enqueue = None
        for p in plan:                                   # every device picture up front: the enqueue thread creates nothing
            poc = int(np.asarray(p["pp"]).reshape(-1)[0]["poc"])
            self.h_org[poc] = self.hp.pic_create(padded=False)
            self.h_rec[poc] = self.hp.pic_create(padded=True)

    def upload(self, frames, in_depth):
        for poc, (y, u, v) in frames.items():
            if poc in self.h_org:
                self.hp.pic_upload(self.h_org[poc], y, u, v, in_depth)

    def upload_s16(self, frames):
        for poc, (y, u, v) in frames.items():
            if poc in self.h_org:
                self.hp.pic_upload_s16(self.h_org[poc], y, u, v)

    def reset(self):
        """before re-encoding the same plan with the same device pictures (every picture of the previous pass must have been fetched)"""
        self.err = None
        for ev in self.enqueued.values():
            ev.clear()

    def enqueue(self, k):
        """enqueue the k-th picture of the plan (coding order); returns at once unless the device is full"""
        try:
            p = self.plan[k]
            pp = np.asarray(p["pp"]).reshape(-1)[0]
            poc = int(pp["poc"])
            rec = picture_record(pp, p["df_pp"], self.h_org[poc], self.h_rec[poc], self.h_rec, deblock=int(p.get("deblock", 1)),
                                 threads=self.threads)
            self.hp.analyze_picture(rec)
            self.enqueued[poc].set()
        except Exception as e:  # noqa: BLE001 -- reported by fetch() / wait()
            self.err = e
            for ev in self.enqueued.values():
                ev.set()
            raise

    def _enqueue_all(self):
        try:
            for k in range(len(self.plan)):
                self.enqueue(k)
        except Exception:  # noqa: BLE001 -- already recorded in self.err
            pass

    def start(self, background=True):
        self.reset()
        if background:
            self.t_enqueue = threading.Thread(target=self._enqueue_all, daemon=True)
            self.t_enqueue.start()
        else:
            self._enqueue_all()
            if self.err:
                raise self.err

    def fetch(self, poc, want_states=False):
        """blocks until picture `poc` is decided: dict(scu, coef, states, cost, stat)"""
        self.enqueued[poc].wait()
        if self.err:
            raise self.err
        return self.hp.picture_fetch(self.h_rec[poc], want_states=want_states)

    def wait(self, poc):
        """blocks until picture `poc` is decided, filtered and border-expanded; the records are dropped -> its xb200_picture_stat"""
        self.enqueued[poc].wait()
        if self.err:
            raise self.err
        return self.hp.picture_wait(self.h_rec[poc])

    def reconstruction(self, poc):
        """deblocked picture `poc` (Y, U, V s16 active areas) from the device"""
        return self.hp.pic_download(self.h_rec[poc], False)

    def close(self):
        if self.t_enqueue is not None:
            self.t_enqueue.join()
            self.t_enqueue = None
        if self.own:
            self.hp.close()
