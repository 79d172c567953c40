"""Seeded synthetic YUV 4:2:0 clips (SURVEY.md 8d recipe).

No natural video is available offline, so every test / bench input is generated here:
a blurred, contrast-stretched noise background that pans at a constant velocity, a few
textured squares that move with their own velocities, and a little per-frame noise.
The result has real (mostly translational) motion, so motion search, sub-pel
interpolation and the residual transform all see non-trivial data.

Frames are planar I420; 8-bit clips are uint8, 10-bit clips are little-endian uint16,
exactly what the reference app reads with ``-d 8`` / ``-d 10``.
"""
from __future__ import annotations

import numpy as np

# name -> (w, h, frames, seed, blur k, gain, pan (x, y), squares [(size, x0, y0, vx, vy)], bit depth)
CLIPS = {
    "cif": dict(w=352, h=288, n=30, seed=1234, k=5, g=3.0, pan=(2, 1),
                squares=[(48, 40, 60, 3, 2)], depth=8),
    "1080p": dict(w=1920, h=1080, n=300, seed=20260925, k=7, g=4.0, pan=(3, 1),
                  squares=[(128, 200, 150, 5, 2), (96, 1400, 300, -4, 3), (64, 900, 800, 2, -1)], depth=8),
    "2160p10": dict(w=3840, h=2160, n=300, seed=20260926, k=7, g=4.0, pan=(3, 1),
                    squares=[(256, 400, 300, 5, 2), (192, 2800, 600, -4, 3), (128, 1800, 1600, 2, -1)], depth=10),
    "2160p": dict(w=3840, h=2160, n=300, seed=20260930, k=7, g=4.0, pan=(3, 1),
                  squares=[(256, 400, 300, 5, 2), (192, 2800, 600, -4, 3), (128, 1800, 1600, 2, -1)], depth=8),
}


def _box_blur(a: np.ndarray, k: int) -> np.ndarray:
    """k x k box blur with edge replication (integral-image form, float64)."""
    p = k // 2
    ap = np.pad(a.astype(np.float64), p, mode="edge")
    c = np.cumsum(np.cumsum(ap, axis=0), axis=1)
    c = np.pad(c, ((1, 0), (1, 0)))
    h, w = a.shape
    s = c[k:k + h, k:k + w] - c[0:h, k:k + w] - c[k:k + h, 0:w] + c[0:h, 0:w]
    return s / (k * k)


class Clip:
    """Frame generator for one named clip; ``frame(n)`` -> (Y, U, V) arrays."""

    def __init__(self, name: str, **override):
        cfg = dict(CLIPS[name])
        cfg.update(override)
        self.cfg = cfg
        self.w, self.h, self.n, self.depth = cfg["w"], cfg["h"], cfg["n"], cfg["depth"]
        rng = np.random.default_rng(cfg["seed"])
        # background is larger than the frame by the total pan distance (capped, wraps after)
        self.mx = 160 + abs(cfg["pan"][0]) * 64
        self.my = 160 + abs(cfg["pan"][1]) * 64
        bg = rng.integers(0, 256, size=(self.h + self.my, self.w + self.mx), dtype=np.uint8)
        bg = _box_blur(bg, cfg["k"])
        bg = np.clip((bg - 128.0) * cfg["g"] + 128.0, 0, 255)
        self.bg = np.floor(bg + 0.5).astype(np.int32)
        self.seed = cfg["seed"]

    def frame(self, n: int):
        cfg = self.cfg
        ox = (cfg["pan"][0] * n) % self.mx
        oy = (cfg["pan"][1] * n) % self.my
        y = self.bg[oy:oy + self.h, ox:ox + self.w].copy()
        for (s, x0, y0, vx, vy) in cfg["squares"]:
            sx = int(np.clip(x0 + vx * n, 0, self.w - s))
            sy = int(np.clip(y0 + vy * n, 0, self.h - s))
            r = np.arange(s, dtype=np.int32)[:, None]
            c = np.arange(s, dtype=np.int32)[None, :]
            y[sy:sy + s, sx:sx + s] = (5 * r + 3 * c + 7 * n) % 256
        rng = np.random.default_rng(self.seed * 1000003 + n)
        y = np.clip(y + rng.integers(-2, 3, size=y.shape, dtype=np.int32), 0, 255)
        ysub = (y[0::2, 0::2] + y[1::2, 0::2] + y[0::2, 1::2] + y[1::2, 1::2] + 2) >> 2
        u = 128 + ((ysub - 128) >> 2)
        v = 128 - ((ysub - 128) >> 3)
        if self.depth == 8:
            return y.astype(np.uint8), u.astype(np.uint8), v.astype(np.uint8)
        # 10-bit: scale by 4 and fill the two LSBs with seeded noise
        sh = self.depth - 8
        lsb = rng.integers(0, 1 << sh, size=y.shape, dtype=np.int32)
        y10 = (y << sh) + lsb
        u10 = (u << sh) + (lsb[0::2, 0::2] & ((1 << sh) - 1))
        v10 = (v << sh) + (lsb[1::2, 1::2] & ((1 << sh) - 1))
        return y10.astype("<u2"), u10.astype("<u2"), v10.astype("<u2")

    def frame_bytes(self, n: int) -> bytes:
        y, u, v = self.frame(n)
        return y.tobytes() + u.tobytes() + v.tobytes()

    def write(self, path: str, frames: int | None = None) -> str:
        frames = self.n if frames is None else frames
        with open(path, "wb") as f:
            for i in range(frames):
                f.write(self.frame_bytes(i))
        return path


def to_internal10(plane: np.ndarray, in_depth: int) -> np.ndarray:
    """Input sample -> the codec's internal 10-bit s16 sample.

    Mirrors the conversion xeve_push applies (reference src_base/xeve_util.c:1552-1571,
    1679-1681: left shift by codec_bit_depth - input_bit_depth).
    """
    return (plane.astype(np.int16) << (10 - in_depth)).astype(np.int16)
