"""Small deterministic fuzz clips for parity tests (edge cases the reference's own behaviour
exposes: odd widths with row padding, flat and duplicate frames, noise -> Cx2/3/6/7 promotion,
gradients -> predictor type 4, scrolls / shifts -> motion vectors, tiny frames)."""
from __future__ import annotations

import numpy as np

from screenpressor_b200.synth import _hash32


def _noise(shape, salt):
    n = int(np.prod(shape))
    return (_hash32(np.arange(n, dtype=np.uint64) + salt * 1000003) & 255).astype(np.uint8).reshape(shape)


def fuzz_clip(w: int, h: int, n: int, seed: int, bpp: int = 32, levels: int = 256):
    """Returns (frames, keyflags).  frames: (n, h, w, 4) for 32 bpp or (n, h, stride) for 24 bpp.
    `levels` < 256 quantises the noise so contexts see repeats early (SmallContext / Cx6 paths)."""
    rgb = np.zeros((h, w, 3), dtype=np.uint8)
    out = []
    keys = np.zeros(n, dtype=np.uint8)
    keys[0] = 1
    for f in range(n):
        r = int(_hash32(seed * 7919 + f * 31 + 5)) % 12
        s = seed * 1000 + f
        if f == 0 or r == 0:  # fresh mixed content
            rgb[:] = _noise((1, 1, 3), s)
            x0, y0 = w // 4, h // 4
            nz = _noise((h - y0, w - x0, 3), s + 1)
            if levels < 256:
                nz = (nz % levels) * (256 // levels)
            rgb[y0:, x0:] = nz
            yy, xx = np.mgrid[0 : h // 2, 0 : w // 2]
            rgb[: h // 2, : w // 2, 0] = (xx * 3 + yy) & 255
            rgb[: h // 2, : w // 2, 1] = (xx + yy * 2) & 255
            rgb[: h // 2, : w // 2, 2] = (xx * 2) & 255
            if r == 0 and f > 0 and (s & 1):
                keys[f] = 1
        elif r == 1:  # flat frame
            rgb[:] = _noise((1, 1, 3), s)
        elif r == 2:  # exact duplicate
            pass
        elif r in (3, 4):  # vertical / horizontal shift of the whole frame
            dy = int(_hash32(s + 11)) % 37 - 18
            dx = int(_hash32(s + 12)) % 23 - 11 if r == 4 else 0
            rgb[:] = np.roll(rgb, (dy, dx), axis=(0, 1))
        elif r in (5, 6):  # random rectangle of a few colours
            x1 = int(_hash32(s + 1)) % w; y1 = int(_hash32(s + 2)) % h
            x2 = min(w, x1 + 1 + int(_hash32(s + 3)) % 70); y2 = min(h, y1 + 1 + int(_hash32(s + 4)) % 50)
            pal = _noise((4, 3), s + 5)
            idx = _noise((y2 - y1, x2 - x1), s + 6) % (1 + int(_hash32(s + 7)) % 4)
            rgb[y1:y2, x1:x2] = pal[idx]
        elif r == 7:  # single pixel
            rgb[int(_hash32(s + 1)) % h, int(_hash32(s + 2)) % w] ^= 0x55
        elif r == 8:  # noise patch
            x1 = int(_hash32(s + 1)) % w; y1 = int(_hash32(s + 2)) % h
            x2 = min(w, x1 + 40); y2 = min(h, y1 + 30)
            nz = _noise((y2 - y1, x2 - x1, 3), s + 3)
            if levels < 256:
                nz = (nz % levels) * (256 // levels)
            rgb[y1:y2, x1:x2] = nz
        elif r == 9:  # shift a sub-pane only (MV blocks next to literal blocks)
            x1, x2, y1, y2 = w // 8, w - w // 8, h // 8, h - h // 8
            rgb[y1:y2, x1:x2] = np.roll(rgb[y1:y2, x1:x2], -16 if s & 1 else 5, axis=0)
        elif r == 10:  # keyframe request on unchanged content
            keys[f] = 1
        else:  # first / last row and column touched (edge predictors)
            rgb[0, :] = _noise((w, 3), s + 1)
            rgb[:, 0] = _noise((h, 3), s + 2)
            rgb[-1, :] ^= 0x0F
        if bpp == 32:
            fr = np.empty((h, w, 4), dtype=np.uint8)
            fr[..., :3] = rgb
            fr[..., 3] = 255
        else:
            stride = (w * 3 + 3) & ~3
            fr = np.zeros((h, stride), dtype=np.uint8)
            fr[:, : w * 3] = rgb.reshape(h, w * 3)
        out.append(fr)
    return np.stack(out), keys


def motion_clip(w: int, h: int, n: int, seed: int):
    """noise-textured window dragged over a textured desktop and dropped now and then: many motion-vector blocks next
    to pixel-coded ones, and non-zero vectors left behind in the reference's persistent mvs[] array (32 bpp)."""
    rng = np.random.default_rng(seed)
    bg = rng.integers(0, 4, (h, w, 1), dtype=np.uint8) * 60 + np.zeros((1, 1, 3), np.uint8)
    win = rng.integers(0, 256, (h // 3, w // 3, 3), dtype=np.uint8)
    clip = np.zeros((n, h, w, 4), np.uint8)
    clip[..., 3] = 255
    for i in range(n):
        f = bg.copy()
        x, y = (16 + 5 * i) % (w - w // 3 - 1), (8 + 3 * i) % (h - h // 3 - 1)
        if i % 12 < 9:
            f[y:y + h // 3, x:x + w // 3] = win
        f[(7 * i) % h, (11 * i) % w] = (i, 255 - i, 3 * i % 256)
        clip[i, ..., :3] = f
    return clip


def to_rgb555(clip32: np.ndarray) -> np.ndarray:
    """(n, h, w, 4) BGRA bytes -> (n, h, w) uint16 words, 5 bits per channel under the masks 0x7C00 / 0x3E0 / 0x1F"""
    c = clip32.astype(np.uint16)
    return ((c[..., 0] >> 3) << 10) | ((c[..., 1] >> 3) << 5) | (c[..., 2] >> 3)


def band_clip(w: int, h: int, n: int, seed: int, bpp: int = 32) -> np.ndarray:
    """flat background, rectangles, a gradient column band and a noise patch: runs that continue across row ends, so the row bands
    of the multi-threaded reference (one new run per band) change the bytes of an I frame"""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        rgb = np.zeros((h, w, 3), np.uint8)
        rgb[:] = rng.integers(0, 256, 3, dtype=np.uint8)
        yy, xx = np.mgrid[0:h, 0:w]
        rgb[:, w // 3:w // 2, 0] = (xx[:, w // 3:w // 2] * 2 + yy[:, w // 3:w // 2]) & 255
        for _ in range(6):
            x1, y1 = int(rng.integers(0, w - 8)), int(rng.integers(0, h - 8))
            rgb[y1:y1 + int(rng.integers(2, h // 3)), x1:x1 + int(rng.integers(2, w // 3))] = rng.integers(0, 256, 3, dtype=np.uint8)
        x1, y1 = int(rng.integers(0, w - 40)), int(rng.integers(0, h - 30))
        rgb[y1:y1 + 30, x1:x1 + 40] = rng.integers(0, 16, (30, 40, 3), dtype=np.uint8) * 16
        if bpp == 32:
            fr = np.full((h, w, 4), 255, np.uint8)
            fr[..., :3] = rgb
        else:
            st = (w * 3 + 3) & ~3
            fr = np.zeros((h, st), np.uint8)
            fr[:, :w * 3] = rgb.reshape(h, w * 3)
        out.append(fr)
    return np.stack(out)
