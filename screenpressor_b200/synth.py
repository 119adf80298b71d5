"""Deterministic synthetic screencast clips (BASELINE.json `configs`).

One generator feeds the CUDA path, the oracle and the compiled reference, so every arm sees the
same bytes.  All pseudo-randomness comes from the `lowbias32` integer hash below (pure uint32
numpy arithmetic), so frames do not depend on numpy's RNG implementation.

Frames are BGRA with alpha = 255 (the codec drops alpha on encode and writes 255 on decode,
reference screencap.cpp:1657-1659, 1721), rows top-to-bottom, no row padding for 32 bpp.
24 bpp frames use the reference's row pitch `(3*W + 3) & ~3` (screencap.cpp:75) with zero padding.

Keyframe schedules follow the VfW policy `npframes + 1 >= interval` (screenpressor.cpp:402-406):
frame 0 and every `interval`-th frame after the previous keyframe are requested as I-frames.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterator, List

import numpy as np

__all__ = ["ClipConfig", "CONFIGS", "clip_frames", "keyframe_flags", "make_clip", "make_clip_range"]


def _hash32(x: np.ndarray | int) -> np.ndarray:
    x = np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x.astype(np.uint32)


def _h(*keys: int) -> int:
    """Scalar hash of a tuple of ints."""
    v = 0x9E3779B9
    for k in keys:
        v = int(_hash32((v ^ (int(k) & 0xFFFFFFFF)) & 0xFFFFFFFF)) + 0x9E3779B9
        v &= 0xFFFFFFFF
    return int(_hash32(v))


@dataclass(frozen=True)
class ClipConfig:
    name: str
    width: int
    height: int
    bpp: int  # 24 or 32
    frames: int
    key_interval: int  # 1 = intra only
    seed: int
    kind: str  # desktop | ide | photo | multimon


# BASELINE.json `configs`, in order (cfg1..cfg5 of SURVEY.md §8(d)).
CONFIGS = {
    "cfg1_720p_rgb24": ClipConfig("cfg1_720p_rgb24", 1280, 720, 24, 60, 500, 1, "desktop"),
    "cfg2_1080p_rgb32": ClipConfig("cfg2_1080p_rgb32", 1920, 1080, 32, 600, 500, 2, "desktop_drag_scroll"),
    "cfg3_2160p_rgb32": ClipConfig("cfg3_2160p_rgb32", 3840, 2160, 32, 3600, 450, 3, "ide"),
    "cfg4_1440p_intra": ClipConfig("cfg4_1440p_intra", 2560, 1440, 32, 120, 1, 4, "photo"),
    "cfg5_5120x1440": ClipConfig("cfg5_5120x1440", 5120, 1440, 32, 600, 500, 5, "multimon"),
}


def keyframe_flags(n: int, interval: int) -> np.ndarray:
    """1 where the host requests an I-frame (screenpressor.cpp:402-406 with ForceInterval)."""
    flags = np.zeros(n, dtype=np.uint8)
    npframes = 0
    for i in range(n):
        if i == 0 or npframes + 1 >= interval:
            flags[i] = 1
            npframes = 0
        else:
            npframes += 1
    return flags


# ------------------------------------------------------------------------------------------
# glyph font: 96 glyphs of 8x16, ink only in the inner 6x12 box
# ------------------------------------------------------------------------------------------
def _font(seed: int) -> np.ndarray:
    idx = np.arange(96 * 16 * 8, dtype=np.uint64) + (seed * 7919)
    bits = (_hash32(idx) % 100) < 38
    f = bits.reshape(96, 16, 8)
    f[:, :2, :] = False
    f[:, 14:, :] = False
    f[:, :, 0] = False
    f[:, :, 7] = False
    f[0] = False  # glyph 0 is a space
    return f


class _TextPane:
    """A text area made of 8x16 glyph cells.  `lines` holds one integer id per text row; a row's
    pixels are a pure function of its id, so scrolling is a shift plus one newly rendered row."""

    def __init__(self, x: int, y: int, w: int, h: int, bg, fgs, seed: int, font: np.ndarray):
        self.x, self.y = x, y
        self.w, self.h = (w // 8) * 8, (h // 16) * 16
        self.cols, self.rows = self.w // 8, self.h // 16
        self.bg = np.array(bg, dtype=np.uint8)
        self.fgs = np.array(fgs, dtype=np.uint8).reshape(-1, 3)
        self.seed = seed
        self.font = font
        self.lines: List[int] = [_h(seed, 1000 + r) for r in range(self.rows)]
        self.next_line = self.rows

    def render_line(self, line_id: int) -> np.ndarray:
        cols = self.cols
        c = np.arange(cols, dtype=np.uint64)
        hv = _hash32(c * 2654435761 + line_id)
        fill = 40 + (_h(line_id, 17) % 61)  # 40..100 % of the columns carry text
        ncol = (cols * fill) // 100
        indent = _h(line_id, 23) % 9
        glyph = (hv % 96).astype(np.int64)
        glyph[(hv >> 8) % 6 == 0] = 0  # spaces between words
        glyph[c >= ncol] = 0
        glyph[c < indent] = 0
        color = ((hv >> 16) % len(self.fgs)).astype(np.int64)
        ink = self.font[glyph]  # cols,16,8
        ink = ink.transpose(1, 0, 2).reshape(16, cols * 8)
        out = np.empty((16, cols * 8, 3), dtype=np.uint8)
        out[:] = self.bg
        colpix = np.repeat(color, 8)
        fg = self.fgs[colpix]  # (w,3)
        out[ink] = np.broadcast_to(fg, (16, cols * 8, 3))[ink]
        return out

    def draw_all(self, screen: np.ndarray) -> None:
        for r, lid in enumerate(self.lines):
            self.draw_row(screen, r)

    def draw_row(self, screen: np.ndarray, r: int) -> None:
        y = self.y + r * 16
        screen[y : y + 16, self.x : self.x + self.w, :3] = self.render_line(self.lines[r])

    def scroll(self, screen: np.ndarray) -> None:
        """Scroll up by one text row (16 px) and append a fresh line at the bottom."""
        x, y, w, h = self.x, self.y, self.w, self.h
        screen[y : y + h - 16, x : x + w] = screen[y + 16 : y + h, x : x + w].copy()
        self.lines = self.lines[1:] + [_h(self.seed, 1000 + self.next_line)]
        self.next_line += 1
        self.draw_row(screen, self.rows - 1)

    def edit(self, screen: np.ndarray, row: int, salt: int) -> None:
        row %= self.rows
        self.lines[row] = _h(self.seed, 555, salt, row)
        self.draw_row(screen, row)


def _cursor_sprite() -> np.ndarray:
    """12x18 arrow; 0 transparent, 1 black outline, 2 white fill."""
    s = np.zeros((18, 12), dtype=np.uint8)
    for r in range(16):
        wid = min(r + 1, 11) if r < 11 else max(1, 16 - r)
        s[r, :wid] = 2
        s[r, 0] = 1
        s[r, wid - 1] = 1
    s[10, :11] = np.where(s[10, :11] > 0, 1, 0)
    return s


def _draw_cursor(frame: np.ndarray, sprite: np.ndarray, cx: int, cy: int) -> None:
    H, W = frame.shape[:2]
    h, w = sprite.shape
    x2, y2 = min(cx + w, W), min(cy + h, H)
    if x2 <= cx or y2 <= cy:
        return
    sp = sprite[: y2 - cy, : x2 - cx]
    reg = frame[cy:y2, cx:x2, :3]
    reg[sp == 1] = (0, 0, 0)
    reg[sp == 2] = (255, 255, 255)


def _desktop(screen: np.ndarray, x0: int, y0: int, w: int, h: int, seed: int, font: np.ndarray, dark: bool = False):
    """Paint a desktop (background, two text windows with title bars, taskbar) into the rectangle
    and return its text panes (main pane first)."""
    bgc = (58, 110, 165) if not dark else (30, 30, 30)
    screen[y0 : y0 + h, x0 : x0 + w, :3] = bgc[::-1]
    # taskbar
    tb = 40
    screen[y0 + h - tb : y0 + h, x0 : x0 + w, :3] = (48, 48, 52)
    for k in range(8):
        bx = x0 + 8 + k * 56
        if bx + 40 < x0 + w:
            col = (60 + 20 * (_h(seed, k) % 8), 90 + 10 * (_h(seed, k, 1) % 12), 120 + 8 * (_h(seed, k, 2) % 14))
            screen[y0 + h - tb + 6 : y0 + h - 6, bx : bx + 40, :3] = col
    panes = []
    # window 1 (main, left) and window 2 (right)
    geo = [
        (x0 + (w * 3) // 100, y0 + (h * 5) // 100, (w * 50) // 100, (h * 75) // 100),
        (x0 + (w * 58) // 100, y0 + (h * 12) // 100, (w * 36) // 100, (h * 60) // 100),
    ]
    pal = [
        ((255, 255, 255), [(0, 0, 0), (160, 0, 0), (0, 0, 160)]),
        ((24, 24, 24), [(220, 220, 220), (86, 182, 194)]),
    ]
    for wi, (wx, wy, ww, wh) in enumerate(geo):
        ww = (ww // 16) * 16
        wx = (wx // 4) * 4
        screen[wy : wy + 24, wx : wx + ww, :3] = (200, 120, 40) if wi == 0 else (90, 90, 96)
        screen[wy + 24 : wy + wh, wx : wx + ww, :3] = (240, 240, 240)
        bg, fgs = pal[wi]
        p = _TextPane(wx + 8, wy + 32, ww - 16, wh - 40, bg[::-1], [f[::-1] for f in fgs], _h(seed, 77, wi), font)
        p.draw_all(screen)
        panes.append(p)
    return panes


def clip_frames(cfg: ClipConfig, n: int | None = None) -> Iterator[np.ndarray]:
    """Yield `n` (default cfg.frames) frames.  32 bpp: (H, W, 4) uint8.  24 bpp: (H, stride) uint8."""
    n = cfg.frames if n is None else n
    W, H = cfg.width, cfg.height
    seed = cfg.seed
    font = _font(seed)
    screen = np.zeros((H, W, 4), dtype=np.uint8)
    screen[..., 3] = 255
    sprite = _cursor_sprite()
    kind = cfg.kind

    drag = None
    noise_rect = None
    tooltip = None
    caret = None
    if kind in ("desktop", "desktop_drag_scroll"):
        panes = _desktop(screen, 0, 0, W, H, seed, font)
        if kind == "desktop_drag_scroll":
            # a small dragged window: 320x200, moves 64x48 px/frame for 30 frames every 100
            drag = {"w": 320, "h": 208, "x": 96, "y": 64}
    elif kind == "ide":
        screen[..., :3] = (30, 30, 30)
        screen[:32, :, :3] = (60, 60, 60)
        screen[:, :320, :3] = (37, 37, 38)
        fgs = [(212, 212, 212), (86, 156, 214), (206, 145, 120), (106, 153, 85), (220, 220, 170)]
        side = _TextPane(16, 48, 288, H - 96, (37, 37, 38)[::-1], [(200, 200, 200)], _h(seed, 5), font)
        side.draw_all(screen)
        main = _TextPane(352, 64, 2400, 1800, (30, 30, 30)[::-1], [f[::-1] for f in fgs], _h(seed, 6), font)
        main.draw_all(screen)
        panes = [main, side]
        # minimap column
        mm = _hash32(np.arange(H * 96, dtype=np.uint64) // 3 + seed).reshape(H, 96)
        screen[:, W - 128 : W - 32, 0] = 30 + (mm % 5) * 20
        screen[:, W - 128 : W - 32, 1] = 30 + ((mm >> 4) % 5) * 20
        screen[:, W - 128 : W - 32, 2] = 30 + ((mm >> 8) % 3) * 20
        caret = (352 + 8 * 40, 64 + 16 * 60)
    elif kind == "photo":
        panes = _desktop(screen, 0, 0, W, H, seed, font)
        # 1280x720 "photo": smooth gradients + 3-bit noise (static)
        py, px, ph, pw = 40, 1240, 720, 1280
        yy, xx = np.mgrid[0:ph, 0:pw]
        nz = _hash32((yy * pw + xx).astype(np.uint64) + seed * 131).astype(np.int64)
        screen[py : py + ph, px : px + pw, 0] = ((xx * 255) // pw + (nz & 7)).clip(0, 255)
        screen[py : py + ph, px : px + pw, 1] = ((yy * 255) // ph + ((nz >> 3) & 7)).clip(0, 255)
        screen[py : py + ph, px : px + pw, 2] = (((xx + yy) * 255) // (pw + ph) + ((nz >> 6) & 7)).clip(0, 255)
        noise_rect = (800, 560, 360, 640)  # y, x, h, w: white noise redrawn each frame
    elif kind == "multimon":
        half = W // 2
        panes = _desktop(screen, 0, 0, half, H, seed, font)
        panes += _desktop(screen, half, 0, W - half, H, seed + 100, font, dark=True)
        tooltip = {"x": 3000, "y": 500, "w": 300, "h": 200, "shown": False, "saved": None}
    else:
        raise ValueError(kind)

    stride24 = (W * 3 + 3) & ~3
    prev_out = None
    cur_x, cur_y = 100, 80
    for f in range(n):
        duplicate = False
        if f > 0:
            if kind == "multimon" and _h(seed, 900, f) % 10 == 0:
                duplicate = True  # 10 % of frames are exact duplicates -> 1-byte P frames
            else:
                cur_x = (cur_x + 13) % (W - 12)
                cur_y = (cur_y + 7) % (H - 18)
            if kind in ("desktop", "desktop_drag_scroll") and f % 10 == 0:
                r = _h(seed, 31, f) % panes[1].rows
                panes[1].edit(screen, r, f)
                panes[1].edit(screen, r + 1, f)
            if kind == "desktop_drag_scroll":
                if f % 30 == 0:
                    panes[0].scroll(screen)
            if kind == "ide":
                panes[0].scroll(screen)
            if tooltip is not None and f % 20 == 0 and not duplicate:
                t = tooltip
                reg = screen[t["y"] : t["y"] + t["h"], t["x"] : t["x"] + t["w"], :3]
                if not t["shown"]:
                    t["saved"] = reg.copy()
                    reg[:] = (225, 255, 255)
                    reg[:2] = reg[-2:] = (80, 80, 80)
                    reg[:, :2] = reg[:, -2:] = (80, 80, 80)
                    tp = _TextPane(t["x"] + 8, t["y"] + 8, t["w"] - 16, t["h"] - 16, (225, 255, 255), [(0, 0, 0)], _h(seed, f), font)
                    tp.draw_all(screen)
                else:
                    reg[:] = t["saved"]
                t["shown"] = not t["shown"]
        if noise_rect is not None:
            ny, nx, nh, nw = noise_rect
            nzv = _hash32(np.arange(nh * nw, dtype=np.uint64) + (f * 977 + seed) * 1000003)
            screen[ny : ny + nh, nx : nx + nw, 0] = (nzv & 255).reshape(nh, nw)
            screen[ny : ny + nh, nx : nx + nw, 1] = ((nzv >> 8) & 255).reshape(nh, nw)
            screen[ny : ny + nh, nx : nx + nw, 2] = ((nzv >> 16) & 255).reshape(nh, nw)

        if duplicate and prev_out is not None:
            yield prev_out.copy()
            continue

        frame = screen.copy()
        if drag is not None:
            ph = f % 100
            if 10 <= ph < 40:
                k = ph - 10
                # bounce inside the screen so the window stays fully visible
                span_x, span_y = W - drag["w"] - 64, H - drag["h"] - 64
                px, py = (k * 64) % (2 * span_x), (k * 48) % (2 * span_y)
                px = px if px < span_x else 2 * span_x - px
                py = py if py < span_y else 2 * span_y - py
                dx, dy = 32 + px, 32 + py
                frame[dy : dy + 24, dx : dx + drag["w"], :3] = (40, 120, 200)
                frame[dy + 24 : dy + drag["h"], dx : dx + drag["w"], :3] = (250, 250, 250)
                tp = _TextPane(dx + 8, dy + 32, drag["w"] - 16, drag["h"] - 40, (250, 250, 250), [(0, 0, 0), (0, 110, 0)], _h(seed, 4242), font)
                tp.draw_all(frame)
        if caret is not None and (f // 15) % 2 == 0:
            frame[caret[1] : caret[1] + 16, caret[0] : caret[0] + 2, :3] = (255, 255, 255)
        _draw_cursor(frame, sprite, cur_x, cur_y)
        if cfg.bpp == 24:
            out = np.zeros((H, stride24), dtype=np.uint8)
            out[:, : W * 3] = frame[..., :3].reshape(H, W * 3)
        else:
            out = frame
        prev_out = out
        yield out


def make_clip(cfg: ClipConfig, n: int | None = None) -> np.ndarray:
    """All frames of a clip as one contiguous array: (n, H, W, 4) or (n, H, stride24)."""
    n = cfg.frames if n is None else n
    it = clip_frames(cfg, n)
    first = next(it)
    out = np.empty((n,) + first.shape, dtype=np.uint8)
    out[0] = first
    for i, fr in enumerate(it, start=1):
        out[i] = fr
    return out


def make_clip_range(cfg: ClipConfig, first: int, count: int) -> np.ndarray:
    """Frames [first, first + count) of the clip (the generator is sequential: earlier frames are produced and dropped).
    What one rank of a frame-range split holds."""
    out = None
    for i, fr in enumerate(clip_frames(cfg, first + count)):
        if i < first:
            continue
        if out is None:
            out = np.empty((count,) + fr.shape, dtype=np.uint8)
        out[i - first] = fr
    return out
