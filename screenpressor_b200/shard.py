"""Frame-range sharding of one clip across the GPUs of a box (BASELINE north_star).

A clip is cut into contiguous frame ranges at keyframes; rank r encodes (or decodes) its ranges with its own
codec object on its own GPU and the per-frame bitstreams are concatenated on the host.  There is no collective
on the data path.  The only thing a byte-identical encode has to pass from one range to the next is the
reference's persistent motion-vector array (`mvs[]` is never cleared, not even by an I frame,
screencap.cpp:96-97) plus two counters: a 64 KB blob at 1080p (`ScreenCodec.ExportRangeState`).  It is relayed
rank to rank with point-to-point `torch.distributed` sends (NCCL on the GPU box, gloo in the CPU tests);
decoding needs no relay at all, a GOP is self-contained.

The relay makes range k's motion search wait for range k-1's: `encode_sharded` therefore encodes the ranges in
order on each rank and overlaps only what does not depend on the relay when the codec offers it.  The hot path
inside a range is untouched.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class FrameRange:
    first: int   # first frame (a keyframe, or frame 0)
    count: int
    rank: int    # owner


def gop_ranges(keyflags: np.ndarray, not_cuttable=()) -> list[tuple[int, int]]:
    """[(first, count)] of the GOPs of a clip: every keyframe starts one (frame 0 always does), except the frames in
    `not_cuttable` (see flat_keyframes)."""
    n = len(keyflags)
    skip = set(int(i) for i in not_cuttable)
    starts = [0] + [i for i in range(1, n) if keyflags[i] and i not in skip]
    return [(s, (starts[k + 1] if k + 1 < len(starts) else n) - s) for k, s in enumerate(starts)]


def flat_keyframes(frames: np.ndarray, keyflags: np.ndarray) -> list[int]:
    """Requested keyframes that are single-colour frames.  The reference codes such a frame as a 4-byte flat frame and
    does NOT start a new GOP with it: its frame counter keeps running and, when the colour repeats the last flat frame's,
    not even the models are renewed (screencap.cpp:1488-1511) -- so the frames after it are P frames on state a fresh
    codec does not have.  A range must not start there; the planner treats the frame as part of the GOP before it."""
    out = []
    for i in range(1, len(keyflags)):
        if keyflags[i]:
            f = np.asarray(frames[i])
            px = f.reshape(-1, f.shape[-1])[:, :3] if f.ndim == 3 else None
            if px is not None and (px == px[0]).all():
                out.append(i)
    return out


def assign_ranges(keyflags: np.ndarray, world: int, not_cuttable=()) -> list[FrameRange]:
    """Contiguous GOP-aligned ranges, one per rank (fewer when the clip has fewer GOPs than ranks), balanced by
    frame count: GOP boundaries are the only cuts that keep the bitstream identical without moving model state."""
    gops = gop_ranges(keyflags, not_cuttable)
    n = len(keyflags)
    parts = min(world, len(gops))
    out, g = [], 0
    for r in range(parts):
        first = gops[g][0]
        target = n * (r + 1) / parts
        cnt = 0
        # take GOPs while the range end stays closer to the ideal cut, leaving one GOP for every later rank
        while g < len(gops) - (parts - 1 - r):
            end = gops[g][0] + gops[g][1]
            if cnt and abs(end - target) > abs(gops[g][0] - target):
                break
            cnt += gops[g][1]
            g += 1
        out.append(FrameRange(first, cnt, r))
    if g < len(gops):  # remainder goes to the last rank
        last = out[-1]
        out[-1] = FrameRange(last.first, n - last.first, last.rank)
    return out


def encode_sharded(codec, frames, keyflags: np.ndarray, rank: int, world: int, dist=None, device_ptr=None, pipelined: bool = True,
                   not_cuttable=(), stats: dict | None = None):
    """Encode this rank's range of the clip.  `frames`: this rank's frames only (host ndarray, or None with a device
    pointer).  Returns (FrameRange | None, stream, sizes, ftypes) for the range.  `dist` = torch.distributed (already
    initialised) or None for a single process; the mvs[] blob travels rank -> rank + 1.

    pipelined: the blob is received right before this range's in-order motion-vector resolve and sent right after it
    (codec hooks), so only the resolves of consecutive ranges are serialised; everything else -- frame scan, motion
    search, pixel typing, model replay, rANS -- runs concurrently on all ranks.  Otherwise whole ranges are serialised."""
    import torch

    ranges = assign_ranges(keyflags, world, not_cuttable)
    mine = next((r for r in ranges if r.rank == rank), None)
    if mine is None:
        return None, np.zeros(0, np.uint8), np.zeros(0, np.uint32), np.zeros(0, np.uint8)
    blob_len = len(codec.ExportRangeState(False))
    has_prev, has_next = rank > 0 and dist is not None, dist is not None and rank + 1 < len(ranges)

    import time
    t_call = time.perf_counter()

    def note(key, t0):
        if stats is not None:
            stats[key] = stats.get(key, 0.0) + (time.perf_counter() - t0) * 1e3

    def recv_blob():
        if stats is not None:
            stats["to_resolve_ms"] = stats.get("to_resolve_ms", 0.0) + (time.perf_counter() - t_call) * 1e3
        if has_prev:
            t0 = time.perf_counter()
            t = torch.empty(blob_len, dtype=torch.uint8)
            dist.recv(t, src=rank - 1)
            note("recv_wait_ms", t0)
            t0 = time.perf_counter()
            codec.ImportRangeState(t.numpy())
            note("import_ms", t0)
        if stats is not None:
            stats["_t_resolve"] = time.perf_counter()

    def send_blob():
        if stats is not None and "_t_resolve" in stats:
            stats["resolve_ms"] = stats.get("resolve_ms", 0.0) + (time.perf_counter() - stats.pop("_t_resolve")) * 1e3
        if has_next:
            t0 = time.perf_counter()
            blob = torch.from_numpy(np.ascontiguousarray(codec.ExportRangeState(False)).copy())
            note("export_ms", t0)
            t0 = time.perf_counter()
            dist.send(blob, dst=rank + 1)
            note("send_ms", t0)

    keys = np.array(keyflags[mine.first:mine.first + mine.count], dtype=np.uint8)
    keys[0] = 1
    if pipelined and hasattr(codec, "set_mvs_hooks"):
        codec.set_mvs_hooks(recv_blob, send_blob)
        try:
            stream, sizes, ftypes = codec.CompressClip(frames, keys, device_ptr=device_ptr, n=mine.count)
        finally:
            codec.set_mvs_hooks(None, None)
    else:
        recv_blob()
        stream, sizes, ftypes = codec.CompressClip(frames, keys, device_ptr=device_ptr, n=mine.count)
        send_blob()
    note("call_ms", t_call)
    return mine, stream, sizes, ftypes


def gather_streams(parts: list[tuple[FrameRange, np.ndarray, np.ndarray, np.ndarray]]):
    """Host-side concatenation of the ranks' outputs in frame order -> (stream, sizes, ftypes) of the whole clip."""
    parts = sorted((p for p in parts if p[0] is not None), key=lambda p: p[0].first)
    return (np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts]),
            np.concatenate([p[3] for p in parts]))
