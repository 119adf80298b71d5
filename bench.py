#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: 1080p RGB32 encode+decode frames/s (bit-exact).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference codec core on the host CPU

N = 1 (config.workload): BASELINE.json configs[1], 1920x1080 RGB32 synthetic desktop capture, 600 frames, keyframes at 0 and
500 (the VfW default interval).  One *step* = encoding all 600 frames and decoding them again.  `value` = frames / second with
the input frames already resident in HBM when the timed region starts (bitstreams still travel to the host and back, they are
tiny); `e2e` = the same through the host-buffer API (pinned host frames in, pinned host frames out), which is what a drop-in
user of the reference's ScreenCodec sees.  Inputs (5 GB per step) exceed the 126 MB L2, so no explicit L2 flush is needed.
Further legs of the N = 1 line (each can be switched off with --skip):
  parity          every frame of the CUDA stream byte-compared with the reference's (oracle/_ref, one thread), in `cpu_baseline`
  roofline        the HBM-bound frame scan kernel, timed alone, against MEASURED_PEAKS.json
  gops_in_flight  decode throughput against the number of independent GOPs in one scpr_decompress_clips call, beside the
                  reference decoding the same clips on as many host threads
  frame_api       the drop-in call pattern: one CompressFrame / DecompressFrame per frame with host buffers
  all_configs     BASELINE configs 1, 3, 4, 5: fps, GOP count, byte parity against the committed reference digests

N > 1: the north-star split -- ONE clip cut by contiguous GOP-aligned frame ranges across the N ranks; every
rank encodes and decodes its range on its own GPU, the persistent motion-vector array travels rank to rank around the in-order
resolve (a 64 KB point-to-point message, no collective on the data path), rank 0 concatenates the bitstreams on the host and
checks them byte for byte against its own single-GPU encode of the whole clip.  Total work is fixed: "scaling": "strong".
"""
from __future__ import annotations

import argparse
import ctypes as C
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

from screenpressor_b200 import synth  # noqa: E402

WORKLOAD = "cfg2_1080p_rgb32"
METRIC = "1080p_rgb32_encode_decode_frames_per_s"
SPLIT_GOPS = 8


def workload_config(frames: int):
    """`config` of the JSON line -- the same dict in both arms (the driver compares them)."""
    cfg = synth.CONFIGS[WORKLOAD]
    keys = synth.keyframe_flags(frames, cfg.key_interval)
    return {"workload": f"{WORKLOAD}: {cfg.width}x{cfg.height} RGB32 synthetic desktop capture, {frames} frames per clip, keyframe interval "
                        f"{cfg.key_interval} ({int(keys.sum())} GOPs), one step = encode all frames + decode them again",
            "frames": frames, "key_interval": cfg.key_interval, "gops_per_clip": int(keys.sum()),
            "l2": "inputs (5 GB/step) exceed L2; no flush needed", "statistic": "mean over the timed steps"}


def split_config(frames: int, interval: int, world: int):
    cfg = synth.CONFIGS[WORKLOAD]
    return {"workload": f"{WORKLOAD} content: {cfg.width}x{cfg.height} RGB32 synthetic desktop capture, ONE clip of {frames} frames, keyframe "
                        f"interval {interval} ({SPLIT_GOPS} GOPs), cut by GOP-aligned frame ranges across the ranks; one step = every rank "
                        "encodes + decodes its range, bitstreams concatenated on the host",
            "frames": frames, "key_interval": interval, "gops_per_clip": SPLIT_GOPS,
            "l2": "inputs (>= 1 GB per rank and step) exceed L2; no flush needed", "statistic": "mean over the timed steps"}


def make_workload(rank: int, frames: int):
    cfg = synth.CONFIGS[WORKLOAD]
    clip = synth.make_clip(cfg, frames)
    keys = synth.keyframe_flags(frames, cfg.key_interval)
    return cfg, clip, keys


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------------
# the reference on the host CPU (the only place bench.py executes oracle/)
# ---------------------------------------------------------------------------------------------------------------------------
def time_reference(clip, keys, cfg, threads: int, keep=None, per_frame=None):
    """encode+decode `clip` with the compiled reference core (oracle/_ref); returns (seconds enc, seconds dec, kind).
    keep: a list that receives every frame's (bytes, ftype) -- the parity check of the CUDA stream.
    per_frame: a list that receives (encode seconds, decode seconds, ftype) per frame."""
    from oracle import pyref  # the one place bench.py may execute oracle/: the CPU baseline

    kind = "reference" if pyref.have_ref() else "port"
    Codec = pyref.RefCodec if kind == "reference" else pyref.OracleCodec
    if kind == "port":
        pyref.build()
    enc, dec = Codec(cfg.width, cfg.height, cfg.bpp, threads=threads), Codec(cfg.width, cfg.height, cfg.bpp, threads=1)
    n = len(clip)
    flat = clip.reshape(n, -1)
    te = td = 0.0
    for i in range(n):
        fr = flat[i].copy()
        t0 = time.perf_counter()
        data, ft = enc.compress(fr, not keys[i])
        t1 = time.perf_counter()
        out = dec.decompress(data, ft)
        t2 = time.perf_counter()
        te += t1 - t0
        td += t2 - t1
        if keep is not None:
            keep.append((data, ft))
        if per_frame is not None:
            per_frame.append((t1 - t0, t2 - t1, ft))
        if i % 97 == 0:
            assert np.array_equal(out, flat[i]), "reference round trip failed"
    enc.close()
    dec.close()
    return te, td, kind


def reference_decode_many(cfg, stream, sizes, ftypes, n_clips: int, n_threads: int):
    """n_clips copies of one clip decoded by the reference on n_threads host threads (oracle/ref_capi.cpp ref_decode_many);
    returns wall seconds, or None when only the plain-C port is available."""
    from oracle import pyref

    if not pyref.have_ref():
        return None
    lib = C.CDLL(pyref.REF_SO)
    if not hasattr(lib, "ref_decode_many"):
        return None
    lib.ref_decode_many.restype = C.c_double
    lib.ref_decode_many.argtypes = [C.c_int] * 3 + [C.c_void_p] * 3 + [C.c_int] * 3
    stream = np.ascontiguousarray(stream, np.uint8)
    sizes = np.ascontiguousarray(sizes, np.uint32)
    ftypes = np.ascontiguousarray(ftypes, np.uint8)
    return float(lib.ref_decode_many(cfg.width, cfg.height, cfg.bpp, stream.ctypes.data, sizes.ctypes.data, ftypes.ctypes.data,
                                     int(sizes.size), n_clips, n_threads))


def run_reference(args):
    """The reference's own CPU implementation of the path (oracle/_ref: the unmodified reference core; the plain-C port only
    if it could not be built) on the same clip, the same step and the same statistic as the CUDA arm: every step encodes
    and decodes all frames, `value` is the mean over the K timed steps after W warm-up steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = synth.CONFIGS[WORKLOAD]
    if world > 1 or args.gpus > 1:
        # the N > 1 workload of the CUDA arm: the one long clip (all of it: the CPU has no ranks to cut it across)
        frames, interval = args.split_frames, args.split_frames // SPLIT_GOPS
        clip = synth.make_clip(cfg, frames)
        keys = synth.keyframe_flags(frames, interval)
        config = split_config(frames, interval, max(world, args.gpus))
    else:
        frames = args.frames
        cfg, clip, keys = make_workload(0, frames)
        config = workload_config(frames)
    nby = (cfg.height + 15) // 16
    nthr = max(1, min(os.cpu_count() or 1, nby))  # cap: the reference's tls[] overflows past nby threads
    # all the host threads it can use: one probe step per thread count, the timed steps run with the faster one
    probe = {}
    for thr in sorted({1, nthr}):
        t = time_reference(clip, keys, cfg, thr)
        probe[thr] = frames / (t[0] + t[1])
    thr = max(probe, key=lambda k: probe[k])
    for _ in range(max(0, args.warmup - 1)):
        time_reference(clip, keys, cfg, thr)
    ts = [time_reference(clip, keys, cfg, thr) for _ in range(args.steps)]
    te, td = sum(t[0] for t in ts), sum(t[1] for t in ts)
    fps = frames * args.steps / (te + td)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": (te + td) / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if (world > 1 or args.gpus > 1) else "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": config,
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": thr, "kind": ts[0][2],
                         "sample": f"all {frames} frames of the clip per step, one CompressFrame + one DecompressFrame call per frame, "
                                   f"mean of {args.steps} steps",
                         "encode_fps": frames * args.steps / te, "decode_fps": frames * args.steps / td,
                         "probe_fps_by_threads": {str(k): v for k, v in probe.items()}},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------------
# N = 1 legs
# ---------------------------------------------------------------------------------------------------------------------------
def leg_gops_in_flight(torch, ScreenCodec, CodecParameters, cfg, d_in, keys, local, gop_frames, cpu: bool):
    """Decode throughput against the number of independent GOPs in one call.  The GOP is the first `gop_frames` frames of the
    clip (its I frame and what follows); G copies of it are G independent clips for scpr_decompress_clips_dev: G chains, one
    thread block each, in one launch.  Beside it the reference decoding the same G clips on min(G, cores) host threads."""
    W, H = cfg.width, cfg.height
    fb = W * H * 4
    L = gop_frames
    enc, dec = ScreenCodec(local), ScreenCodec(local)
    for c in (enc, dec):
        c.Init(CodecParameters(W, H, 32))
    s, sizes, fts = enc.CompressClip(None, keys[:L], device_ptr=d_in.data_ptr(), n=L)
    s, sizes, fts = s.copy(), sizes.copy(), fts.copy()
    free, _ = torch.cuda.mem_get_info()
    cores = os.cpu_count() or 1
    rows = []
    for G in (1, 2, 8, 32, 148):
        if G * L * fb > free - (8 << 30):
            rows.append({"gops": G, "skipped": "not enough free HBM for the decoded frames"})
            continue
        d_out = torch.empty(G * L * fb, dtype=torch.uint8, device="cuda")
        clips = [(s, sizes, fts)] * G
        dec.DecompressClips(clips, device_ptr=d_out.data_ptr())  # warm-up (workspaces)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 2
        e0.record()
        for _ in range(reps):
            res, _ = dec.DecompressClips(clips, device_ptr=d_out.data_ptr())
        e1.record()
        e1.synchronize()
        ms = e0.elapsed_time(e1) / reps
        assert all(r == 1 for r in res)
        ref_px = d_in[:L * fb]
        assert torch.equal(d_out[:L * fb], ref_px) and torch.equal(d_out[(G - 1) * L * fb:], ref_px), "multi-clip decode is not bit-exact"
        row = {"gops": G, "frames": G * L, "ms": ms, "decode_fps": G * L / (ms / 1e3)}
        if cpu:
            thr = min(G, cores)
            sec = reference_decode_many(cfg, s, sizes, fts, G, thr)
            if sec:
                row.update({"cpu_ref_decode_fps": G * L / sec, "cpu_threads": thr})
        rows.append(row)
        del d_out
        torch.cuda.empty_cache()
    return {"gop_frames": L, "note": "decode only, frames resident in HBM; one scpr_decompress_clips_dev call per row; cpu_ref = the "
                                     "reference decoding the same clips, one clip per host thread (oracle/_ref ref_decode_many)",
            "host_cores": cores, "curve": rows}


def leg_frame_api(torch, ScreenCodec, CodecParameters, cfg, clip, keys, local, n_api, ref_per_frame):
    """The drop-in call pattern (screenpressor.cpp:425, 620): one scpr_compress_frame / scpr_decompress_frame per frame, pinned
    host buffers both ways, wall clock per call."""
    W, H = cfg.width, cfg.height
    fb = W * H * 4
    n = min(n_api, len(clip))
    enc, dec = ScreenCodec(local), ScreenCodec(local)
    for c in (enc, dec):
        c.Init(CodecParameters(W, H, 32))
    h_src = torch.empty(fb, dtype=torch.uint8, pin_memory=True)
    h_dst = torch.empty(fb, dtype=torch.uint8, pin_memory=True)
    h_bits = torch.empty(W * H * 6, dtype=torch.uint8, pin_memory=True)
    times = []
    for rep in range(2):  # the first pass warms the workspaces up
        enc.Reset()
        dec.Reset()
        times = []
        for i in range(n):
            h_src.numpy()[:] = clip[i].reshape(-1)
            ft = C.c_int(0 if keys[i] else 1)
            t0 = time.perf_counter()
            sz = enc._lib.scpr_compress_frame(enc._h, h_src.data_ptr(), h_bits.data_ptr(), h_bits.numel(), C.byref(ft), 0)
            t1 = time.perf_counter()
            assert sz > 0, sz
            r = dec._lib.scpr_decompress_frame(dec._h, h_bits.data_ptr(), sz, h_dst.data_ptr(), W * 4, ft.value)
            t2 = time.perf_counter()
            assert r == 1, r
            if rep == 1:
                assert torch.equal(h_dst, h_src), f"frame {i}: per-frame decode is not bit-exact"
            times.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, ft.value, int(sz)))
    p = [t for t in times if t[2] == 1]
    i_fr = [t for t in times if t[2] == 0]
    out = {"frames": n, "enc_ms_p50": float(np.median([t[0] for t in p])), "dec_ms_p50": float(np.median([t[1] for t in p])),
           "enc_ms_p90": float(np.percentile([t[0] for t in p], 90)), "dec_ms_p90": float(np.percentile([t[1] for t in p], 90)),
           "enc_ms_max": float(max(t[0] for t in p)), "dec_ms_max": float(max(t[1] for t in p)),
           "I_enc_ms": float(np.mean([t[0] for t in i_fr])), "I_dec_ms": float(np.mean([t[1] for t in i_fr])),
           "note": "P-frame statistics over the P frames of the first frames of the clip (cursor moves, text edits, a window drag, scrolls)"}
    if ref_per_frame:
        rp = [t for t in ref_per_frame[:n] if t[2] == 1]
        ri = [t for t in ref_per_frame[:n] if t[2] == 0]
        out["reference_1_thread"] = {"enc_ms_p50": float(np.median([t[0] for t in rp]) * 1e3), "dec_ms_p50": float(np.median([t[1] for t in rp]) * 1e3),
                                     "enc_ms_p90": float(np.percentile([t[0] for t in rp], 90) * 1e3),
                                     "dec_ms_p90": float(np.percentile([t[1] for t in rp], 90) * 1e3),
                                     "I_enc_ms": float(np.mean([t[0] for t in ri]) * 1e3), "I_dec_ms": float(np.mean([t[1] for t in ri]) * 1e3)}
    return out


def leg_all_configs(torch, ScreenCodec, CodecParameters, local, cpu: bool):
    """BASELINE configs 1, 3, 4, 5 at the lengths tests/golden holds reference digests for: fps with frames resident in HBM,
    GOP count, byte parity of every frame against what the unmodified reference wrote (committed md5s), bit-exact decode."""
    gold = {}
    for name in ("ref_digests_long.json", "ref_digests.json"):
        with open(os.path.join(ROOT, "tests", "golden", name)) as f:
            gold.update(json.load(f))
    cases = [("cfg1_720p_rgb24_f24", "cfg1_720p_rgb24", 24, 500), ("cfg3_2160p_rgb32_f64_k450", "cfg3_2160p_rgb32", 64, 450),
             ("cfg4_1440p_intra_f4", "cfg4_1440p_intra", 4, 1), ("cfg5_5120x1440_f120_k500", "cfg5_5120x1440", 120, 500)]
    rows = []
    for gname, cname, n, interval in cases:
        cfg = synth.CONFIGS[cname]
        clip = synth.make_clip(cfg, n)
        keys = synth.keyframe_flags(n, interval)
        d_in = torch.from_numpy(clip.reshape(-1)).cuda()
        d_out = torch.empty_like(d_in)
        enc, dec = ScreenCodec(local), ScreenCodec(local)
        for c in (enc, dec):
            c.Init(CodecParameters(cfg.width, cfg.height, cfg.bpp))
        enc.reserve_clip_output(64 << 20)
        for rep in range(2):
            enc.Reset()
            dec.Reset()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            s, sizes, fts = enc.CompressClip(None, keys, device_ptr=d_in.data_ptr(), n=n)
            e1.record()
            dec.DecompressClip(s, sizes, fts, device_ptr=d_out.data_ptr())
            e2.record()
            e2.synchronize()
        ok_dec = bool(torch.equal(d_out, d_in))
        pos, same = 0, True
        for i, (gft, gsz, gmd5) in enumerate(gold[gname]["frames"]):
            data = bytes(s[pos:pos + int(sizes[i])])
            pos += int(sizes[i])
            same = same and int(fts[i]) == gft and len(data) == gsz and hashlib.md5(data).hexdigest() == gmd5
        row = {"config": cname, "size": f"{cfg.width}x{cfg.height}x{cfg.bpp}", "frames": n, "gops": int(keys.sum()),
               "encode_fps": n / (e0.elapsed_time(e1) / 1e3), "decode_fps": n / (e1.elapsed_time(e2) / 1e3), "stream_bytes": int(sizes.sum()),
               "identical_to_reference": bool(same), "decode_bit_exact": ok_dec}
        if cpu:
            m = min(n, 24 if cname != "cfg4_1440p_intra" else 2)
            te, td, kind = time_reference(clip[:m], keys[:m], cfg, 1)
            row.update({"ref_encode_fps": m / te, "ref_decode_fps": m / td, "ref_sample": f"first {m} frames, 1 thread, {kind}"})
        rows.append(row)
        assert same and ok_dec, f"{cname}: parity failure"
        del d_in, d_out
        torch.cuda.empty_cache()
    return rows


def run_cuda(args):
    import torch
    import torch.distributed as dist

    from screenpressor_b200.codec import CodecParameters, ScreenCodec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        return run_cuda_split(args, torch, dist, world, rank, local)
    skip = set(args.skip.split(",")) if args.skip else set()
    frames = args.frames
    cfg, clip, keys = make_workload(rank, frames)
    W, H = cfg.width, cfg.height
    fb = W * H * 4
    # pinned host copies for the end-to-end leg, device-resident copy for `value`
    h_in = torch.empty(frames * fb, dtype=torch.uint8, pin_memory=True)
    h_in.numpy()[:] = clip.reshape(-1)
    h_out = torch.empty(frames * fb, dtype=torch.uint8, pin_memory=True)
    d_in = h_in.cuda(non_blocking=True)
    d_out = torch.empty(frames * fb, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.current_stream()
    enc, dec = ScreenCodec(local), ScreenCodec(local)
    for c in (enc, dec):
        c.Init(CodecParameters(W, H, 32))
        c.set_stream(stream.cuda_stream)
    enc.reserve_clip_output(64 << 20)
    stage = {"enc_ms": 0.0, "dec_ms": 0.0, "e2e_enc_s": 0.0, "e2e_dec_s": 0.0}

    def fresh():
        # a new clip: Deinit + Init semantics (prev, models, frame counters), device workspaces are kept
        enc.Reset()
        dec.Reset()

    def step_device():
        fresh()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        s, sizes, fts = enc.CompressClip(None, keys, device_ptr=d_in.data_ptr(), n=frames)
        e1.record(stream)
        dec.DecompressClip(s, sizes, fts, device_ptr=d_out.data_ptr())
        e2.record(stream)
        e2.synchronize()
        stage["enc_ms"] += e0.elapsed_time(e1)
        stage["dec_ms"] += e1.elapsed_time(e2)
        return s, sizes, fts

    def step_host():
        fresh()
        t0 = time.perf_counter()
        s, sizes, fts = enc.CompressClip(h_in.numpy(), keys)
        t1 = time.perf_counter()
        out = dec._lib.scpr_decompress_clip(dec._h, s.ctypes.data, sizes.ctypes.data, fts.ctypes.data, frames, h_out.data_ptr(), W * 4)
        t2 = time.perf_counter()
        assert out == 1, out
        stage["e2e_enc_s"] += t1 - t0  # both calls are synchronous: host clocks bracket them exactly
        stage["e2e_dec_s"] += t2 - t1
        return s, sizes, fts

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            res = fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1), res

    # correctness gate before timing: decode(encode(clip)) must reproduce the clip bit for bit
    step_device()
    torch.cuda.synchronize()
    assert torch.equal(d_out, d_in), "decode(encode(x)) != x on the device path"
    for _ in range(max(0, args.warmup - 1)):
        step_device()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = enc.kernel_launches() + dec.kernel_launches()
    stage["enc_ms"] = stage["dec_ms"] = 0.0
    ms, (s, sizes, fts) = timed(step_device, args.steps)
    s_keep, sizes_keep, fts_keep = s.copy(), sizes.copy(), fts.copy()
    launches = enc.kernel_launches() + dec.kernel_launches() - l0
    clocks = sampler.stop()
    value = frames * args.steps / (ms / 1e3)
    enc_fps = frames * args.steps / (stage["enc_ms"] / 1e3)
    dec_fps = frames * args.steps / (stage["dec_ms"] / 1e3)
    dec_share = stage["dec_ms"] / ms

    # end-to-end leg
    step_host()
    torch.cuda.synchronize()
    assert np.array_equal(h_out.numpy(), h_in.numpy()), "decode(encode(x)) != x on the host-buffer path"
    stage["e2e_enc_s"] = stage["e2e_dec_s"] = 0.0
    ms_e2e, (s2, sizes2, fts2) = timed(step_host, args.steps)
    assert np.array_equal(s2, s_keep) and np.array_equal(sizes2, sizes_keep), "host-buffer stream differs from the device-resident one"
    e2e_enc_fps = frames * args.steps / stage["e2e_enc_s"]
    e2e_dec_fps = frames * args.steps / stage["e2e_dec_s"]
    e2e = frames * args.steps / (ms_e2e / 1e3)
    stream_bytes = int(sizes_keep.sum())

    # roofline leg: the frame-scan kernel (stage A pass 1), timed alone with CUDA events on its stream
    fresh()
    scan_ms = enc.bench_frame_scan(d_in.data_ptr(), frames, 5)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # Algorithmic bytes of the frame scan: every pixel of every frame once, 4 B/px.  SURVEY.md 8(d) quotes W*H*(4+4) per frame
    # (current + previous frame, how the reference and the round-1 kernel read them); k_frame_scan_tma carries the previous
    # frame's tile in registers from one frame to the next, so the second read is not part of the algorithm any more -- the
    # 8 B/px figure is reported beside it as `reference_equivalent_GBps` (it exceeds the HBM peak because those bytes never move).
    alg_bytes = frames * W * H * 4
    achieved = alg_bytes / (scan_ms / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r02_frame_scan_tma_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = int((tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["frames_in_launch"] * frames)
    roofline = {"kernel": "k_frame_scan_tma", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": scan_ms,
                "share_of_step": scan_ms / (ms / args.steps),
                "reference_equivalent_GBps": 2 * achieved,
                "note": "TMA tile stream (cp.async.bulk.tensor.3d + mbarrier rings, one persistent CTA per SM); algorithmic bytes = "
                        "4 B/px, every frame once (the previous frame's tile stays in registers); `traffic` = dram read + write bytes of "
                        "the same launch from ncu (profiles/r02_frame_scan_tma_traffic.json). This kernel is a fraction of a percent of "
                        "the step; the step is the decoder's serial GOP chain, see dominant_kernel"}
    if traffic:
        roofline.update({"dram_GBps": traffic / (scan_ms / 1e3) / 1e9, "dram_frac": traffic / (scan_ms / 1e3) / 1e9 / peak})

    cpu_on = not args.no_cpu_baseline
    cpu = parity = None
    ref_per_frame = []
    if cpu_on:
        # the reference on one host thread (its canonical bitstream) over the same frames; every frame it codes is
        # byte-compared with the CUDA stream of the timed steps
        sample = min(args.ref_frames, frames)
        ref_frames = []
        te, td, kind = time_reference(clip[:sample], keys[:sample], cfg, 1, keep=ref_frames, per_frame=ref_per_frame)
        cpu = {"value": sample / (te + td), "unit": "frames/s", "cores": 1, "kind": kind,
               "sample": f"the first {sample} of {frames} frames of the same clip, one CompressFrame + DecompressFrame call per frame, 1 thread "
                         "(canonical bitstream), one pass",
               "encode_fps": sample / te, "decode_fps": sample / td}
        if kind == "reference" and "timing_split" not in args.skip:
            # the reference's own TIMING mode on the first frames of the clip: where one host core spends an I and a P frame
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            try:
                import ref_timing_split
                cpu["timing_split"] = ref_timing_split.collect(WORKLOAD, 40)
            except Exception as ex:  # a baseline detail must not take the bench line down
                cpu["timing_split"] = {"unavailable": repr(ex)[:200]}
        pos, bad = 0, []
        for i, (data, ft) in enumerate(ref_frames):
            got = bytes(s_keep[pos:pos + int(sizes_keep[i])])
            pos += int(sizes_keep[i])
            if got != data or int(fts_keep[i]) != ft:
                bad.append(i)
        parity = {"frames": len(ref_frames), "identical": not bad,
                  "against": "oracle/_ref (unmodified reference, 1 thread)" if kind == "reference" else "oracle port",
                  "first_mismatches": bad[:5], "stream_md5": hashlib.md5(bytes(s_keep)).hexdigest()}
        assert not bad, f"CUDA bitstream differs from the reference at frames {bad[:5]}"

    gops = frame_api = all_cfg = None
    if "gops" not in skip:
        gops = leg_gops_in_flight(torch, ScreenCodec, CodecParameters, cfg, d_in, keys, local, args.gop_frames, cpu_on)
    if "frame_api" not in skip:
        frame_api = leg_frame_api(torch, ScreenCodec, CodecParameters, cfg, clip, keys, local, args.api_frames, ref_per_frame)
    del d_out, h_out
    torch.cuda.empty_cache()
    if "all_configs" not in skip:
        all_cfg = leg_all_configs(torch, ScreenCodec, CodecParameters, local, cpu_on)

    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": workload_config(frames),
        "stream_bytes_per_clip": stream_bytes,
        "parity": parity,
        "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": frames * fb + stream_bytes,
                "d2h_bytes_per_step": frames * fb + stream_bytes, "ms_per_step": ms_e2e / args.steps,
                "encode_fps": e2e_enc_fps, "decode_fps": e2e_dec_fps,
                "note": "this rank's host-buffer calls: pinned frames in -> bitstream out, bitstream in -> pinned frames out"},
        "gpu_launches": launches,
        "encode_fps": enc_fps, "decode_fps": dec_fps,
        "mpix_per_s": value * W * H / 1e6,
        "clocks": clocks,
        "roofline": roofline,
        "dominant_kernel": {"kernel": "k_dec_chain", "share_of_step": dec_share,
                            "bound": "serial dependency chain of one GOP: one chain warp per GOP at IPC 0.2 (dependent-chain latency, "
                                     "profiles/r02_chain_variants.txt) + a reconstruction warp and copy warps; throughput comes from chains in "
                                     "flight, see gops_in_flight",
                            "algorithmic_decode_bytes_per_step": frames * W * H * 8,
                            "achieved_GBps": frames * W * H * 8 / (stage["dec_ms"] / args.steps / 1e3) / 1e9},
        "gops_in_flight": gops,
        "frame_api": frame_api,
        "all_configs": all_cfg,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------------
# N > 1: one clip cut by frame ranges across the ranks (strong scaling)
# ---------------------------------------------------------------------------------------------------------------------------
def run_cuda_split(args, torch, dist, world, rank, local):
    from screenpressor_b200 import shard
    from screenpressor_b200.codec import CodecParameters, ScreenCodec

    # NCCL writes its version banner to stdout when the first communicator comes up: send it to stderr, rank 0's
    # stdout carries exactly one JSON line.  CPU tensors (the 64 KB hand-off blob) go over gloo, the barrier over NCCL.
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
        dist.barrier()
        torch.cuda.synchronize()
        cpu_group = dist.new_group(backend="gloo")  # for waits during which rank 0 drives every GPU itself (no spinning NCCL kernel)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    cfg = synth.CONFIGS[WORKLOAD]
    W, H = cfg.width, cfg.height
    fb = W * H * 4
    frames, interval = args.split_frames, args.split_frames // SPLIT_GOPS
    keys = synth.keyframe_flags(frames, interval)
    ranges = shard.assign_ranges(keys, world)
    mine = next((r for r in ranges if r.rank == rank), None)
    # rank 0 holds the whole clip (it also produces the single-GPU stream everything is compared with), the others their range
    if rank == 0:
        clip = synth.make_clip(cfg, frames)
        mine_np = clip[mine.first:mine.first + mine.count]
    else:
        clip = None
        mine_np = synth.make_clip_range(cfg, mine.first, mine.count) if mine else None
    n_mine = mine.count if mine else 0
    h_in = torch.empty(max(n_mine, 1) * fb, dtype=torch.uint8, pin_memory=True)
    if n_mine:
        h_in.numpy()[:n_mine * fb] = mine_np.reshape(-1)
    h_out = torch.empty(max(n_mine, 1) * fb, dtype=torch.uint8, pin_memory=True)
    d_in = h_in.cuda()
    d_out = torch.empty_like(d_in)
    stream = torch.cuda.current_stream()
    enc, dec = ScreenCodec(local), ScreenCodec(local)
    for c in (enc, dec):
        c.Init(CodecParameters(W, H, 32))
        c.set_stream(stream.cuda_stream)
    enc.reserve_clip_output(64 << 20)
    stage = {"enc": 0.0, "dec": 0.0}
    relay_stats = {}

    def one_step(host: bool):
        """every rank: encode its range (the mvs[] blob arrives before / leaves after its in-order resolve), decode it again"""
        enc.Reset()
        dec.Reset()
        t0 = time.perf_counter()
        if host:
            rng, s, sizes, fts = shard.encode_sharded(enc, h_in.numpy()[:n_mine * fb] if n_mine else None, keys, rank, world, dist)
        else:
            rng, s, sizes, fts = shard.encode_sharded(enc, None, keys, rank, world, dist, device_ptr=d_in.data_ptr() if n_mine else None,
                                                      stats=relay_stats)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if n_mine:
            if host:
                r = dec._lib.scpr_decompress_clip(dec._h, s.ctypes.data, sizes.ctypes.data, fts.ctypes.data, n_mine, h_out.data_ptr(), W * 4)
                assert r == 1, r
            else:
                dec.DecompressClip(s, sizes, fts, device_ptr=d_out.data_ptr())
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        stage["enc"] += t1 - t0
        stage["dec"] += t2 - t1
        return rng, s, sizes, fts

    def timed(host: bool, steps: int):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stage["enc"] = stage["dec"] = 0.0
        e0.record(stream)
        for _ in range(steps):
            res = one_step(host)
        e1.record(stream)
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1), stage["enc"] * 1e3, stage["dec"] * 1e3], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()], res

    def timed_phases(host: bool, steps: int):
        """encode and decode rates with the ranks in step: a barrier in front of each phase, so that a rank does not count the
        time it waits for a neighbour that is still decoding the previous step as encode time (the ranks' GOPs differ, their
        decode times with them).  -> [encode ms, decode ms] per `steps`, max over ranks, device events."""
        tot = torch.zeros(2, device="cuda")
        for _ in range(steps):
            enc.Reset()
            dec.Reset()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            dist.barrier()
            torch.cuda.synchronize()
            ev[0].record(stream)
            if host:
                rng_, s_, sizes_, fts_ = shard.encode_sharded(enc, h_in.numpy()[:n_mine * fb] if n_mine else None, keys, rank, world, dist)
            else:
                rng_, s_, sizes_, fts_ = shard.encode_sharded(enc, None, keys, rank, world, dist, device_ptr=d_in.data_ptr() if n_mine else None,
                                                              stats=relay_stats)
            ev[1].record(stream)
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            ev[2].record(stream)
            if n_mine:
                if host:
                    r = dec._lib.scpr_decompress_clip(dec._h, s_.ctypes.data, sizes_.ctypes.data, fts_.ctypes.data, n_mine, h_out.data_ptr(), W * 4)
                    assert r == 1, r
                else:
                    dec.DecompressClip(s_, sizes_, fts_, device_ptr=d_out.data_ptr())
            ev[3].record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            tot += t
        return [float(x) for x in tot.tolist()]

    # correctness gate + warm-up
    rng, s, sizes, fts = one_step(False)
    if n_mine:
        assert torch.equal(d_out[:n_mine * fb], d_in[:n_mine * fb]), "decode(encode(range)) != range"
    for _ in range(max(0, args.warmup - 1)):
        one_step(False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = enc.kernel_launches() + dec.kernel_launches()
    (ms, _, _), (rng, s, sizes, fts) = timed(False, args.steps)
    launches = enc.kernel_launches() + dec.kernel_launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    (ms_e2e, _, _), (_, s2, sizes2, _) = timed(True, args.steps)
    # the split of a step into its encode and its decode phase, ranks in step
    relay_stats.clear()
    enc_ms, dec_ms = timed_phases(False, args.steps)
    my_relay = {k: round(v / args.steps, 3) for k, v in relay_stats.items() if not k.startswith("_")}
    e2e_enc_ms, e2e_dec_ms = timed_phases(True, args.steps)
    if n_mine:
        assert np.array_equal(h_out.numpy()[:n_mine * fb], h_in.numpy()[:n_mine * fb]), "host-buffer decode(encode(range)) != range"
        assert np.array_equal(s2, s), "host-buffer stream differs from the device-resident one"
    # host-side concatenation on rank 0
    parts = [None] * world
    dist.gather_object((rng, np.array(s), np.array(sizes), np.array(fts), my_relay), parts if rank == 0 else None, dst=0)
    if rank != 0:
        # leave the GPU to rank 0's one-process leg: free this rank's device buffers and wait on the CPU
        del d_in, d_out, enc, dec
        torch.cuda.empty_cache()
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()
        return
    relay = [p[4] for p in parts]
    cs, csz, cft = shard.gather_streams([p[:4] for p in parts])
    # the same clip on this rank's GPU alone: the single-GPU stream (parity) and the single-GPU time (what N ranks are set against)
    del d_in, d_out
    torch.cuda.empty_cache()
    d_all = torch.from_numpy(clip.reshape(-1)).cuda()
    d_all_out = torch.empty_like(d_all)
    whole, wdec = ScreenCodec(local), ScreenCodec(local)
    for c in (whole, wdec):
        c.Init(CodecParameters(W, H, 32))
    whole.reserve_clip_output(256 << 20)
    single = {}
    for rep in range(2):
        whole.Reset()
        wdec.Reset()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        ws, wsz, wft = whole.CompressClip(None, keys, device_ptr=d_all.data_ptr(), n=frames)
        e1.record()
        wdec.DecompressClip(ws, wsz, wft, device_ptr=d_all_out.data_ptr())
        e2.record()
        e2.synchronize()
        single = {"value": frames / (e0.elapsed_time(e2) / 1e3), "encode_fps": frames / (e0.elapsed_time(e1) / 1e3),
                  "decode_fps": frames / (e1.elapsed_time(e2) / 1e3), "unit": "frames/s",
                  "note": "the whole clip on rank 0's GPU alone, same box, one pass after a warm-up pass (all 8 GOP chains decode "
                          "concurrently on one GPU)"}
    assert torch.equal(d_all_out, d_all)
    same = bool(np.array_equal(cs, ws) and np.array_equal(csz, wsz) and np.array_equal(cft, wft))
    assert same, "sharded bitstream differs from the single-GPU bitstream"
    stream_bytes = int(csz.sum())
    value = frames * args.steps / (ms / 1e3)
    # The same split driven from ONE host process in C++ (csrc/multi.cu: a codec object and a host thread per range, mvs[] relayed
    # device to device, host-side concatenation): rank 0 alone over all N GPUs, pinned host frames in, pinned host frames out.
    one_process = None
    if "one_process" not in args.skip:
        try:
            from screenpressor_b200.codec import MultiCodec
            del d_all, d_all_out, whole, wdec, h_in, h_out
            torch.cuda.empty_cache()
            hp_in = torch.empty(frames * fb, dtype=torch.uint8, pin_memory=True)
            hp_in.numpy()[:] = clip.reshape(-1)
            hp_out = torch.empty(frames * fb, dtype=torch.uint8, pin_memory=True)
            mc = MultiCodec(CodecParameters(W, H, 32), list(range(world)))
            best = None
            for rep in range(3):
                t0 = time.perf_counter()
                ms_, msz, mft, firsts = mc.compress_clip(hp_in.data_ptr(), keys)
                t1 = time.perf_counter()
                ms_ = ms_.copy()
                t1b = time.perf_counter()
                mc.decompress_clip(ms_, msz, mft, hp_out.data_ptr())
                t2 = time.perf_counter()
                cur = {"value": frames / ((t1 - t0) + (t2 - t1b)), "encode_fps": frames / (t1 - t0), "decode_fps": frames / (t2 - t1b)}
                if rep and (best is None or cur["value"] > best["value"]):
                    best = cur
            ident = bool(np.array_equal(ms_, ws) and np.array_equal(msz, wsz) and np.array_equal(mft, wft))
            exact = bool(np.array_equal(hp_out.numpy(), hp_in.numpy()))
            assert ident and exact, "the one-process multi-GPU entry disagrees with the single-GPU stream"
            one_process = dict(best, unit="frames/s", devices=list(range(world)), range_first=firsts, identical=ident, decode_bit_exact=exact,
                               note="scpr_multi_compress_clip / scpr_multi_decompress_clip: rank 0 alone drives all N GPUs from C++ host threads "
                                    "(the other ranks wait on the CPU); wall clock around the calls, host buffers both ways, best of 2 after a warm-up")
            mc.close()
        except AssertionError:
            raise
        except Exception as ex:  # a detail leg must not take the scaling line down
            one_process = {"error": repr(ex)[:300]}
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": split_config(frames, interval, world),
        "ranges": [[r.first, r.count] for r in ranges],
        "stream_bytes_per_clip": stream_bytes,
        "parity": {"frames": frames, "identical": same, "against": "this rank's single-GPU encode of the whole clip (pinned to the reference by "
                   "tests/golden and the N = 1 bench line)", "stream_md5": hashlib.md5(cs.tobytes()).hexdigest()},
        "encode_fps": frames * args.steps / (enc_ms / 1e3), "decode_fps": frames * args.steps / (dec_ms / 1e3),
        "mpix_per_s": value * W * H / 1e6,
        "phases": "encode_fps / decode_fps (also inside e2e): extra steps with a barrier in front of each phase so that the ranks are in step, "
                  "max over ranks, device events; value and e2e.value: free-running steps (a rank may decode while its successor still encodes)",
        "single_gpu": single,
        "encode_relay_ms_per_rank": relay,
        "one_process": one_process,
        "e2e": {"value": frames * args.steps / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": frames * fb + stream_bytes,
                "d2h_bytes_per_step": frames * fb + stream_bytes, "ms_per_step": ms_e2e / args.steps,
                "encode_fps": frames * args.steps / (e2e_enc_ms / 1e3), "decode_fps": frames * args.steps / (e2e_dec_ms / 1e3),
                "note": "every rank: its pinned host frames in -> bitstream out, bitstream in -> pinned host frames out; bytes are the sum over ranks"},
        "gpu_launches": launches,
        "clocks": clocks,
        "serialised_by": "decode: the GOP is one dependency chain -- a rank with one GOP takes as long as a GPU with all of them (one chain per "
                         "SM), so frame ranges only pay once a GPU holds more GOPs than SMs, or for the host transfers (e2e); encode: the in-order "
                         "motion-vector resolves of consecutive ranges run one after the other (mvs[] hand-off), everything else overlaps",
        "roofline": None, "cpu_baseline": None,
    }
    print(json.dumps(line), flush=True)
    dist.barrier(group=cpu_group)
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--frames", type=int, default=synth.CONFIGS[WORKLOAD].frames)
    ap.add_argument("--ref-frames", type=int, default=600, help="frames the cpu_baseline / parity leg of the CUDA arm runs through the reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip", default="", help="comma list of legs to leave out: gops,frame_api,all_configs,timing_split (N = 1), one_process (N > 1)")
    ap.add_argument("--gop-frames", type=int, default=50, help="frames per GOP in the gops_in_flight leg")
    ap.add_argument("--api-frames", type=int, default=120, help="frames of the frame_api leg")
    ap.add_argument("--split-frames", type=int, default=1200, help="N > 1: frames of the one clip that is cut across the ranks (8 GOPs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
