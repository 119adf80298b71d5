#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: 1080p RGB32 encode+decode frames/s (bit-exact).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference codec core on the host CPU

Workload (config.workload): BASELINE.json configs[1], 1920x1080 RGB32 synthetic desktop capture,
600 frames, keyframes at 0 and 500 (the VfW default interval).  One *step* = encoding all 600
frames and decoding them again.  `value` = frames / second with the input frames already resident
in HBM when the timed region starts (bitstreams still travel to the host and back, they are tiny);
`e2e` = the same through the host-buffer API (pinned host frames in, pinned host frames out), which
is what a drop-in user of the reference's ScreenCodec sees.  Inputs (5 GB per step) exceed the
126 MB L2, so no explicit L2 flush is needed between steps.

Multi-GPU: clips are independent, so each rank encodes+decodes its own clip (seed + rank); there is
no data-path collective (weak scaling).  torch.distributed is only used for the timing barrier.
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

from screenpressor_b200 import synth  # noqa: E402

WORKLOAD = "cfg2_1080p_rgb32"
METRIC = "1080p_rgb32_encode_decode_frames_per_s"


def workload_config(frames: int):
    """`config` of the JSON line -- the same dict in both arms (the driver compares them)."""
    cfg = synth.CONFIGS[WORKLOAD]
    keys = synth.keyframe_flags(frames, cfg.key_interval)
    return {"workload": f"{WORKLOAD}: {cfg.width}x{cfg.height} RGB32 synthetic desktop capture, {frames} frames per clip, keyframe interval "
                        f"{cfg.key_interval} ({int(keys.sum())} GOPs), one step = encode all frames + decode them again",
            "frames": frames, "key_interval": cfg.key_interval, "gops_per_clip": int(keys.sum()),
            "l2": "inputs (5 GB/step) exceed L2; no flush needed", "statistic": "mean over the timed steps"}


def make_workload(rank: int, frames: int):
    cfg = synth.CONFIGS[WORKLOAD]
    if rank:
        cfg = synth.ClipConfig(cfg.name, cfg.width, cfg.height, cfg.bpp, cfg.frames, cfg.key_interval, cfg.seed + 100 * rank, cfg.kind)
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


def time_reference(clip, keys, cfg, threads: int, keep=None):
    """encode+decode `clip` with the compiled reference core (oracle/_ref); returns (seconds enc, seconds dec, kind).
    keep: a list that receives every frame's (bytes, ftype) -- the parity check of the CUDA stream."""
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
        if i % 97 == 0:
            assert np.array_equal(out, flat[i]), "reference round trip failed"
    enc.close()
    dec.close()
    return te, td, kind


def run_reference(args):
    """The reference's own CPU implementation of the path (oracle/_ref: the unmodified reference core; the plain-C port only
    if it could not be built) on the same clip, the same step and the same statistic as the CUDA arm: every step encodes
    and decodes all frames, `value` is the mean over the K timed steps after W warm-up steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # other ranks exit 0 without work
    frames = args.frames
    cfg, clip, keys = make_workload(0, frames)
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
        "warmup": args.warmup, "ms_per_step": (te + td) / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": workload_config(frames),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": thr, "kind": ts[0][2],
                         "sample": f"all {frames} frames of the clip per step, one CompressFrame + one DecompressFrame call per frame, "
                                   f"mean of {args.steps} steps",
                         "encode_fps": frames * args.steps / te, "decode_fps": frames * args.steps / td,
                         "probe_fps_by_threads": {str(k): v for k, v in probe.items()}},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


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
        # NCCL writes its version banner to stdout when the first communicator comes up: send it to stderr, rank 0's
        # stdout carries exactly one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    frames = args.frames
    cfg, clip, keys = make_workload(rank, frames)
    W, H = cfg.width, cfg.height
    fb = W * H * 4
    # pinned host copies for the end-to-end leg, device-resident copy for `value`
    h_in = torch.empty(frames * fb, dtype=torch.uint8, pin_memory=True)
    h_in.numpy()[:] = clip.reshape(-1)
    if world > 1:
        clip = None  # N ranks share the host's memory: keep only the pinned copy (the CPU baseline runs at N = 1 only)
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            res = fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, res

    # correctness gate before timing: decode(encode(clip)) must reproduce the clip bit for bit
    step_device()
    torch.cuda.synchronize()
    assert torch.equal(d_out, d_in), "decode(encode(x)) != x on the device path"
    for _ in range(max(0, args.warmup - 1)):
        step_device()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = enc.kernel_launches() + dec.kernel_launches()
    stage["enc_ms"] = stage["dec_ms"] = 0.0
    ms, (s, sizes, fts) = timed(step_device, args.steps)
    s_keep, sizes_keep, fts_keep = s.copy(), sizes.copy(), fts.copy()
    launches = enc.kernel_launches() + dec.kernel_launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    value = world * frames * args.steps / (ms / 1e3)
    enc_fps = frames * args.steps / (stage["enc_ms"] / 1e3)
    dec_fps = frames * args.steps / (stage["dec_ms"] / 1e3)

    # end-to-end leg
    step_host()
    torch.cuda.synchronize()
    assert np.array_equal(h_out.numpy(), h_in.numpy()), "decode(encode(x)) != x on the host-buffer path"
    stage["e2e_enc_s"] = stage["e2e_dec_s"] = 0.0
    ms_e2e, (s2, sizes2, fts2) = timed(step_host, args.steps)
    assert np.array_equal(s2, s_keep) and np.array_equal(sizes2, sizes_keep), "host-buffer stream differs from the device-resident one"
    e2e_enc_fps = frames * args.steps / stage["e2e_enc_s"]
    e2e_dec_fps = frames * args.steps / stage["e2e_dec_s"]
    e2e = world * frames * args.steps / (ms_e2e / 1e3)
    stream_bytes = int(sizes.sum())

    # several independent clips in flight on one GPU (decode parallelism is per GOP chain):
    # `multi` codec pairs, one CUDA stream + host thread each, same clip
    multi = None
    if args.multi > 1 and world == 1:
        import threading as th

        pairs = []
        for k in range(args.multi):
            st_k = torch.cuda.Stream()
            e_k, d_k = ScreenCodec(local), ScreenCodec(local)
            for c in (e_k, d_k):
                c.Init(CodecParameters(W, H, 32))
                c.set_stream(st_k.cuda_stream)
            e_k.reserve_clip_output(64 << 20)
            pairs.append((e_k, d_k, torch.empty(frames * fb, dtype=torch.uint8, device="cuda")))

        def work(e_k, d_k, out_k):
            e_k.Reset(); d_k.Reset()
            s_k, sz_k, ft_k = e_k.CompressClip(None, keys, device_ptr=d_in.data_ptr(), n=frames)
            d_k.DecompressClip(s_k, sz_k, ft_k, device_ptr=out_k.data_ptr())

        def run_all():
            ts = [th.Thread(target=work, args=p) for p in pairs]
            [t.start() for t in ts]
            [t.join() for t in ts]

        run_all()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        run_all()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert all(torch.equal(p[2], d_in) for p in pairs)
        multi = {"clips_in_flight": args.multi, "value": args.multi * frames / dt, "unit": "frames/s",
                 "note": "same metric with several independent 600-frame clips decoded/encoded concurrently on one GPU (wall clock)"}
        del pairs

    # roofline leg: the frame-scan kernel (stage A pass 1), timed alone with CUDA events on its stream
    fresh()
    scan_ms = enc.bench_frame_scan(d_in.data_ptr(), frames, 5)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes = frames * W * H * 8  # cur + prev, 4 B/px each
    achieved = alg_bytes / (scan_ms / 1e3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "frame_scan_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch_600f")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = parity = None
    if world == 1 and not args.no_cpu_baseline:
        # the reference on one host thread (its canonical bitstream) over the same frames; every frame it codes is
        # byte-compared with the CUDA stream of the timed steps
        sample = min(args.ref_frames, frames)
        ref_frames = []
        te, td, kind = time_reference(clip[:sample], keys[:sample], cfg, 1, keep=ref_frames)
        cpu = {"value": sample / (te + td), "unit": "frames/s", "cores": 1, "kind": kind,
               "sample": f"the first {sample} of {frames} frames of the same clip, one CompressFrame + DecompressFrame call per frame, 1 thread "
                         "(canonical bitstream), one pass",
               "encode_fps": sample / te, "decode_fps": sample / td}
        pos, bad = 0, []
        for i, (data, ft) in enumerate(ref_frames):
            got = bytes(s_keep[pos:pos + int(sizes_keep[i])])
            pos += int(sizes_keep[i])
            if got != data or int(fts_keep[i]) != ft:
                bad.append(i)
        parity = {"frames": len(ref_frames), "identical": not bad, "against": "oracle/_ref (unmodified reference, 1 thread)" if kind == "reference" else "oracle port",
                  "first_mismatches": bad[:5], "stream_md5": hashlib.md5(bytes(s_keep)).hexdigest()}
        assert not bad, f"CUDA bitstream differs from the reference at frames {bad[:5]}"
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
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
        "multi_clip": multi,
        "clocks": clocks,
        "roofline": {"kernel": "k_frame_scan32", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": scan_ms,
                     "share_of_step": scan_ms / (ms / args.steps),
                     "note": "the HBM-bound stage (delta / changed-block detection); the step itself is dominated by the "
                             "decoder's serial GOP chain, see dominant_kernel"},
        "dominant_kernel": {"kernel": "k_dec_chain", "share_of_step": (stage["dec_ms"] / args.steps) / (ms / args.steps),
                            "bound": "serial dependency chain of one GOP: one chain warp per GOP (~6 cycles per issued instruction) + helper warps for "
                                     "motion-vector copies; see profiles/ and multi_clip for the throughput with more chains in flight",
                            "algorithmic_decode_bytes_per_step": frames * W * H * 8,
                            "achieved_GBps": frames * W * H * 8 / (stage["dec_ms"] / args.steps / 1e3) / 1e9},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
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
    ap.add_argument("--multi", type=int, default=8, help="also measure N independent clips in flight on one GPU (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
