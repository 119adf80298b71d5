"""GPU tool: encode / decode rates of every BASELINE.json config next to the compiled reference on one host core.

    python tools/config_table.py > profiles/r01_config_table.json

GPU side: frames resident in HBM, second of two repetitions, whole clip per call (CUDA events are not needed at these
durations: wall clock around synchronous calls).  CPU side: oracle/_ref, one thread (the canonical bitstream), a bounded
sample of the same clip, encode then decode per frame; the GPU streams are compared byte for byte with the reference's
on that sample."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import pyref
from screenpressor_b200 import synth
from screenpressor_b200.codec import CodecParameters, ScreenCodec

CASES = [("cfg1_720p_rgb24", 60, 500, 60), ("cfg2_1080p_rgb32", 600, 500, 60), ("cfg3_2160p_rgb32", 60, 450, 12),
         ("cfg4_1440p_intra", 12, 1, 3), ("cfg5_5120x1440", 120, 500, 16)]
pyref.build()
rows = []
for name, n, interval, sample in CASES:
    cfg = synth.CONFIGS[name]
    clip = synth.make_clip(cfg, n); keys = synth.keyframe_flags(n, interval)
    flat = clip.reshape(n, -1)
    d_in = torch.from_numpy(flat.reshape(-1)).cuda(); d_out = torch.empty_like(d_in)
    for rep in range(2):
        enc, dec = ScreenCodec(0), ScreenCodec(0)
        for c in (enc, dec): c.Init(CodecParameters(cfg.width, cfg.height, cfg.bpp))
        enc.reserve_clip_output(512 << 20)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s, sizes, fts = enc.CompressClip(None, keys, device_ptr=d_in.data_ptr(), n=n)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        dec.DecompressClip(s, sizes, fts, device_ptr=d_out.data_ptr())
        torch.cuda.synchronize(); t2 = time.perf_counter()
        s = s.copy()
    assert torch.equal(d_in, d_out), name
    ref_e, ref_d = pyref.RefCodec(cfg.width, cfg.height, cfg.bpp, threads=1), pyref.RefCodec(cfg.width, cfg.height, cfg.bpp, threads=1)
    te = td = 0.0; pos = 0; same = True
    for i in range(sample):
        fr = flat[i].copy()
        a = time.perf_counter(); data, ft = ref_e.compress(fr, not keys[i]); b = time.perf_counter()
        out = ref_d.decompress(data, ft); c2 = time.perf_counter()
        te += b - a; td += c2 - b
        same &= data == bytes(s[pos:pos + int(sizes[i])]) and ft == fts[i]; pos += int(sizes[i])
    rows.append({"config": name, "size": f"{cfg.width}x{cfg.height}x{cfg.bpp}", "frames": n, "gops": int(keys.sum()), "stream_bytes": int(sizes.sum()),
                 "gpu_encode_fps": n / (t1 - t0), "gpu_decode_fps": n / (t2 - t1), "ref_1thread_sample_frames": sample,
                 "ref_encode_fps": sample / te, "ref_decode_fps": sample / td, "bitstream_identical_on_sample": bool(same)})
    print(json.dumps(rows[-1]), flush=True)
    del d_in, d_out
