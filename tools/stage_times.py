"""Developer tool (GPU): per-stage CUDA-event timings of one encode + decode of a clip (SCPR_TIMING=1)."""
import os, sys, time
os.environ["SCPR_TIMING"] = "1"
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from screenpressor_b200 import synth
from screenpressor_b200.codec import CodecParameters, ScreenCodec
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_1080p_rgb32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 600
interval = int(sys.argv[3]) if len(sys.argv) > 3 else synth.CONFIGS[name].key_interval
cfg = synth.CONFIGS[name]
clip = synth.make_clip(cfg, n); keys = synth.keyframe_flags(n, interval)
d_in = torch.from_numpy(clip.reshape(-1)).cuda(); d_out = torch.empty_like(d_in)
for rep in range(2):
    enc, dec = ScreenCodec(0), ScreenCodec(0)
    enc.Init(CodecParameters(cfg.width, cfg.height, 32)); dec.Init(CodecParameters(cfg.width, cfg.height, 32))
    enc.reserve_clip_output(256 << 20)
    torch.cuda.synchronize(); t0 = time.time()
    s, sizes, fts = enc.CompressClip(None, keys, device_ptr=d_in.data_ptr(), n=n)
    torch.cuda.synchronize(); t1 = time.time()
    dec.DecompressClip(s, sizes, fts, device_ptr=d_out.data_ptr())
    torch.cuda.synchronize(); t2 = time.time()
    print(f"rep {rep}: encode {1e3*(t1-t0):.1f} ms ({n/(t1-t0):.0f} fps), decode {1e3*(t2-t1):.1f} ms ({n/(t2-t1):.0f} fps), bytes {int(sizes.sum())}", file=sys.stderr)
    assert torch.equal(d_in, d_out)
