"""Developer tool (GPU): one encode + decode of a short clip, for ncu captures of single kernels."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from screenpressor_b200 import synth
from screenpressor_b200.codec import CodecParameters, ScreenCodec
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_1080p_rgb32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cfg = synth.CONFIGS[name]
clip = synth.make_clip(cfg, n); keys = synth.keyframe_flags(n, cfg.key_interval)
d_in = torch.from_numpy(clip.reshape(-1)).cuda(); d_out = torch.empty_like(d_in)
enc, dec = ScreenCodec(0), ScreenCodec(0)
enc.Init(CodecParameters(cfg.width, cfg.height, 32)); dec.Init(CodecParameters(cfg.width, cfg.height, 32))
enc.reserve_clip_output(256 << 20)
s, sizes, fts = enc.CompressClip(None, keys, device_ptr=d_in.data_ptr(), n=n)
dec.DecompressClip(s, sizes, fts, device_ptr=d_out.data_ptr())
torch.cuda.synchronize()
assert torch.equal(d_in, d_out)
print("ok", n, int(sizes.sum()))
