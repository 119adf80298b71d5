"""GPU tool: the drop-in path -- one CompressFrame / DecompressFrame call per frame (host buffers), per-call latency."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from screenpressor_b200 import synth
from screenpressor_b200.codec import CodecParameters, ScreenCodec
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_1080p_rgb32"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cfg = synth.CONFIGS[name]
clip = synth.make_clip(cfg, n); keys = synth.keyframe_flags(n, cfg.key_interval)
flat = clip.reshape(n, -1)
pin = torch.empty(flat.shape[1], dtype=torch.uint8, pin_memory=True)
for rep in range(2):
    enc, dec = ScreenCodec(0), ScreenCodec(0)
    for c in (enc, dec): c.Init(CodecParameters(cfg.width, cfg.height, cfg.bpp))
    te, td, data = [], [], []
    for i in range(n):
        pin.numpy()[:] = flat[i]
        t0 = time.perf_counter(); d, ft = enc.CompressFrame(pin.numpy(), 0 if keys[i] else 1); t1 = time.perf_counter()
        out = dec.DecompressFrame(d, None, ft); t2 = time.perf_counter()
        assert np.array_equal(out, flat[i])
        te.append(t1 - t0); td.append(t2 - t1)
    te, td = np.array(te) * 1e3, np.array(td) * 1e3
    print(f"rep {rep} {name}: I frame encode {te[0]:.2f} ms decode {td[0]:.2f} ms | P frames: encode median {np.median(te[1:]):.2f} ms (p90 {np.percentile(te[1:], 90):.2f}), "
          f"decode median {np.median(td[1:]):.2f} ms (p90 {np.percentile(td[1:], 90):.2f}) | {n / (te.sum() / 1e3):.0f} / {n / (td.sum() / 1e3):.0f} fps")
if os.environ.get("SCPR_TIMING"):
    pass
