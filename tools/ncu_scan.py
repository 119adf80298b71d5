"""Developer tool (GPU): the frame-scan kernel alone over a device-resident cfg2 clip, for the ncu capture behind bench.py's
roofline.traffic (profiles/r02_frame_scan_tma_traffic.json).  argv: frames (default 600), mode (0 = encoder's choice, 1 = plain loads)."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from screenpressor_b200 import synth
from screenpressor_b200.codec import CodecParameters, ScreenCodec
n = int(sys.argv[1]) if len(sys.argv) > 1 else 600
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cfg = synth.CONFIGS["cfg2_1080p_rgb32"]
clip = synth.make_clip(cfg, n)
d = torch.from_numpy(clip.reshape(-1)).cuda()
prev = torch.from_numpy(np.ascontiguousarray(clip[0]).reshape(-1)).cuda()
sc = ScreenCodec(0); sc.Init(CodecParameters(cfg.width, cfg.height, 32))
ms, _, _ = sc.debug_frame_scan(mode, d.data_ptr(), prev.data_ptr(), n, 3, fetch=False)
print("ok", n, mode, round(ms, 4))
