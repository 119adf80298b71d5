"""Developer tool (GPU): the two frame-scan kernels side by side on one box -- plain 128-bit loads (k_frame_scan32, mode 1) against
the TMA tile stream (k_frame_scan_tma, mode 2) -- at several batch sizes, with the outputs compared.  One JSON line per row."""
import json, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from screenpressor_b200 import synth
from screenpressor_b200.codec import CodecParameters, ScreenCodec
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_1080p_rgb32"
n_max = int(sys.argv[2]) if len(sys.argv) > 2 else 600
cfg = synth.CONFIGS[name]
clip = synth.make_clip(cfg, n_max)
d = torch.from_numpy(clip.reshape(-1)).cuda()
prev = torch.from_numpy(np.ascontiguousarray(clip[0]).reshape(-1)).cuda()
sc = ScreenCodec(0); sc.Init(CodecParameters(cfg.width, cfg.height, 32))
fb = cfg.width * cfg.height * 4
for n in [n_max, n_max // 8, 8, 1]:
    row = {"clip": name, "frames": n, "algorithmic_bytes": 2 * fb * n}
    outs = {}
    for mode, tag in ((1, "ld"), (2, "tma")):
        reps = 10 if n > 8 else 50
        ms, bi, sm = sc.debug_frame_scan(mode, d.data_ptr(), prev.data_ptr(), n, reps)
        ms2, _, _ = sc.debug_frame_scan(mode, d.data_ptr(), prev.data_ptr(), n, reps, fetch=False)
        ms = min(ms, ms2)
        outs[tag] = (bi, sm[:, :3])
        row[tag + "_ms"] = round(ms, 4)
        row[tag + "_GBps_algorithmic"] = round(2 * fb * n / ms / 1e6, 1)
    row["identical"] = bool(np.array_equal(outs["ld"][0], outs["tma"][0]) and np.array_equal(outs["ld"][1], outs["tma"][1]))
    print(json.dumps(row))
