"""Regenerate tests/golden/ref_digests.json and ref_streams.npz FROM THE UNMODIFIED REFERENCE.

Run in the build container (needs /root/reference to build oracle/_ref):
    python tests/golden/make_golden.py
Every case is encoded by oracle/_ref/libscpr_ref.so with dwNumberOfProcessors = 1 (the only
deterministic configuration of the reference).  The JSON holds, per frame,
[frame type, byte length, md5]; the NPZ holds complete bitstreams of a few tiny clips so that the
decoders can be tested without the reference being present.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _clips import fuzz_clip  # noqa: E402
from oracle.pyref import RefCodec, build  # noqa: E402
from screenpressor_b200 import synth  # noqa: E402

# (name, kind, args)  synth: (config, frames, key interval)  fuzz: (w, h, n, seed, bpp, levels)
CASES = [
    ("cfg1_720p_rgb24_f24", "synth", ("cfg1_720p_rgb24", 24, 500)),
    ("cfg2_1080p_rgb32_f36_k16", "synth", ("cfg2_1080p_rgb32", 36, 16)),
    ("cfg3_2160p_rgb32_f4", "synth", ("cfg3_2160p_rgb32", 4, 450)),
    ("cfg4_1440p_intra_f2", "synth", ("cfg4_1440p_intra", 2, 1)),
    ("cfg5_5120x1440_f24", "synth", ("cfg5_5120x1440", 24, 500)),
    ("fuzz_97x45_rgb32", "fuzz", (97, 45, 40, 11, 32, 256)),
    ("fuzz_1001x37_rgb24_l4", "fuzz", (1001, 37, 40, 12, 24, 4)),
    ("fuzz_64x64_rgb32_l16", "fuzz", (64, 64, 40, 13, 32, 16)),
    ("fuzz_1366x50_rgb32_l64", "fuzz", (1366, 50, 30, 14, 32, 64)),
    ("fuzz_640x360_rgb32", "fuzz", (640, 360, 30, 15, 32, 256)),
    ("fuzz_33x17_rgb24", "fuzz", (33, 17, 40, 16, 24, 256)),
    ("fuzz_130x130_rgb32_l4", "fuzz", (130, 130, 40, 17, 32, 4)),
]
STREAM_CASES = {"fuzz_97x45_rgb32", "fuzz_64x64_rgb32_l16", "fuzz_33x17_rgb24", "fuzz_130x130_rgb32_l4"}
# full-length cases of the BASELINE configs (ref_digests_long.json): the headline clip with its 500-frame GOP of model
# adaptation and mvs[] staleness, a 4K scrolling GOP, intra-only photo/noise frames, the sparse multi-monitor clip
LONG_CASES = [
    ("cfg2_1080p_rgb32_f600_k500", "synth", ("cfg2_1080p_rgb32", 600, 500)),
    ("cfg3_2160p_rgb32_f64_k450", "synth", ("cfg3_2160p_rgb32", 64, 450)),
    ("cfg4_1440p_intra_f4", "synth", ("cfg4_1440p_intra", 4, 1)),
    ("cfg5_5120x1440_f120_k500", "synth", ("cfg5_5120x1440", 120, 500)),
]


def load_case(kind, args):
    if kind == "synth":
        name, n, interval = args
        cfg = synth.CONFIGS[name]
        return synth.make_clip(cfg, n), synth.keyframe_flags(n, interval), cfg.width, cfg.height, cfg.bpp
    w, h, n, seed, bpp, levels = args
    clip, keys = fuzz_clip(w, h, n, seed, bpp, levels)
    return clip, keys, w, h, bpp


def encode_cases(cases, keep_streams=()):
    digests, streams = {}, {}
    for name, kind, args in cases:
        clip, keys, w, h, bpp = load_case(kind, args)
        ref = RefCodec(w, h, bpp, threads=1)
        rows, blobs = [], []
        for i in range(len(clip)):
            data, ft = ref.compress(np.ascontiguousarray(clip[i]).reshape(-1).copy(), not keys[i])
            rows.append([ft, len(data), hashlib.md5(data).hexdigest()])
            blobs.append(data)
        digests[name] = {"kind": kind, "args": list(args), "frames": rows}
        if name in keep_streams:
            streams[name + "/sizes"] = np.array([len(b) for b in blobs], dtype=np.int32)
            streams[name + "/types"] = np.array([r[0] for r in rows], dtype=np.uint8)
            streams[name + "/data"] = np.frombuffer(b"".join(blobs), dtype=np.uint8)
        print(name, len(rows), "frames", sum(r[1] for r in rows), "bytes")
    return digests, streams


def main():
    build()
    if "--long-only" not in sys.argv:
        digests, streams = encode_cases(CASES, STREAM_CASES)
        with open(os.path.join(HERE, "ref_digests.json"), "w") as f:
            json.dump(digests, f, indent=0)
        np.savez_compressed(os.path.join(HERE, "ref_streams.npz"), **streams)
    digests, _ = encode_cases(LONG_CASES)
    with open(os.path.join(HERE, "ref_digests_long.json"), "w") as f:
        json.dump(digests, f, separators=(",", ":"))


if __name__ == "__main__":
    main()
