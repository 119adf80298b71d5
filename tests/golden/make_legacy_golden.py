"""Regenerate tests/golden/legacy_streams.npz: v3 and v2 bitstreams FROM THE UNMODIFIED REFERENCE.

    python tests/golden/make_legacy_golden.py        (build container: needs /root/reference for oracle/_ref)

ScreenCodec::CompressFrame always creates a v4 codec (screencap.cpp:1646-1648), but DecompressFrame creates the codec
the stream header names (screencap.cpp:1700-1701).  Feeding the reference object one 4-byte flat frame with header
0x21 (v3) or 0x11 (v2) first therefore makes its CompressFrame emit that generation -- the same code path old
ScreenPressor releases used (CScreenCapt<UseANS> with f0 = 64, CScreenCapt<UseRC> with RangeCoderSub).  One thread.
Decoding is lossless, so the expected output of a decoder is the source clip itself; the NPZ only holds the streams.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from _clips import fuzz_clip, motion_clip, to_rgb555  # noqa: E402
from oracle.pyref import RefCodec, build  # noqa: E402

# name -> (w, h, bpp, clip, keys)
def cases():
    out = {}
    for name, (w, h, n, seed, bpp, lv) in {"fuzz_97x45_rgb32": (97, 45, 30, 11, 32, 256), "fuzz_33x17_rgb24": (33, 17, 30, 16, 24, 256),
                                            "fuzz_130x130_rgb32_l4": (130, 130, 30, 17, 32, 4), "fuzz_200x120_rgb32_l16": (200, 120, 30, 41, 32, 16)}.items():
        clip, keys = fuzz_clip(w, h, n, seed, bpp, lv)
        out[name] = (w, h, bpp, clip, keys)
    clip = motion_clip(320, 192, 24, 5)
    keys = np.zeros(24, np.uint8); keys[[0, 13]] = 1
    out["motion_320x192_rgb32"] = (320, 192, 32, clip, keys)
    return out


def cases16():
    """16 bpp (5-5-5) versions of two clips; odd width included (RGB24 rows then carry padding)"""
    out = {}
    for name, (w, h, n, seed, lv) in {"fuzz_98x45": (98, 45, 24, 11, 256), "fuzz_131x60_l16": (131, 60, 24, 23, 16)}.items():
        clip, keys = fuzz_clip(w, h, n, seed, 32, lv)
        out[name] = (w, h, 16, to_rgb555(clip), keys)
    out["motion_320x192"] = (320, 192, 16, to_rgb555(motion_clip(320, 192, 16, 7)), np.array([1] + [0] * 15, np.uint8))
    return out


def main():
    build()
    streams = {}
    for ver, prime in ((3, 0x21), (2, 0x11)):
        for name, (w, h, bpp, clip, keys) in cases().items():
            enc, dec = RefCodec(w, h, bpp, threads=1), RefCodec(w, h, bpp, threads=1)
            enc.decompress(bytes([prime, 1, 2, 3]), 0)
            blobs, types = [], []
            for i in range(len(clip)):
                data, ft = enc.compress(np.ascontiguousarray(clip[i]).reshape(-1).copy(), not keys[i])
                assert i > 0 or data[0] == 2 + (ver - 1) * 16, hex(data[0])
                assert np.array_equal(dec.decompress(data, ft), np.ascontiguousarray(clip[i]).reshape(-1)), (ver, name, i)
                blobs.append(data); types.append(ft)
            key = f"v{ver}/{name}"
            streams[key + "/sizes"] = np.array([len(b) for b in blobs], dtype=np.int32)
            streams[key + "/types"] = np.array(types, dtype=np.uint8)
            streams[key + "/data"] = np.frombuffer(b"".join(blobs), dtype=np.uint8)
            print(key, len(blobs), "frames", sum(len(b) for b in blobs), "bytes")
    # 16 bpp clients (current v4 streams): the reference splits every word with the channel masks (screencap.cpp:1665-1678)
    for name, (w, h, bpp, clip, keys) in cases16().items():
        enc, dec = RefCodec(w, h, 16, threads=1), RefCodec(w, h, 16, threads=1)
        blobs, types = [], []
        for i in range(len(clip)):
            fr = np.zeros(h * enc.pitch, np.uint8)
            fr[: h * w * 2] = clip[i].reshape(-1).view(np.uint8)
            data, ft = enc.compress(fr, not keys[i])
            out = dec.decompress(data, ft, pitch=w * 2)
            assert np.array_equal(out.view(np.uint16), clip[i].reshape(-1)), (name, i)
            blobs.append(data); types.append(ft)
        key = f"v4rgb16/{name}"
        streams[key + "/sizes"] = np.array([len(b) for b in blobs], dtype=np.int32)
        streams[key + "/types"] = np.array(types, dtype=np.uint8)
        streams[key + "/data"] = np.frombuffer(b"".join(blobs), dtype=np.uint8)
        print(key, len(blobs), "frames", sum(len(b) for b in blobs), "bytes")
    np.savez_compressed(os.path.join(HERE, "legacy_streams.npz"), **streams)


if __name__ == "__main__":
    main()
