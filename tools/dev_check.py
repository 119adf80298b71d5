"""Developer check (GPU): CUDA encoder vs the oracle, with stage-level localisation of mismatches."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from _clips import fuzz_clip
from oracle import pyref
from screenpressor_b200 import synth
from screenpressor_b200.codec import CodecParameters, ScreenCodec

pyref.build()


def oracle_events(orc):
    lib = orc.lib
    lib.orc_last_events.restype = C.c_size_t
    lib.orc_last_events.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint32))]
    lib.orc_last_freqs.restype = C.c_size_t
    lib.orc_last_freqs.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint32))]
    p = C.POINTER(C.c_uint32)()
    n = lib.orc_last_events(orc.h_, C.byref(p))
    ev = np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.uint32)
    n2 = lib.orc_last_freqs(orc.h_, C.byref(p))
    fq = np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.uint32)
    return ev, fq


def check(name, clip, keys, w, h, bpp, mode):
    n = len(clip)
    orc = pyref.OracleCodec(w, h, bpp)
    exp, evs = [], []
    for i in range(n):
        d, ft = orc.compress(np.ascontiguousarray(clip[i]).reshape(-1).copy(), not keys[i])
        exp.append((d, ft)); evs.append(oracle_events(orc))
    sc = ScreenCodec(); sc.Init(CodecParameters(w, h, bpp))
    t0 = time.time()
    got = []
    if mode == "frame":
        for i in range(n):
            got.append(sc.CompressFrame(clip[i], 0 if keys[i] else 1))
    else:
        stream, sizes, ftypes = sc.CompressClip(clip, keys)
        pos = 0
        for i in range(n):
            got.append((bytes(stream[pos:pos + sizes[i]]), int(ftypes[i]))); pos += int(sizes[i])
    dt = time.time() - t0
    bad = [i for i in range(n) if got[i] != exp[i]]
    print(f"{name} [{mode}] {w}x{h}x{bpp} n={n}: {'OK' if not bad else 'MISMATCH at ' + str(bad[:8])}  ({dt*1e3:.1f} ms)")
    if bad and mode == "clip":
        i = bad[0]
        print("  frame", i, "type exp/got", exp[i][1], got[i][1], "len", len(exp[i][0]), len(got[i][0]), "hdr", exp[i][0][:1], got[i][0][:1])
        ev, iv = sc.debug_events(i)
        oev, ofq = evs[i]
        print("  events exp/got", len(oev), len(ev))
        m = min(len(oev), len(ev))
        d = np.nonzero(oev[:m] != ev[:m])[0]
        if len(d):
            k = d[0]
            print("  first event diff at", k, "exp", [(int(x) >> 16, int(x) & 0xFFFF) for x in oev[max(0,k-3):k+4]], "got", [(int(x) >> 16, int(x) & 0xFFFF) for x in ev[max(0,k-3):k+4]])
        elif len(oev) == len(ev):
            # orc_freq is {freq, cum} little endian = (cum<<16)|freq
            d = np.nonzero(ofq != iv)[0]
            if len(d):
                k = d[0]
                print("  first interval diff at", k, "event", (int(ev[k]) >> 16, int(ev[k]) & 0xFFFF), "exp", hex(int(ofq[k])), "got", hex(int(iv[k])), "ndiff", len(d))
            else:
                print("  events and intervals equal -> rANS/assembly differs")
    sc.Deinit()
    # decoder: reference-exact streams in, source frames out
    dec = ScreenCodec(); dec.Init(CodecParameters(w, h, bpp))
    flat = [np.ascontiguousarray(clip[i]).reshape(-1) for i in range(n)]
    t0 = time.time()
    dbad = []
    if mode == "frame":
        for i in range(n):
            out = dec.DecompressFrame(exp[i][0], None, exp[i][1])
            if not np.array_equal(out, flat[i]): dbad.append(i)
    else:
        stream = np.frombuffer(b"".join(e[0] for e in exp), dtype=np.uint8)
        sizes = np.array([len(e[0]) for e in exp], dtype=np.uint32); fts = np.array([e[1] for e in exp], dtype=np.uint8)
        outs = dec.DecompressClip(stream, sizes, fts)
        for i in range(n):
            if not np.array_equal(outs[i], flat[i]): dbad.append(i)
    dt = time.time() - t0
    print(f"   decode [{mode}]: {'OK' if not dbad else 'MISMATCH at ' + str(dbad[:8])}  ({dt*1e3:.1f} ms)")
    if dbad and mode == "clip":
        i = dbad[0]; d = np.nonzero(outs[i] != flat[i])[0]
        pitch = flat[i].size // h
        print("    frame", i, "type", exp[i][1], "hdr", exp[i][0][0], "ndiff", len(d), "first (y,xbyte)", divmod(int(d[0]), pitch), "last", divmod(int(d[-1]), pitch))
    dec.Deinit()
    return not bad and not dbad


ok = True
cases = []
for (w, h, n, seed, bpp, lv) in [(97, 45, 30, 11, 32, 256), (64, 64, 30, 13, 32, 16), (130, 130, 30, 17, 32, 4), (1001, 37, 30, 12, 24, 4),
                                 (33, 17, 30, 16, 24, 256), (640, 360, 20, 15, 32, 256), (1366, 50, 20, 14, 32, 64), (256, 256, 40, 99, 32, 16)]:
    clip, keys = fuzz_clip(w, h, n, seed, bpp, lv)
    cases.append((f"fuzz{seed}", clip, keys, w, h, bpp))
for cname, nf, interval in [("cfg1_720p_rgb24", 24, 500), ("cfg2_1080p_rgb32", 36, 16), ("cfg5_5120x1440", 12, 500), ("cfg3_2160p_rgb32", 4, 450), ("cfg4_1440p_intra", 2, 1)]:
    cfg = synth.CONFIGS[cname]
    cases.append((cname, synth.make_clip(cfg, nf), synth.keyframe_flags(nf, interval), cfg.width, cfg.height, cfg.bpp))
only = sys.argv[1:] 
for (name, clip, keys, w, h, bpp) in cases:
    if only and not any(o in name for o in only):
        continue
    for mode in ("clip", "frame"):
        ok &= check(name, clip, keys, w, h, bpp, mode)
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
