"""The reference's own TIMING mode (screencap.h:47, output of screencap.cpp:399, 1261-1268, 1685-1689) on one I frame and the P
frames after it: the stage split BASELINE.md 3.3 asks for.  TEST / BENCH INFRASTRUCTURE: runs oracle/_ref/libscpr_ref_timing.so
(the unmodified reference compiled with -DTIMING by `make -C oracle ref_timing`), one thread.

    python tools/ref_timing_split.py [config] [frames]      -> one JSON object on stdout

Called as a subprocess by bench.py (the reference prints its timings with printf; the child's stdout is parsed here)."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "oracle", "_ref", "libscpr_ref_timing.so")


def child(name: str, n: int) -> None:
    sys.path.insert(0, ROOT)
    import numpy as np
    from screenpressor_b200 import synth

    cfg = synth.CONFIGS[name]
    clip = synth.make_clip(cfg, n)
    lib = C.CDLL(SO)
    libc = C.CDLL(None)
    lib.ref_create.restype = C.c_void_p
    lib.ref_create.argtypes = [C.c_int] * 5
    lib.ref_compress.restype = C.c_int
    lib.ref_compress.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int]
    lib.ref_decompress.restype = C.c_int
    lib.ref_decompress.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
    h = lib.ref_create(cfg.width, cfg.height, 32, 0, 1)
    cap = cfg.width * cfg.height * 6 + 64
    dst = np.empty(cap, np.uint8)
    for i in range(n):
        fr = np.ascontiguousarray(clip[i]).reshape(-1).copy()
        ft = C.c_int(0 if i == 0 else 1)
        libc.printf(b"\n@frame %d ", i)
        sz = lib.ref_compress(h, fr.ctypes.data, dst.ctypes.data, cap, C.byref(ft), 0)
        libc.printf(b" @type %d @size %d\n", ft.value, sz)
    libc.fflush(None)


def parse(text: str):
    frames = []
    for line in text.split("\n@frame ")[1:]:
        vals = {k: float(v) for k, v in re.findall(r"(\w+)=([0-9.eE+-]+)", line)}
        m = re.search(r"@type (\d+) @size (\d+)", line)
        if not m:
            continue
        frames.append({"type": "I" if m.group(1) == "0" else "P", "bytes": int(m.group(2)), "seconds": vals})
    return frames


def collect(name: str = "cfg2_1080p_rgb32", n: int = 40):
    """-> dict for bench.py: stage times in milliseconds, the I frame and the median over the P frames that coded something"""
    if not os.path.exists(SO):
        return {"unavailable": "oracle/_ref/libscpr_ref_timing.so is not built (make -C oracle ref_timing)"}
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", name, str(n)], capture_output=True, text=True, timeout=600)
    if r.returncode != 0:
        return {"unavailable": "TIMING run failed: " + r.stderr[-200:]}
    frames = parse(r.stdout)
    out = {"config": name, "frames": len(frames), "threads": 1, "unit": "ms",
           "source": "the reference's TIMING printf lines (screencap.cpp:399, 1261-1268, 1685-1689), oracle/_ref/libscpr_ref_timing.so"}
    names = {"rgb24": "rgb32_to_24", "clsfy": "classify_pixels", "encode": "model_and_rans_serialise", "cmp_prev": "cmp_prev",
             "decideblocks": "decide_blocks_and_motion_search", "bts": "block_types", "blocks": "blocks_model_and_rans", "memcpy_prev": "memcpy_prev",
             "CF": "codec_total"}
    for kind in ("I", "P"):
        sel = [f for f in frames if f["type"] == kind and f["bytes"] > 1]
        if not sel:
            continue
        row = {"frames": len(sel)}
        for k, label in names.items():
            xs = sorted(f["seconds"][k] for f in sel if k in f["seconds"])
            if xs:
                row[label] = round(1e3 * xs[len(xs) // 2], 4)
        out[kind + "_frame"] = row
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]))
    else:
        print(json.dumps(collect(sys.argv[1] if len(sys.argv) > 1 else "cfg2_1080p_rgb32", int(sys.argv[2]) if len(sys.argv) > 2 else 40)))
