"""Regenerate tests/golden/ref_threads_digests.json FROM THE UNMODIFIED REFERENCE running with several worker threads.

    python tests/golden/make_threads_golden.py        (build container: needs /root/reference to build oracle/_ref)

Intra-only clips (tests/_clips.band_clip) encoded by oracle/_ref/libscpr_ref.so created with dwNumberOfProcessors = n: the
I-frame bytes of the multi-threaded reference depend on n (one row band per worker, every band starts a new run, SURVEY.md
0.1 / 8(f)5) but not on timing, so they can be pinned.  Every thread count runs in a process of its own: the reference reads
the processor count when a codec codes its first frame.  Per case and thread count: per frame [type, size, md5].
"""
import hashlib
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [(320, 192, 32, 6, 1220), (161, 90, 24, 6, 1061), (640, 360, 32, 4, 1540)]   # w, h, bpp, frames, seed
THREADS = [1, 2, 3, 5, 8]


def child(w, h, bpp, n, seed, threads):
    import numpy as np
    from _clips import band_clip
    from oracle.pyref import RefCodec

    clip = band_clip(w, h, n, seed, bpp)
    ref = RefCodec(w, h, bpp, threads=threads)
    rows = []
    for i in range(n):
        data, ft = ref.compress(np.ascontiguousarray(clip[i]).reshape(-1).copy(), False)
        rows.append([ft, len(data), hashlib.md5(data).hexdigest()])
    print(json.dumps(rows))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(*[int(x) for x in sys.argv[1:]])
        sys.exit(0)
    from oracle.pyref import build

    build()
    out = {}
    for (w, h, bpp, n, seed) in CASES:
        key = f"band_{w}x{h}_rgb{bpp}"
        out[key] = {"args": [w, h, bpp, n, seed], "threads": {}}
        for t in THREADS:
            if t > (h + 15) // 16:
                continue
            r = subprocess.run([sys.executable, os.path.abspath(__file__)] + [str(x) for x in (w, h, bpp, n, seed, t)], capture_output=True, text=True, check=True)
            out[key]["threads"][str(t)] = json.loads(r.stdout.strip().splitlines()[-1])
        assert out[key]["threads"]["1"] != out[key]["threads"]["3"], "the clip does not exercise the band breaks"
    with open(os.path.join(HERE, "ref_threads_digests.json"), "w") as f:
        json.dump(out, f, indent=0)
    print("wrote ref_threads_digests.json:", {k: sorted(v["threads"]) for k, v in out.items()})
