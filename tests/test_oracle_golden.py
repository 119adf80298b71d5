"""CPU: the plain-C oracle must reproduce the reference's golden bitstreams byte for byte and
decode them back to the source frames (fixtures: tests/golden/, made from the unmodified
reference by tests/golden/make_golden.py)."""
import numpy as np
import pytest

import _golden

DIGESTS = _golden.digests()
# the 1440p intra and 4K cases cost seconds each on the scalar oracle; keep them, they are the
# BASELINE.json configs
CASES = sorted(DIGESTS)


@pytest.mark.parametrize("name", CASES)
def test_oracle_encode_matches_reference_golden(name, oracle_built):
    entry = DIGESTS[name]
    clip, keys, w, h, bpp = _golden.load_case(entry)
    enc = oracle_built.OracleCodec(w, h, bpp)
    dec = oracle_built.OracleCodec(w, h, bpp)
    produced = []
    for i in range(len(clip)):
        fr = np.ascontiguousarray(clip[i]).reshape(-1)
        data, ft = enc.compress(fr.copy(), not keys[i])
        produced.append((data, ft))
        out = dec.decompress(data, ft)
        assert np.array_equal(out, fr), f"{name} frame {i}: oracle decode is not bit-exact"
    _golden.check_frames(name, entry["frames"], produced)


def test_oracle_decodes_reference_streams(oracle_built):
    st = _golden.streams()
    names = sorted({k.split("/")[0] for k in st.files})
    assert names
    for name in names:
        entry = DIGESTS[name]
        clip, keys, w, h, bpp = _golden.load_case(entry)
        sizes, types, data = st[name + "/sizes"], st[name + "/types"], st[name + "/data"].tobytes()
        dec = oracle_built.OracleCodec(w, h, bpp)
        pos = 0
        for i, (sz, ft) in enumerate(zip(sizes, types)):
            out = dec.decompress(data[pos : pos + sz], int(ft))
            pos += sz
            assert np.array_equal(out, np.ascontiguousarray(clip[i]).reshape(-1)), (name, i)


def test_stage_hooks_reproduce_frame_bytes(oracle_built):
    """events -> orc_replay_events -> orc_rans_encode == the I-frame's payload (stage separation)."""
    import ctypes as C

    from _clips import fuzz_clip

    clip, keys = fuzz_clip(130, 70, 1, 5, 32, 16)
    enc = oracle_built.OracleCodec(130, 70, 32)
    data, ft = enc.compress(clip[0].reshape(-1).copy(), False)
    lib = enc.lib
    lib.orc_last_events.restype = C.c_size_t
    lib.orc_last_events.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint32))]
    ptr = C.POINTER(C.c_uint32)()
    n = lib.orc_last_events(enc.h_, C.byref(ptr))
    ev = np.ctypeslib.as_array(ptr, shape=(n,)).copy()
    fq = np.zeros(n, dtype=np.uint32)
    lib.orc_replay_events.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
    lib.orc_replay_events(ev.ctypes.data, n, fq.ctypes.data, 32)
    out = np.zeros(n * 2 + 64, dtype=np.uint8)
    lib.orc_rans_encode.restype = C.c_size_t
    lib.orc_rans_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    sz = lib.orc_rans_encode(fq.ctypes.data, n, out.ctypes.data)
    assert data[0] == 0x32 and bytes(out[:sz]) == data[1:]


def test_legacy_generation_fixtures(oracle_built):
    """tests/golden/legacy_streams.npz (v3 and v2 streams written by the reference): the C oracle decodes the v3 ones
    (same ANS coder, f0 = 64), and the compiled reference -- when oracle/_ref is present -- decodes both generations.
    The CUDA decoder is checked against the same fixtures in tests/test_gpu_parity.py."""
    import os
    import sys

    sys.path.insert(0, _golden.GOLDEN_DIR)
    import make_legacy_golden as mk

    st = np.load(os.path.join(_golden.GOLDEN_DIR, "legacy_streams.npz"))
    cases = mk.cases()
    keys = sorted({"/".join(k.split("/")[:2]) for k in st.files if k.startswith(("v2/", "v3/"))})
    assert len(keys) == 10
    for key in keys:
        ver, name = key.split("/")
        w, h, bpp, clip, _ = cases[name]
        data, sizes, types = st[key + "/data"].tobytes(), st[key + "/sizes"], st[key + "/types"]
        assert data[0] == 2 + (int(ver[1]) - 1) * 16
        decs = []
        if ver == "v3":
            decs.append(oracle_built.OracleCodec(w, h, bpp))
        if oracle_built.have_ref():
            decs.append(oracle_built.RefCodec(w, h, bpp))
        for dec in decs:
            pos = 0
            for i in range(len(clip)):
                out = dec.decompress(data[pos:pos + int(sizes[i])], int(types[i]))
                assert np.array_equal(out, np.ascontiguousarray(clip[i]).reshape(-1)), (key, i, type(dec).__name__)
                pos += int(sizes[i])
