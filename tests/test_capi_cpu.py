"""CPU: the C-ABI library builds, loads and exports every symbol include/scpr_c.h declares; the host
mirror refuses to work without a GPU instead of falling back to anything."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "scpr_c.h")
LIB = os.path.join(ROOT, "screenpressor_b200", "libscpr_b200.so")


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "screenpressor_b200", "csrc")], check=True)
    return C.CDLL(LIB)


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(scpr_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_reference_entry_points():
    syms = declared_symbols()
    for name in ("scpr_create", "scpr_destroy", "scpr_compress_frame", "scpr_decompress_frame"):
        assert name in syms


def test_library_exports_every_declared_symbol(lib):
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} is declared in include/scpr_c.h but not exported"


def test_library_does_not_link_the_oracle():
    out = subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout
    assert "oracle" not in out and "scpr_ref" not in out
    nm = subprocess.run(["nm", "-D", LIB], capture_output=True, text=True).stdout
    assert "orc_" not in nm and "ref_compress" not in nm


def test_sm100a_code_is_present():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_no_cpu_fallback_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from screenpressor_b200.codec import CodecParameters, ScprError, ScreenCodec

    sc = ScreenCodec()
    with pytest.raises(ScprError) as e:
        sc.Init(CodecParameters(64, 64, 32))
    assert e.value.code in (-1004, -1000)


def test_max_compressed_size_matches_vfw_contract(lib):
    from screenpressor_b200.codec import _Params

    lib.scpr_max_compressed_size.restype = C.c_size_t
    p = _Params(1920, 1080, 32, 0, 0, 0, 256, 256, 8, 8, 0)
    assert lib.scpr_max_compressed_size(C.byref(p)) == 1920 * 1080 * 6  # screenpressor.cpp:386-388


def test_vfw_policy_functions():
    """quality -> loss table and InferFrameType of the VfW layer (screenpressor.cpp:410-422, 579-589): pure host code"""
    from screenpressor_b200 import codec

    for q, want in [(0, 4), (2000, 4), (2001, 3), (4000, 3), (4001, 2), (6000, 2), (6001, 1), (8000, 1), (8001, 0), (10000, 0), (99999, 0)]:
        assert codec.quality_to_loss(q) == want, q
    assert [codec.infer_frame_type(b, 100) for b in (0, 1, 0x02, 0x11, 0x12, 0x32, 0x31, 0x22)] == [1, 1, 0, 0, 0, -1, -1, -1]
    assert codec.infer_frame_type(1, 4) == 0 and codec.infer_frame_type(1, 5) == 1


def test_avi_container_round_trip(tmp_path):
    """the RIFF layout is checked with an independent parser written here, then read back through the library"""
    import struct

    from screenpressor_b200 import codec

    rng = np.random.default_rng(3)
    chunks = [bytes(rng.integers(0, 256, int(n), dtype=np.uint8)) for n in (4, 1, 777, 1000, 1, 13)]
    keys = [True, False, False, True, False, False]
    path = str(tmp_path / "clip.avi")
    with codec.AviWriter(path, 322, 200, 16, fps=(25, 1), masks=(0xF800, 0x7E0, 0x1F)) as w:
        for c, k in zip(chunks, keys):
            w.write(c, k)
    raw = open(path, "rb").read()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"AVI " and struct.unpack("<I", raw[4:8])[0] == len(raw) - 8

    def walk(pos, end, out, depth=0):
        while pos + 8 <= end:
            cid, sz = raw[pos:pos + 4], struct.unpack("<I", raw[pos + 4:pos + 8])[0]
            if cid == b"LIST":
                out.append((depth, raw[pos + 8:pos + 12], pos, sz))
                walk(pos + 12, pos + 8 + sz, out, depth + 1)
            else:
                out.append((depth, cid, pos, sz))
            pos += 8 + sz + (sz & 1)
        return out

    tree = walk(12, len(raw), [])
    names = [t[1] for t in tree]
    assert names[:5] == [b"hdrl", b"avih", b"strl", b"strh", b"strf"] and b"movi" in names and names[-1] == b"idx1"
    strh = next(t for t in tree if t[1] == b"strh")[2] + 8
    assert raw[strh:strh + 8] == b"vidsSCPR"
    strf = next(t for t in tree if t[1] == b"strf")
    bi = strf[2] + 8
    assert strf[3] == 52 and struct.unpack("<IiiHH4s", raw[bi:bi + 20]) == (52, 322, 200, 1, 16, b"SCPR")
    assert struct.unpack("<III", raw[bi + 40:bi + 52]) == (0xF800, 0x7E0, 0x1F)
    movi = next(t for t in tree if t[1] == b"movi")[2] + 8
    got = [(raw[t[2] + 8:t[2] + 8 + t[3]]) for t in tree if t[1] == b"00dc"]
    assert got == chunks
    idx = next(t for t in tree if t[1] == b"idx1")
    ents = [struct.unpack("<4sIII", raw[idx[2] + 8 + 16 * i:idx[2] + 24 + 16 * i]) for i in range(idx[3] // 16)]
    assert [e[1] == 0x10 for e in ents] == keys and [e[3] for e in ents] == [len(c) for c in chunks]
    assert all(raw[movi + e[2]:movi + e[2] + 4] == b"00dc" for e in ents)
    avih = next(t for t in tree if t[1] == b"avih")[2] + 8
    assert struct.unpack("<I", raw[avih + 16:avih + 20])[0] == 6 and struct.unpack("<I", raw[avih:avih + 4])[0] == 40000
    with codec.AviReader(path) as r:
        assert len(r) == 6 and (r.info.width, r.info.height, r.info.bits_per_pixel, r.info.fourcc) == (322, 200, 16, codec.FOURCC_SCPR)
        assert (r.info.redmask, r.info.greenmask, r.info.bluemask, r.info.fps_num, r.info.fps_den) == (0xF800, 0x7E0, 0x1F, 25, 1)
        assert [r.read(i) for i in range(6)] == list(zip(chunks, keys))
    # a file without an index (truncated by a crash): chunks are found by scanning, frame types left to the decoder
    cut = raw[:idx[2]]
    open(path, "wb").write(cut)
    with codec.AviReader(path) as r:
        assert [r.read(i)[0] for i in range(len(r))] == chunks
    # the same chunks grouped in 'rec ' lists (interleaved files), no index: the scan descends into the lists
    movi_t = next(t for t in tree if t[1] == b"movi")
    body = raw[movi_t[2] + 12:movi_t[2] + 8 + movi_t[3]]
    rec = b"LIST" + struct.pack("<I", 4 + len(body)) + b"rec " + body
    new_movi = b"LIST" + struct.pack("<I", 4 + len(rec)) + b"movi" + rec
    grouped = raw[:movi_t[2]] + new_movi
    grouped = grouped[:4] + struct.pack("<I", len(grouped) - 8) + grouped[8:]
    open(path, "wb").write(grouped)
    with codec.AviReader(path) as r:
        assert [r.read(i)[0] for i in range(len(r))] == chunks
    # hostile sizes: an index chunk that claims 4 GB, index entries that point past the end of the file
    bad = bytearray(raw)
    bad[idx[2] + 4:idx[2] + 8] = struct.pack("<I", 0xFFFFFFF0)
    bad[idx[2] + 8 + 16 + 12:idx[2] + 8 + 16 + 16] = struct.pack("<I", 0x7FFFFFFF)   # second entry: absurd length
    open(path, "wb").write(bytes(bad))
    with codec.AviReader(path) as r:
        got = [r.read(i)[0] for i in range(len(r))]
        assert got == chunks[:1] + chunks[2:]
    with pytest.raises(codec.ScprError):
        codec.AviReader(str(tmp_path / "missing.avi"))


def test_cpp_facade_compiles_and_fails_loudly_without_gpu(lib):
    """include/screencodec_b200.h (the reference's ScreenCodec class over the C ABI) driven by a CodecInst-shaped C++ caller
    (tests/cpp/vfw_caller.cpp, call sites of screenpressor.cpp:381, 425, 620-636): compiles with g++ -Wall -Werror and links the
    library; without a device Init() marks the object crashed and CompressFrame returns 0 -- no CPU path."""
    import torch

    import _cppharness

    exe = _cppharness.build()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([exe, "--nodevice"], capture_output=True, text=True)
    assert r.returncode == 0 and "status -1004 size 0" in r.stdout, (r.stdout, r.stderr)
