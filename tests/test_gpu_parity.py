"""GPU: the CUDA path, called through the C ABI, against (a) the golden bitstreams generated from the
unmodified reference, (b) the plain-C oracle on seeded fuzz clips (bytes, events, intervals, block
tables), (c) size-independent properties at BASELINE.json's full sizes (encode -> decode round trip)."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import _golden
from _clips import band_clip, fuzz_clip, motion_clip

pytestmark = pytest.mark.gpu

DIGESTS = _golden.digests()


@pytest.fixture(scope="module")
def scpr():
    import __graft_entry__ as ge

    ge.build()
    from screenpressor_b200 import codec

    return codec


def _new(scpr, w, h, bpp):
    sc = scpr.ScreenCodec(0)
    sc.Init(scpr.CodecParameters(w, h, bpp))
    return sc


def _split(stream, sizes, ftypes):
    out, pos = [], 0
    for sz, ft in zip(sizes, ftypes):
        out.append((bytes(stream[pos:pos + int(sz)]), int(ft)))
        pos += int(sz)
    return out


@pytest.mark.parametrize("name", sorted(DIGESTS))
def test_encode_matches_reference_golden_clip_api(scpr, name):
    """whole clip in one call; bytes must equal the reference's, frame by frame"""
    entry = DIGESTS[name]
    clip, keys, w, h, bpp = _golden.load_case(entry)
    enc = _new(scpr, w, h, bpp)
    produced = _split(*enc.CompressClip(clip, keys))
    _golden.check_frames(name, entry["frames"], produced)
    dec = _new(scpr, w, h, bpp)
    stream = np.frombuffer(b"".join(p[0] for p in produced), dtype=np.uint8)
    out = dec.DecompressClip(stream, np.array([len(p[0]) for p in produced], np.uint32), np.array([p[1] for p in produced], np.uint8))
    assert np.array_equal(out.reshape(len(clip), -1), clip.reshape(len(clip), -1)), f"{name}: decode is not bit-exact"


@pytest.mark.parametrize("name", [n for n in sorted(DIGESTS) if n.startswith("fuzz") or n.startswith("cfg1")])
def test_encode_matches_reference_golden_frame_api(scpr, name):
    """the drop-in calls: one CompressFrame / DecompressFrame per frame, state carried across calls"""
    entry = DIGESTS[name]
    clip, keys, w, h, bpp = _golden.load_case(entry)
    enc, dec = _new(scpr, w, h, bpp), _new(scpr, w, h, bpp)
    produced = []
    for i in range(len(clip)):
        data, ft = enc.CompressFrame(clip[i], 0 if keys[i] else 1)
        produced.append((data, ft))
        out = dec.DecompressFrame(data, None, ft)
        assert np.array_equal(out, clip[i].reshape(-1)), f"{name} frame {i}: decode is not bit-exact"
    _golden.check_frames(name, entry["frames"], produced)


def test_decodes_reference_streams(scpr):
    st = _golden.streams()
    for name in sorted({k.split("/")[0] for k in st.files}):
        clip, keys, w, h, bpp = _golden.load_case(DIGESTS[name])
        dec = _new(scpr, w, h, bpp)
        out = dec.DecompressClip(st[name + "/data"], st[name + "/sizes"].astype(np.uint32), st[name + "/types"])
        assert np.array_equal(out.reshape(len(clip), -1), clip.reshape(len(clip), -1)), name


def test_mixed_batch_sizes_equal_single_calls(scpr):
    """splitting a clip into arbitrary batches must not change a byte (model / prev / mvs carry-over)"""
    clip, keys = fuzz_clip(200, 120, 40, 21, 32, 16)
    a = _new(scpr, 200, 120, 32)
    whole = _split(*a.CompressClip(clip, keys))
    b = _new(scpr, 200, 120, 32)
    parts = []
    for lo, hi in [(0, 1), (1, 2), (2, 9), (9, 10), (10, 27), (27, 40)]:
        parts += _split(*b.CompressClip(clip[lo:hi], keys[lo:hi]))
    assert parts == whole
    d = _new(scpr, 200, 120, 32)
    outs = []
    for lo, hi in [(0, 3), (3, 4), (4, 30), (30, 40)]:
        p = whole[lo:hi]
        outs.append(d.DecompressClip(np.frombuffer(b"".join(x[0] for x in p), np.uint8), np.array([len(x[0]) for x in p], np.uint32),
                                     np.array([x[1] for x in p], np.uint8)))
    assert np.array_equal(np.concatenate(outs).reshape(clip.shape), clip)


@pytest.mark.parametrize("size", [(97, 45), (1001, 37), (33, 17), (16, 16), (255, 31), (18, 2), (3, 40), (5, 5), (4, 3), (130, 130)])
@pytest.mark.parametrize("bpp", [32, 24])
def test_fuzz_against_oracle(scpr, oracle_built, size, bpp):
    """odd widths (row padding), tiny frames, noise levels that drive every model promotion"""
    w, h = size
    for levels in (256, 16, 4):
        clip, keys = fuzz_clip(w, h, 24, w * 5 + levels, bpp, levels)
        orc = oracle_built.OracleCodec(w, h, bpp)
        want = [orc.compress(np.ascontiguousarray(clip[i]).reshape(-1).copy(), not keys[i]) for i in range(len(clip))]
        enc = _new(scpr, w, h, bpp)
        got = _split(*enc.CompressClip(clip, keys))
        assert got == want, (size, bpp, levels, [i for i in range(len(got)) if got[i] != want[i]][:5])
        dec = _new(scpr, w, h, bpp)
        out = dec.DecompressClip(np.frombuffer(b"".join(x[0] for x in want), np.uint8), np.array([len(x[0]) for x in want], np.uint32),
                                 np.array([x[1] for x in want], np.uint8))
        assert np.array_equal(out.reshape(len(clip), -1), clip.reshape(len(clip), -1)), (size, bpp, levels)


def test_stage_outputs_match_oracle(scpr, oracle_built):
    """per-stage differential: event list (stage A), intervals (stage B), block tables"""
    w, h = 320, 200
    clip, keys = fuzz_clip(w, h, 16, 31, 32, 16)
    orc = oracle_built.OracleCodec(w, h, 32)
    lib = orc.lib
    lib.orc_last_events.restype = C.c_size_t
    lib.orc_last_events.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint32))]
    lib.orc_last_freqs.restype = C.c_size_t
    lib.orc_last_freqs.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_uint32))]
    lib.orc_last_bts.restype = C.POINTER(C.c_uint8)
    lib.orc_last_bts.argtypes = [C.c_void_p]
    lib.orc_last_sxy.restype = C.POINTER(C.c_int)
    lib.orc_last_sxy.argtypes = [C.c_void_p, C.c_int]
    lib.orc_last_mvs.restype = C.POINTER(C.c_int)
    lib.orc_last_mvs.argtypes = [C.c_void_p, C.c_int]
    checked_blocks = checked_mvs = 0
    enc = _new(scpr, w, h, 32)
    nb = ((w + 15) // 16) * ((h + 15) // 16)
    for i in range(len(clip)):
        data, ft = orc.compress(clip[i].reshape(-1).copy(), not keys[i])
        p = C.POINTER(C.c_uint32)()
        n = lib.orc_last_events(orc.h_, C.byref(p))
        oev = np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.uint32)
        lib.orc_last_freqs(orc.h_, C.byref(p))
        ofq = np.ctypeslib.as_array(p, shape=(n,)).copy() if n else np.zeros(0, np.uint32)
        got, gft = enc.CompressFrame(clip[i], 0 if keys[i] else 1)
        ev, iv = enc.debug_events(0)
        assert np.array_equal(ev, oev), f"frame {i}: event list differs"
        assert np.array_equal(iv, ofq), f"frame {i}: intervals differ"
        assert (got, gft) == (data, ft)
        if ft == 1 and data[0] == 1:
            bts, sxy, mv = enc.debug_blocks(0)
            obts = np.ctypeslib.as_array(lib.orc_last_bts(orc.h_), shape=(nb,))
            assert np.array_equal(bts, obts), f"frame {i}: block types differ"
            osxy = np.stack([np.ctypeslib.as_array(lib.orc_last_sxy(orc.h_, k), shape=(nb,)) for k in range(4)], axis=1)
            omv = np.stack([np.ctypeslib.as_array(lib.orc_last_mvs(orc.h_, k), shape=(nb,)) for k in range(2)], axis=1)
            chg, moved = obts != 0, obts >= 3
            checked_blocks += int(chg.sum())
            checked_mvs += int(moved.sum())
            assert np.array_equal(sxy[chg], osxy[chg]), f"frame {i}: changed sub-rects differ"
            assert np.array_equal(mv[moved], omv[moved]), f"frame {i}: motion vectors differ"
    assert checked_blocks > 100 and checked_mvs > 10, (checked_blocks, checked_mvs)


LONG = _golden.digests_long()


@pytest.mark.parametrize("name", sorted(LONG))
def test_full_length_configs_match_reference(scpr, name):
    """BASELINE configs at full resolution and length (the headline clip with its 500-frame GOP, a 4K scrolling GOP,
    intra-only photo / noise frames, the sparse multi-monitor clip): every frame's type, size and md5 must equal what the
    unmodified reference wrote with one worker thread (tests/golden/ref_digests_long.json), through the clip calls with
    the frames resident on the device and again through the host-buffer calls; decode must be bit-exact."""
    import torch

    entry = LONG[name]
    clip, keys, w, h, bpp = _golden.load_case(entry)
    n = len(clip)
    enc = _new(scpr, w, h, bpp)
    d_in = torch.from_numpy(clip.reshape(-1)).cuda()
    stream, sizes, fts = enc.CompressClip(None, keys, device_ptr=d_in.data_ptr(), n=n)
    _golden.check_frames(name, entry["frames"], _split(stream, sizes, fts))
    d_out = torch.empty_like(d_in)
    dec = _new(scpr, w, h, bpp)
    dec.DecompressClip(stream, sizes, fts, device_ptr=d_out.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(d_out, d_in), f"{name}: decode is not bit-exact"
    del d_out
    enc2 = _new(scpr, w, h, bpp)
    stream2, sizes2, fts2 = enc2.CompressClip(clip, keys)
    assert np.array_equal(stream2, stream) and np.array_equal(sizes2, sizes) and np.array_equal(fts2, fts), f"{name}: host-buffer call differs"


@pytest.mark.parametrize("size", [(128, 16), (132, 17), (640, 360), (1000, 50), (1920, 1080), (260, 33)])
def test_tma_frame_scan_equals_plain_load_scan(scpr, size):
    """The TMA tile stream (csrc/frame_scan_tma.cu: cp.async.bulk.tensor tiles, the previous frame's tile kept in registers along
    a run of frames) against the plain-load kernel (csrc/frame_scan.cu) on the same device frames: block words, flat / changed
    flags and pixel 0 of every frame must be identical -- ragged right and bottom tiles, runs of 1..32 frames, flat frames,
    duplicates, single-pixel changes in tile corners."""
    import torch

    w, h = size
    rng = np.random.default_rng(w * 7 + h)
    for n in (1, 2, 5, 37, 70 if w * h < 500000 else 9):
        clip = np.zeros((n, h, w, 4), np.uint8)
        base = rng.integers(0, 4, (h, w, 3), dtype=np.uint8) * 50
        for i in range(n):
            kind = rng.integers(0, 6)
            if kind == 0 and i:
                pass                                         # duplicate of the previous frame
            elif kind == 1:
                base = base.copy(); base[:] = rng.integers(0, 256, 3, dtype=np.uint8)   # flat frame
            elif kind == 2:
                base = base.copy()
                for _ in range(4):                           # single pixels, corners of tiles and of the frame among them
                    y = int(rng.choice([0, h - 1, 15, 16 % h, rng.integers(0, h)])); x = int(rng.choice([0, w - 1, 127 % w, 128 % w, rng.integers(0, w)]))
                    base[y, x] ^= rng.integers(1, 256, 3, dtype=np.uint8)
            elif kind == 3:
                base = rng.integers(0, 3, (h, w, 3), dtype=np.uint8) * 90
            else:
                base = base.copy()
                y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
                base[y:y + int(rng.integers(1, 40)), x:x + int(rng.integers(1, 200))] = rng.integers(0, 256, 3, dtype=np.uint8)
            clip[i, ..., :3] = base
            clip[i, ..., 3] = rng.integers(0, 256)           # alpha is not part of the picture
        prev = np.ascontiguousarray(clip[-1] if n > 1 else clip[0] ^ 1)
        d = torch.from_numpy(clip).cuda()
        dp = torch.from_numpy(prev).cuda()
        sc = _new(scpr, w, h, 32)
        _, bi1, sm1 = sc.debug_frame_scan(1, d.data_ptr(), dp.data_ptr(), n)
        _, bi2, sm2 = sc.debug_frame_scan(2, d.data_ptr(), dp.data_ptr(), n)
        assert np.array_equal(bi1, bi2), (size, n, np.argwhere(bi1 != bi2)[:5])
        assert np.array_equal(sm1[:, :3], sm2[:, :3]), (size, n, sm1[:4], sm2[:4])
        # and against numpy: changed / notflat flags, pixel 0
        px = clip.view(np.uint32).reshape(n, h, w) & 0xFFFFFF
        pv = np.concatenate([prev.view(np.uint32).reshape(1, h, w) & 0xFFFFFF, px[:-1]])
        assert np.array_equal(sm2[:, 1] != 0, (px != pv).reshape(n, -1).any(1))
        assert np.array_equal(sm2[:, 0] != 0, (px != px[:, :1, :1]).reshape(n, -1).any(1))
        assert np.array_equal(sm2[:, 2], px[:, 0, 0])


def test_maximum_frame_size_matches_reference(scpr, oracle_built):
    """4096 x 4096 = 65 536 blocks, the most the format can address (the changed range of a P frame is two 16-bit block indices,
    screencap.cpp:1145-1150): an I frame, a P frame that scrolls the lower half (motion vectors up to the last block) and
    touches the first and the last pixel, a duplicate.  Bytes against the reference, decode bit-exact.  One block more is refused."""
    w = h = 4096
    clip = band_clip(w, h, 1, 4096, 32)
    f0 = clip[0]
    f1 = f0.copy()
    f1[h // 2:, :, :3] = np.roll(f0[h // 2:, :, :3], -16, axis=0)
    f1[0, 0, :3] ^= 0x5A
    f1[h - 1, w - 1, :3] ^= 0xA5
    frames = np.stack([f0, f1, f1])
    keys = np.array([1, 0, 0], np.uint8)
    ref = oracle_built.RefCodec(w, h, 32) if oracle_built.have_ref() else oracle_built.OracleCodec(w, h, 32)
    want = [ref.compress(np.ascontiguousarray(frames[i]).reshape(-1).copy(), i > 0) for i in range(3)]
    enc = _new(scpr, w, h, 32)
    got = _split(*enc.CompressClip(frames, keys))
    assert [g[1] for g in got] == [x[1] for x in want]
    assert got == want, [i for i in range(3) if got[i] != want[i]]
    assert len(want[2][0]) == 1          # the duplicate is a one-byte frame
    dec = _new(scpr, w, h, 32)
    for i in range(3):
        assert np.array_equal(dec.DecompressFrame(want[i][0], None, want[i][1]), frames[i].reshape(-1)), i
    sc = scpr.ScreenCodec(0)
    with pytest.raises(scpr.ScprError):
        sc.Init(scpr.CodecParameters(4096 + 16, 4096, 32))


def test_full_size_round_trip_properties(scpr):
    """BASELINE configs at full resolution, more frames than the oracle could check in seconds:
    decode(encode(x)) == x, duplicate frames cost one byte, flat frames four."""
    from screenpressor_b200 import synth

    for cname, n in [("cfg2_1080p_rgb32", 150), ("cfg5_5120x1440", 60), ("cfg3_2160p_rgb32", 12), ("cfg4_1440p_intra", 4)]:
        cfg = synth.CONFIGS[cname]
        clip = synth.make_clip(cfg, n)
        keys = synth.keyframe_flags(n, min(cfg.key_interval, 64))
        enc, dec = _new(scpr, cfg.width, cfg.height, 32), _new(scpr, cfg.width, cfg.height, 32)
        stream, sizes, fts = enc.CompressClip(clip, keys)
        out = dec.DecompressClip(stream, sizes, fts)
        assert np.array_equal(out.reshape(n, -1), clip.reshape(n, -1)), cname
        for i in range(1, n):
            if np.array_equal(clip[i], clip[i - 1]) and not keys[i]:
                assert sizes[i] == 1
        assert hashlib.md5(stream.tobytes()).hexdigest()  # stream is materialised


def test_flat_and_duplicate_frames(scpr, oracle_built):
    w, h = 64, 48
    frames = np.zeros((8, h, w, 4), np.uint8)
    frames[..., 3] = 255
    frames[0, ..., :3] = (10, 20, 30)          # flat
    frames[1] = frames[0]                      # same flat colour again
    frames[2, ..., :3] = (10, 20, 31)          # another flat colour -> renew
    frames[3] = frames[2]; frames[3, 5, 7, 1] = 99   # first non-flat frame: forced I
    frames[4] = frames[3]                      # duplicate -> 1 byte
    frames[5] = frames[3]; frames[5, 40:44, 10:30, :3] = 200
    frames[6, ..., :3] = (1, 2, 3)             # flat in the middle of a GOP
    frames[7] = frames[5]
    keys = np.array([1, 0, 0, 0, 0, 0, 0, 0], np.uint8)
    orc = oracle_built.OracleCodec(w, h, 32)
    want = [orc.compress(frames[i].reshape(-1).copy(), not keys[i]) for i in range(8)]
    enc = _new(scpr, w, h, 32)
    got = _split(*enc.CompressClip(frames, keys))
    assert got == want
    assert [len(g[0]) for g in got][:3] == [4, 4, 4] and len(got[4][0]) == 1
    dec = _new(scpr, w, h, 32)
    for i in range(8):
        assert np.array_equal(dec.DecompressFrame(got[i][0], None, got[i][1]), frames[i].reshape(-1)), i


@pytest.mark.parametrize("loss", [1, 2, 3, 4])
def test_lossy_modes_match_oracle(scpr, oracle_built, loss):
    """quality -> loss bits (screenpressor.cpp:418-422): masked pixels, masked frame becomes `prev`"""
    for (w, h, bpp, seed) in [(200, 100, 32, 77), (201, 100, 24, 78), (320, 200, 32, 79)]:
        clip, keys = fuzz_clip(w, h, 20, seed + loss, bpp, 16)
        orc = oracle_built.OracleCodec(w, h, bpp, loss)
        want = [orc.compress(np.ascontiguousarray(clip[i]).reshape(-1).copy(), not keys[i]) for i in range(len(clip))]
        enc = scpr.ScreenCodec(0)
        enc.Init(scpr.CodecParameters(w, h, bpp, loss=loss))
        got = _split(*enc.CompressClip(clip, keys))
        assert got == want, (w, h, bpp, loss, [i for i in range(len(got)) if got[i] != want[i]][:5])
        enc2 = scpr.ScreenCodec(0)
        enc2.Init(scpr.CodecParameters(w, h, bpp))
        got2 = [enc2.CompressFrame(clip[i], 0 if keys[i] else 1, loss) for i in range(len(clip))]
        assert got2 == want
        dec = _new(scpr, w, h, bpp)
        dorc = oracle_built.OracleCodec(w, h, bpp, loss)
        for i, (data, ft) in enumerate(want):
            assert np.array_equal(dec.DecompressFrame(data, None, ft), dorc.decompress(data, ft)), (loss, i)


def test_frame_range_sharding_is_byte_identical(scpr):
    """a clip cut at keyframes and encoded by different codec objects equals the single-codec stream once
    the persistent mvs[] array is handed over -- serially, and pipelined through the resolve hooks"""
    from screenpressor_b200 import shard

    w, h, n = 320, 192, 48
    clip = motion_clip(w, h, n, 5)
    keys = np.zeros(n, np.uint8)
    keys[[0, 13, 30]] = 1
    whole = _split(*_new(scpr, w, h, 32).CompressClip(clip, keys))
    ranges = shard.assign_ranges(keys, 3)
    assert [(r.first, r.count) for r in ranges] == [(0, 13), (13, 17), (30, 18)]
    # (a) serial hand-off
    blob, parts = None, []
    for r in ranges:
        c = _new(scpr, w, h, 32)
        if blob is not None:
            c.ImportRangeState(blob)
        parts += _split(*c.CompressClip(clip[r.first:r.first + r.count], keys[r.first:r.first + r.count]))
        blob = c.ExportRangeState(False)
        assert blob.size == 40 + 8 * ((w + 15) // 16) * ((h + 15) // 16) or blob.size > 0
    assert parts == whole
    # (b) through the hooks: the vectors arrive right before the resolve, leave right after it
    box, parts, calls = {"blob": None}, [], []
    for r in ranges:
        c = _new(scpr, w, h, 32)

        def wait(c=c):
            calls.append("wait")
            if box["blob"] is not None:
                c.ImportRangeState(box["blob"])

        def ready(c=c):
            calls.append("ready")
            box["next"] = c.ExportRangeState(False).copy()

        c.set_mvs_hooks(wait, ready)
        parts += _split(*c.CompressClip(clip[r.first:r.first + r.count], keys[r.first:r.first + r.count]))
        box["blob"] = box["next"]
    assert parts == whole and calls == ["wait", "ready"] * 3
    # (c) the hand-off matters on this clip: without it at least one later range differs
    lone = []
    for r in ranges:
        lone += _split(*_new(scpr, w, h, 32).CompressClip(clip[r.first:r.first + r.count], keys[r.first:r.first + r.count]))
    assert lone[:13] == whole[:13]
    if lone == whole:
        pytest.skip("mvs[] did not influence this clip (hand-off still verified above)")


def test_sharding_with_a_flat_frame_at_a_keyframe(scpr):
    """a requested keyframe that is a flat frame (here: repeating the previous flat colour, so not even the models
    are renewed) does not start a GOP in the reference.  The planner keeps it inside the previous range; a codec that is
    nevertheless handed a small blob there refuses the P frame that has no models to continue instead of coding garbage."""
    from screenpressor_b200 import shard

    w, h, n = 160, 96, 20
    clip = motion_clip(w, h, n, 9)
    for f in (9, 10):                      # two flat frames of one colour; frame 10 is a requested keyframe
        clip[f, ..., :3] = (40, 50, 60)
    keys = np.zeros(n, np.uint8)
    keys[[0, 10, 15]] = 1
    whole = _split(*_new(scpr, w, h, 32).CompressClip(clip, keys))
    assert [len(x[0]) for x in whole[9:11]] == [4, 4] and whole[11][1] == 1   # flat, flat, then a P frame on the old models
    skip = shard.flat_keyframes(clip, keys)
    assert skip == [10]
    ranges = shard.assign_ranges(keys, 3, skip)
    assert [(r.first, r.count) for r in ranges] == [(0, 15), (15, 5)]
    blob, parts = None, []
    for r in ranges:
        c = _new(scpr, w, h, 32)
        if blob is not None:
            c.ImportRangeState(blob)
        parts += _split(*c.CompressClip(clip[r.first:r.first + r.count], keys[r.first:r.first + r.count]))
        blob = c.ExportRangeState(False)
    assert parts == whole
    # the bad cut: range [10, 15) on a small blob
    a = _new(scpr, w, h, 32)
    a.CompressClip(clip[:10], keys[:10])
    b = _new(scpr, w, h, 32)
    b.ImportRangeState(a.ExportRangeState(False))
    with pytest.raises(scpr.ScprError):
        b.CompressClip(clip[10:15], keys[10:15])
    # ... and the same cut on a full blob is exact
    b2 = _new(scpr, w, h, 32)
    b2.ImportRangeState(a.ExportRangeState(True))
    assert _split(*b2.CompressClip(clip[10:15], keys[10:15])) == whole[10:15]


def test_failed_call_leaves_a_decodable_stream(scpr):
    """an error after device work has started (here: destination too small) drops the frames of that call; the
    next coded frame is an I frame with fresh models and the stream the caller holds decodes bit-exactly."""
    w, h, n = 160, 96, 12
    clip = motion_clip(w, h, n, 3)
    keys = np.zeros(n, np.uint8)
    keys[0] = 1
    enc = _new(scpr, w, h, 32)
    got = _split(*enc.CompressClip(clip[:4], keys[:4]))
    sizes = np.zeros(4, np.uint32)
    fts = np.zeros(4, np.uint8)
    tiny = np.zeros(8, np.uint8)
    r = enc._lib.scpr_compress_clip(enc._h, clip[4:8].ctypes.data, 4, keys[4:8].ctypes.data, tiny.ctypes.data, tiny.size, sizes.ctypes.data,
                                    fts.ctypes.data)
    assert r == scpr.SCPR_E_DSTSIZE
    rest = _split(*enc.CompressClip(clip[8:], keys[8:]))
    assert rest[0][1] == 0 and rest[0][0][0] == 0x32          # forced I frame
    dec = _new(scpr, w, h, 32)
    frames = got + rest
    out = dec.DecompressClip(np.frombuffer(b"".join(x[0] for x in frames), np.uint8), np.array([len(x[0]) for x in frames], np.uint32),
                             np.array([x[1] for x in frames], np.uint8))
    assert np.array_equal(out.reshape(8, -1), np.concatenate([clip[:4], clip[8:]]).reshape(8, -1))


def test_many_clips_in_one_call(scpr):
    """scpr_decompress_clips: every GOP of every clip is one thread block of one launch; each clip decodes as by a fresh
    codec, a clip that starts with a P frame is refused without touching the others, host and device destinations agree"""
    import torch

    w, h = 160, 96
    specs = [(20, 3, [0, 7]), (9, 4, [0]), (31, 5, [0, 10, 20, 30]), (12, 6, [0, 5])]
    clips, want = [], []
    for n, seed, kf in specs:
        clip = motion_clip(w, h, n, seed)
        keys = np.zeros(n, np.uint8)
        keys[kf] = 1
        stream, sizes, fts = _new(scpr, w, h, 32).CompressClip(clip, keys)
        clips.append((stream.copy(), sizes.copy(), fts.copy()))
        want.append(clip.reshape(n, -1))
    # a fifth "clip" that starts on a P frame: the tail of clip 0
    s0, z0, f0 = clips[0]
    cut = int(z0[:3].sum())
    clips.insert(2, (s0[cut:], z0[3:], f0[3:]))
    want.insert(2, None)
    dec = _new(scpr, w, h, 32)
    results, frames = dec.DecompressClips(clips)
    assert results == [1, 1, 0, 1, 1]
    for k, wnt in enumerate(want):
        if wnt is not None:
            assert np.array_equal(frames[k], wnt), k
    # device destination: clips back to back
    total = sum(int(c[1].size) for c in clips)
    d_out = torch.zeros(total * w * h * 4, dtype=torch.uint8, device="cuda")
    results, _ = dec.DecompressClips(clips, device_ptr=d_out.data_ptr())
    assert results == [1, 1, 0, 1, 1]
    got = d_out.cpu().numpy().reshape(total, -1)
    pos = 0
    for k, wnt in enumerate(want):
        n = int(clips[k][1].size)
        if wnt is not None:
            assert np.array_equal(got[pos:pos + n], wnt), k
        pos += n
    # the codec is usable as a single-stream decoder afterwards (its state was reset, not corrupted)
    out = dec.DecompressClip(*clips[1])
    assert np.array_equal(out, want[1])
    # a bad stream version in one clip
    bad = clips[1][0].copy()
    bad[0] = 0x72  # version 8
    results, frames = dec.DecompressClips([clips[0], (bad, clips[1][1], clips[1][2]), clips[3]])
    assert results == [1, -8, 1] and np.array_equal(frames[0], want[0]) and np.array_equal(frames[2], want[3])


def test_pitch_per_call_and_untouched_row_padding(scpr):
    """the reference takes the output pitch per call and writes width * bytes-per-pixel bytes per row (screencap.cpp:1705-1738):
    the pitch may change between calls of one stream, and the caller's row padding is never written"""
    w, h, n = 150, 64, 10
    clip = motion_clip(w, h, n, 11)
    keys = np.zeros(n, np.uint8)
    keys[0] = 1
    stream, sizes, fts = _new(scpr, w, h, 32).CompressClip(clip, keys)
    frames = _split(stream, sizes, fts)
    dec = _new(scpr, w, h, 32)
    for i, (data, ft) in enumerate(frames):
        pitch = w * 4 + (0, 16, 64)[i % 3]
        out = np.full(h * pitch, 0xA5, np.uint8)
        src = np.frombuffer(data, np.uint8)
        r = dec._lib.scpr_decompress_frame(dec._h, src.ctypes.data, len(data), out.ctypes.data, pitch, ft)
        assert r == 1, (i, r)
        rows = out.reshape(h, pitch)
        assert np.array_equal(rows[:, :w * 4], clip[i].reshape(h, w * 4)), i
        assert (rows[:, w * 4:] == 0xA5).all(), f"frame {i}: row padding was written"
    # clip call with a padded pitch
    dec2 = _new(scpr, w, h, 32)
    pitch = w * 4 + 32
    out = dec2.DecompressClip(stream, sizes, fts, pitch=pitch).reshape(n, h, pitch)
    assert np.array_equal(out[:, :, :w * 4], clip.reshape(n, h, w * 4))


def test_rejected_decode_call_leaves_the_decoder_state_alone(scpr):
    """a call that is refused while it is being planned (here: a truncated flat frame in the middle of a batch) must
    not change what the next valid call does"""
    w, h = 64, 48
    frames = np.zeros((6, h, w, 4), np.uint8)
    frames[..., 3] = 255
    frames[0, ..., :3] = (10, 20, 30)
    frames[0, 3, 4, 0] = 9
    frames[1] = frames[0]; frames[1, 8:12, 8:20, :3] = 77
    frames[2, ..., :3] = (1, 2, 3)     # flat
    frames[3, ..., :3] = (1, 2, 3)     # same flat colour: no renewal
    frames[4] = frames[1]
    frames[5] = frames[4]; frames[5, 20:24, 8:20, :3] = 99
    keys = np.array([1, 0, 0, 0, 0, 0], np.uint8)
    coded = _split(*_new(scpr, w, h, 32).CompressClip(frames, keys))
    dec = _new(scpr, w, h, 32)
    for i in range(3):
        assert np.array_equal(dec.DecompressFrame(coded[i][0], None, coded[i][1]), frames[i].reshape(-1))
    # a batch of a flat frame of ANOTHER colour followed by a truncated flat frame: refused as a whole.  Had the first of them
    # been remembered as "the last flat colour", the next flat frame would renew the models and the P frames after it break.
    bad_stream = np.frombuffer(bytes([0x31, 5, 5, 5]) + bytes([0x31, 9]), np.uint8)
    with pytest.raises(scpr.ScprError):
        dec.DecompressClip(bad_stream, np.array([4, 2], np.uint32), np.array([0, 0], np.uint8))
    for i in range(3, 6):
        assert np.array_equal(dec.DecompressFrame(coded[i][0], None, coded[i][1]), frames[i].reshape(-1)), i


def test_multi_device_entry_is_byte_identical(scpr):
    """scpr_compress_clip_multi / scpr_decompress_clip_multi: the frame-range split driven from one host process, one codec
    object and host thread per range, mvs[] relayed device to device around the resolves (here all ranges on device 0: the
    relay logic is the same, cudaMemcpyPeerAsync within one device).  Byte-identical to one codec; a flat keyframe is not a cut."""
    w, h, n = 320, 192, 60
    clip = motion_clip(w, h, n, 5)
    clip[30, ..., :3] = (9, 8, 7)          # a flat frame exactly where a keyframe is requested
    keys = np.zeros(n, np.uint8)
    keys[[0, 13, 30, 41, 52]] = 1
    params = scpr.CodecParameters(w, h, 32)
    whole = _new(scpr, w, h, 32).CompressClip(clip, keys)
    want = (whole[0].copy(), whole[1].copy(), whole[2].copy())
    for devs in ([0], [0, 0], [0, 0, 0], [0] * 8):
        stream, sizes, fts, firsts = scpr.compress_clip_multi(params, devs, clip, keys)
        assert np.array_equal(stream, want[0]) and np.array_equal(sizes, want[1]) and np.array_equal(fts, want[2]), devs
        assert 30 not in firsts and all(keys[f] for f in firsts) and len(firsts) == min(len(devs), 4), (devs, firsts)
        out = scpr.decompress_clip_multi(params, devs, stream, sizes, fts)
        assert np.array_equal(out.reshape(n, -1), clip.reshape(n, -1)), devs
    # the standing set (scpr_multi_*): several clips through the same objects
    mc = scpr.MultiCodec(params, [0, 0, 0])
    flat = np.ascontiguousarray(clip.reshape(-1))
    out = np.zeros_like(flat)
    for rep in range(3):
        src = flat if rep != 1 else np.ascontiguousarray(clip[::-1].reshape(-1))   # another clip in between
        stream, sizes, fts, firsts = mc.compress_clip(src.ctypes.data, keys)
        if rep != 1:
            assert np.array_equal(stream, want[0]) and np.array_equal(sizes, want[1]) and np.array_equal(fts, want[2]), rep
        mc.decompress_clip(stream.copy(), sizes, fts, out.ctypes.data)
        assert np.array_equal(out, src), rep
    mc.close()


def test_cpp_facade_driven_like_the_vfw_layer(scpr, oracle_built, tmp_path):
    """include/screencodec_b200.h through a compiled C++ caller shaped like CodecInst (tests/cpp/vfw_caller.cpp: CompressBegin /
    Compress with the forced keyframe interval and quality -> loss / Decompress with InferFrameType and the BadVersionException
    handler, screenpressor.cpp:343-437, 579-640) -- no ctypes on this path.  Bytes, key flags and decoded frames must equal what the
    unmodified reference produces under the same driving."""
    import subprocess

    import _cppharness

    exe = _cppharness.build()
    for (w, h, bpp, n, kf, quality) in [(322, 200, 32, 40, 12, 10000), (161, 90, 24, 30, 500, 10000), (256, 128, 32, 20, 7, 5000)]:
        clip, _ = fuzz_clip(w, h, n, 77 + w, bpp=bpp, levels=16)
        raw = tmp_path / "frames.raw"
        clip.tofile(raw)
        outs = [tmp_path / x for x in ("out.stream", "out.index", "out.decoded")]
        r = subprocess.run([exe, str(w), str(h), str(bpp), str(n), str(kf), str(quality), str(raw)] + [str(o) for o in outs], capture_output=True, text=True)
        assert r.returncode == 0, (r.stdout, r.stderr)
        assert "bogus stream: result -2 version 10" in r.stdout, r.stdout
        index = np.fromfile(outs[1], dtype=np.uint32).reshape(n, 2)
        stream = np.fromfile(outs[0], dtype=np.uint8)
        decoded = np.fromfile(outs[2], dtype=np.uint8).reshape(n, -1)
        # the reference under the same policy: npframes + 1 >= interval forces a keyframe, loss from the quality
        loss = min((10000 - quality) // 2000, 4)
        ref = oracle_built.RefCodec(w, h, bpp, loss=loss) if oracle_built.have_ref() else oracle_built.OracleCodec(w, h, bpp, loss=loss)
        refdec = oracle_built.RefCodec(w, h, bpp) if oracle_built.have_ref() else oracle_built.OracleCodec(w, h, bpp)
        npf, pos = 0, 0
        for i in range(n):
            want_p = not (npf + 1 >= kf)
            data, ft = ref.compress(np.ascontiguousarray(clip[i]).reshape(-1).copy(), want_p, loss)
            npf = 0 if ft == 0 else npf + 1
            sz = int(index[i, 0])
            assert bytes(stream[pos:pos + sz]) == data, (w, h, bpp, i)
            assert (int(index[i, 1]) == 0x10) == (ft == 0), (w, h, bpp, i)
            assert np.array_equal(decoded[i], refdec.decompress(data, ft)), (w, h, bpp, i)
            pos += sz
        assert pos == stream.size


@pytest.mark.parametrize("case", [(320, 192, 32), (161, 90, 24), (640, 360, 32)])
def test_threads_layout_matches_multithreaded_reference_i_frames(scpr, oracle_built, case):
    """the reference with n worker threads splits an I frame into n row bands and every band starts a new run
    (squad.cpp:16-31, screencap.cpp:862-866, 876-919, 365-388).  scpr_set_threads_layout(n) must write the same bytes -- checked
    against the unmodified reference created with n threads on intra-only clips (its multi-threaded P frames are timing dependent,
    SURVEY 0.1, so only I frames can be pinned), frame and clip API; any decoder reads them."""
    Checker = oracle_built.RefCodec if oracle_built.have_ref() else oracle_built.OracleCodec   # (the port restates the bands too)
    w, h, bpp = case
    n = 6
    clip = band_clip(w, h, n, 900 + w, bpp)
    keys = np.ones(n, np.uint8)
    nby = (h + 15) // 16
    canonical = None
    for threads in (1, 2, 3, 5, nby):
        ref = Checker(w, h, bpp, threads=threads)   # (the reference reads the thread count at this codec's first CompressFrame)
        want = [ref.compress(np.ascontiguousarray(clip[i]).reshape(-1).copy(), False) for i in range(n)]
        if threads == 1:
            canonical = want
        else:
            assert want != canonical, "the clip does not exercise the band breaks"
        enc = _new(scpr, w, h, bpp)
        enc.set_threads_layout(threads)
        got = _split(*enc.CompressClip(clip, keys))
        flat = [i for i in range(n) if len(want[i][0]) <= 4]
        assert len(flat) < n
        assert got == want, (case, threads, [i for i in range(n) if got[i] != want[i]])
        enc2 = _new(scpr, w, h, bpp)
        enc2.set_threads_layout(threads)
        assert [enc2.CompressFrame(clip[i], 0) for i in range(n)] == want, (case, threads)
        dec = _new(scpr, w, h, bpp)
        for i in range(n):
            assert np.array_equal(dec.DecompressFrame(want[i][0], None, want[i][1]), clip[i].reshape(-1)), (case, threads, i)
    bad = _new(scpr, w, h, bpp)
    with pytest.raises(scpr.ScprError):
        bad.set_threads_layout(nby + 1)
    with pytest.raises(scpr.ScprError):
        bad.set_threads_layout(0)


def test_threads_layout_matches_committed_reference_digests(scpr):
    """the same against the committed fixture (tests/golden/ref_threads_digests.json, written by make_threads_golden.py from the
    unmodified reference with 1, 2, 3, 5, 8 worker threads): does not need oracle/_ref on the box"""
    import json
    import os

    with open(os.path.join(_golden.GOLDEN_DIR, "ref_threads_digests.json")) as f:
        fixture = json.load(f)
    for name, entry in fixture.items():
        w, h, bpp, n, seed = entry["args"]
        clip = band_clip(w, h, n, seed, bpp)
        for threads, rows in entry["threads"].items():
            enc = _new(scpr, w, h, bpp)
            enc.set_threads_layout(int(threads))
            produced = _split(*enc.CompressClip(clip, np.ones(n, np.uint8)))
            _golden.check_frames(f"{name} threads={threads}", rows, produced)


def test_full_state_checkpoint_resume_at_any_frame(scpr):
    """full = 1: previous frame + adaptive models + mvs[]; an encode resumed in another codec object continues byte-exactly"""
    w, h, n = 200, 120, 30
    clip, keys = fuzz_clip(w, h, n, 41, 32, 16)
    whole = _split(*_new(scpr, w, h, 32).CompressClip(clip, keys))
    for cut in (1, 7, 19):
        a = _new(scpr, w, h, 32)
        parts = _split(*a.CompressClip(clip[:cut], keys[:cut]))
        blob = a.ExportRangeState(True)
        b = _new(scpr, w, h, 32)
        b.ImportRangeState(blob)
        parts += _split(*b.CompressClip(clip[cut:], keys[cut:]))
        assert parts == whole, cut
    with pytest.raises(scpr.ScprError):
        _new(scpr, w + 16, h, 32).ImportRangeState(blob)


def _legacy_cases():
    import os

    path = os.path.join(_golden.GOLDEN_DIR, "legacy_streams.npz")
    st = np.load(path)
    return st, sorted({"/".join(k.split("/")[:2]) for k in st.files})


@pytest.mark.parametrize("version", [3, 2])
def test_decodes_legacy_stream_generations(scpr, version):
    """old files: v3 (same ANS coder, Cx6 start frequency 64) and v2 (range coder, no MV-repeat flag) streams written by
    the reference itself (tests/golden/make_legacy_golden.py) decode to the source clips, per frame and per clip"""
    import sys
    sys.path.insert(0, _golden.GOLDEN_DIR)
    import make_legacy_golden as mk

    st, keys = _legacy_cases()
    cases = mk.cases()
    seen = 0
    for key in keys:
        ver, name = key.split("/")
        if ver != f"v{version}":
            continue
        w, h, bpp, clip, _ = cases[name]
        data, sizes, types = st[key + "/data"], st[key + "/sizes"].astype(np.uint32), st[key + "/types"]
        assert data[0] == 2 + (version - 1) * 16
        out = _new(scpr, w, h, bpp).DecompressClip(data, sizes, types)
        assert np.array_equal(out.reshape(len(clip), -1), clip.reshape(len(clip), -1)), key
        dec, pos = _new(scpr, w, h, bpp), 0
        for i in range(len(clip)):
            fr = dec.DecompressFrame(bytes(data[pos:pos + int(sizes[i])]), None, int(types[i]))
            assert np.array_equal(fr, np.ascontiguousarray(clip[i]).reshape(-1)), (key, i)
            pos += int(sizes[i])
        seen += 1
    assert seen == 5


def test_rgb16_clients_match_reference(scpr):
    """16 bpp frames (5-5-5 masks): bitstream bytes equal the reference's, decode returns the words, any output pitch"""
    import sys
    sys.path.insert(0, _golden.GOLDEN_DIR)
    import make_legacy_golden as mk

    st, _ = _legacy_cases()
    for name, (w, h, bpp, clip, keys) in mk.cases16().items():
        key = "v4rgb16/" + name
        data, sizes, types = st[key + "/data"], st[key + "/sizes"].astype(np.uint32), st[key + "/types"]
        n = len(clip)
        enc = _new(scpr, w, h, 16)
        stream, gsizes, gtypes = enc.CompressClip(clip.view(np.uint8).reshape(n, -1), keys)
        assert np.array_equal(gsizes, sizes) and np.array_equal(gtypes, types) and np.array_equal(stream, data), name
        enc2, pos = _new(scpr, w, h, 16), 0
        for i in range(n):   # the per-frame drop-in call
            d, ft = enc2.CompressFrame(clip[i].view(np.uint8).reshape(-1), 0 if keys[i] else 1)
            assert d == bytes(data[pos:pos + int(sizes[i])]) and ft == types[i], (name, i)
            pos += int(sizes[i])
        out = _new(scpr, w, h, 16).DecompressClip(data, sizes, types)
        assert np.array_equal(out.view(np.uint16).reshape(clip.shape), clip), name
        pitch = (w * 2 + 3 & ~3) + 8
        out = _new(scpr, w, h, 16).DecompressClip(data, sizes, types, pitch=pitch)
        assert np.array_equal(out.reshape(n, h, pitch)[:, :, : 2 * w].copy().view(np.uint16), clip), name


def test_vfw_session_and_avi_container(scpr, tmp_path):
    """the layer above the codec object: forced keyframe interval (npframes + 1 >= interval, screenpressor.cpp:402-406),
    quality -> loss, AVI written with AVIIF_KEYFRAME where Compress said so, decoded back through Decompress with the
    frame type inferred from the data"""
    from screenpressor_b200 import synth

    w, h, n = 200, 120, 30
    clip, _ = fuzz_clip(w, h, n, 41, 32, 16)
    inst = scpr.CodecInst(scpr.CodecParameters(w, h, 32), kf_interval=8)
    path = str(tmp_path / "cap.avi")
    produced = []
    with scpr.AviWriter(path, w, h, 32) as wr:
        for i in range(n):
            data, key = inst.Compress(clip[i])
            wr.write(data, key)
            produced.append((data, 0 if key else 1))
    keys = synth.keyframe_flags(n, 8)
    # flat frames are I frames whatever was asked (screencap.cpp:1488-1499): the session's counter restarts there too
    want = _split(*_new(scpr, w, h, 32).CompressClip(clip, keys))
    flat = [i for i in range(n) if len(want[i][0]) == 4 and want[i][1] == 0]
    if not flat:
        assert produced == want
    assert [p[1] for p in produced][:9] == [0, 1, 1, 1, 1, 1, 1, 1, 0][:9] or flat
    dec = scpr.CodecInst(scpr.CodecParameters(w, h, 32))
    with scpr.AviReader(path) as rd:
        assert len(rd) == n and rd.info.fourcc == scpr.FOURCC_SCPR and (rd.info.width, rd.info.height) == (w, h)
        for i in range(n):
            data, key = rd.read(i)
            assert (data, 0 if key else 1) == produced[i]
            out = dec.Decompress(data, w * 4, not key)
            assert np.array_equal(out, clip[i].reshape(-1)), i
    # quality drives the loss when it is not forced: 5000 -> 2 bits
    lossy = scpr.CodecInst(scpr.CodecParameters(w, h, 32), force_loss=False)
    ref = scpr.ScreenCodec(0)
    ref.Init(scpr.CodecParameters(w, h, 32))
    for i in range(4):
        assert lossy.Compress(clip[i], quality=5000)[0] == ref.CompressFrame(clip[i], 0 if i == 0 else 1, loss=2)[0]


def test_corrupt_streams_do_not_take_the_device_down(scpr):
    """truncated and random payloads: the decoder may produce garbage but must return, stay inside its buffers (reads past
    the stream see zeros) and leave the CUDA context usable"""
    w, h, n = 200, 120, 16
    clip, keys = fuzz_clip(w, h, n, 41, 32, 16)
    stream, sizes, fts = _new(scpr, w, h, 32).CompressClip(clip, keys)
    frames, pos = [], 0
    for i in range(n):
        frames.append(bytes(stream[pos:pos + int(sizes[i])])); pos += int(sizes[i])
    rng = np.random.default_rng(7)
    for trial in range(6):
        bad = list(frames)
        for i in range(n):
            if len(bad[i]) > 8 and rng.random() < 0.6:
                body = bytearray(bad[i])
                if trial % 2 == 0:
                    body = body[: 1 + len(body) // 3]                      # truncated
                else:
                    k = rng.integers(1, len(body), size=max(1, len(body) // 8))
                    for j in k:
                        body[j] = int(rng.integers(0, 256))                 # noise, header byte kept
                bad[i] = bytes(body)
        dec = _new(scpr, w, h, 32)
        data = np.frombuffer(b"".join(bad), np.uint8)
        out = dec.DecompressClip(data, np.array([len(b) for b in bad], np.uint32), fts)
        assert out.shape == (n, w * h * 4)
    for ver_byte in (0x32, 0x22, 0x12):   # pure noise behind a valid I-frame header of every generation
        dec = _new(scpr, w, h, 32)
        junk = bytes([ver_byte]) + bytes(rng.integers(0, 256, 4000, dtype=np.uint8))
        dec.DecompressFrame(junk, None, 0)
        dec.DecompressFrame(b"\x01" + bytes(rng.integers(0, 256, 500, dtype=np.uint8)), None, 1)
    good = _new(scpr, w, h, 32).DecompressClip(stream, sizes, fts)
    assert np.array_equal(good.reshape(n, -1), clip.reshape(n, -1))


def test_error_behaviour(scpr):
    dec = _new(scpr, 64, 48, 32)
    with pytest.raises(scpr.ScprError):          # P before any I: the reference returns 0 (screencap.cpp:1699)
        dec.DecompressFrame(b"\x01abcdefgh", None, 1)
    with pytest.raises(scpr.BadVersionException) as e:   # version nibble 7 -> BadVersionException(8)
        dec.DecompressFrame(b"\x72abcdefgh", None, 0)
    assert e.value.version == 8
    with pytest.raises(scpr.BadVersionException) as e:   # version nibble 0 -> BadVersionException(1): the pre-2.0 coder
        dec.DecompressFrame(b"\x02abcdefgh", None, 0)
    assert e.value.version == 1
    sc = scpr.ScreenCodec(0)
    with pytest.raises(scpr.ScprError):          # BadVersionException(48) in the reference (screencap.cpp:1607-1609)
        sc.Init(scpr.CodecParameters(64, 48, 8))


def test_smoke_entry(scpr):
    import __graft_entry__ as ge

    ge.smoke()
