"""CPU: live differential of the plain-C oracle against the compiled, unmodified reference
(oracle/_ref/libscpr_ref.so).  Skipped when neither the prebuilt library nor the reference sources
are available."""
import ctypes as C

import numpy as np
import pytest

from _clips import fuzz_clip


@pytest.fixture(scope="module")
def both(oracle_built):
    if not oracle_built.have_ref():
        pytest.skip("oracle/_ref not built (reference sources absent)")
    return oracle_built


def _run(py, w, h, n, seed, bpp, levels, loss=0):
    clip, keys = fuzz_clip(w, h, n, seed, bpp, levels)
    ref, orc = py.RefCodec(w, h, bpp, loss), py.OracleCodec(w, h, bpp, loss)
    dref, dorc = py.RefCodec(w, h, bpp, loss), py.OracleCodec(w, h, bpp, loss)
    for i in range(n):
        fr = np.ascontiguousarray(clip[i]).reshape(-1)
        a, fa = ref.compress(fr.copy(), not keys[i])
        b, fb = orc.compress(fr.copy(), not keys[i])
        assert (a, fa) == (b, fb), (w, h, seed, bpp, levels, loss, i, len(a), len(b))
        o1, o2 = dorc.decompress(a, fa), dref.decompress(a, fa)
        assert np.array_equal(o1, o2)
        if loss == 0:
            assert np.array_equal(o1, fr)


# widths with and without row padding, tiny frames, frames smaller than one block.  X = 2 is left
# out: the reference itself walks off the row there (screencap.cpp:894-899 increments x past X).
SIZES = [(97, 45), (1001, 37), (64, 64), (33, 17), (16, 16), (255, 31), (18, 2), (3, 40), (5, 5), (4, 3)]


@pytest.mark.parametrize("size", SIZES)
@pytest.mark.parametrize("bpp", [32, 24])
def test_fuzz_byte_exact(both, size, bpp):
    for levels in (256, 4, 16, 64):
        _run(both, size[0], size[1], 30, size[0] * 7 + levels, bpp, levels)


@pytest.mark.parametrize("loss", [1, 2, 3, 4])
def test_lossy_modes_byte_exact(both, loss):
    _run(both, 200, 100, 24, 77 + loss, 32, 16, loss)
    _run(both, 201, 100, 24, 78 + loss, 24, 16, loss)


def test_all_model_promotions_exercised(both):
    """The fuzz content must drive every colour-model promotion of ans_contexts.cpp:3-50."""
    _run(both, 640, 200, 6, 3, 32, 256)
    _run(both, 320, 200, 6, 4, 32, 64)
    _run(both, 640, 360, 30, 15, 32, 256)  # the one fuzz clip that takes a context through Cx2 -> Cx3 -> Cx7 (this test stands alone)
    lib = C.CDLL(both.ORACLE_SO)
    lib.orc_transition_count.restype = C.c_ulong
    for a, b in [(0, 1), (1, 2), (1, 4), (1, 5), (2, 3), (2, 6), (3, 7), (4, 5), (5, 6), (6, 7)]:
        assert lib.orc_transition_count(a, b) > 0, f"promotion {a}->{b} never happened"


@pytest.mark.parametrize("case", [(320, 192, 32), (161, 90, 24), (640, 360, 32)])
def test_oracle_band_layout_matches_multithreaded_reference(oracle_built, case):
    """SURVEY 8(f)5 on the CPU: the plain-C restatement with `threads` row bands against the unmodified reference created with as
    many worker threads, intra-only clips (the reference's multi-threaded P frames are timing dependent), bytes and decode."""
    if not oracle_built.have_ref():
        pytest.skip("oracle/_ref not built")
    from _clips import band_clip

    w, h, bpp = case
    n = 4
    clip = band_clip(w, h, n, 900 + w, bpp)
    canonical = None
    for threads in (1, 2, 3, 5, (h + 15) // 16):
        ref = oracle_built.RefCodec(w, h, bpp, threads=threads)
        want = [ref.compress(np.ascontiguousarray(clip[i]).reshape(-1).copy(), False) for i in range(n)]
        orc = oracle_built.OracleCodec(w, h, bpp, threads=threads)
        got = [orc.compress(np.ascontiguousarray(clip[i]).reshape(-1).copy(), False) for i in range(n)]
        assert got == want, (case, threads, [i for i in range(n) if got[i] != want[i]])
        if threads == 1:
            canonical = want
        else:
            assert want != canonical
        dec = oracle_built.OracleCodec(w, h, bpp)
        for i in range(n):
            assert np.array_equal(dec.decompress(want[i][0], want[i][1]), clip[i].reshape(-1)), (case, threads, i)
