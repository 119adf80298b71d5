"""CPU: the C-ABI library builds, loads and exports every symbol include/scpr_c.h declares; the host
mirror refuses to work without a GPU instead of falling back to anything."""
import ctypes as C
import os
import re
import subprocess

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
