"""ctypes bindings for the two CPU checkers -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

RefCodec    : the unmodified reference core (oracle/_ref/libscpr_ref.so, built by `make ref`).
OracleCodec : the plain-C restatement (oracle/libscpr_oracle.so, built by `make oracle`).
Both expose compress(frame, want_p) -> (bytes, ftype) and decompress(data, ftype) -> ndarray.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libscpr_ref.so")
ORACLE_SO = os.path.join(HERE, "libscpr_oracle.so")


def build(quiet: bool = True) -> None:
    """Compile the C restatement and, when the reference sources are present, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle", "ref", "ref_timing"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _stride(w: int, bpp: int) -> int:
    return (w * bpp // 8 + 3) & ~3


class _Base:
    prefix = ""
    so = ""

    def __init__(self, width: int, height: int, bpp: int = 32, loss: int = 0, threads: int = 1):
        if not os.path.exists(self.so):
            build()
        self.lib = C.CDLL(self.so)
        p = self.prefix
        self._create = getattr(self.lib, p + "create")
        self._create.restype = C.c_void_p
        self._create.argtypes = [C.c_int] * 5
        self._destroy = getattr(self.lib, p + "destroy")
        self._destroy.argtypes = [C.c_void_p]
        self._compress = getattr(self.lib, p + "compress")
        self._compress.restype = C.c_int
        self._compress.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_int]
        self._decompress = getattr(self.lib, p + "decompress")
        self._decompress.restype = C.c_int
        self._decompress.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        self.w, self.h, self.bpp, self.loss = width, height, bpp, loss
        self.pitch = _stride(width, bpp)
        self.h_ = self._create(width, height, bpp, loss, threads)
        self.cap = width * height * 6 + 64  # CompressGetSize, screenpressor.cpp:386-388
        self.dst = np.empty(self.cap, dtype=np.uint8)

    def close(self) -> None:
        if self.h_:
            self._destroy(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compress(self, frame: np.ndarray, want_p: bool, loss: int | None = None):
        """frame: contiguous uint8 of H*pitch bytes.  NOTE the reference may modify a 24 bpp source."""
        assert frame.flags["C_CONTIGUOUS"] and frame.size == self.h * self.pitch, (frame.shape, self.pitch)
        ft = C.c_int(1 if want_p else 0)
        n = self._compress(self.h_, frame.ctypes.data, self.dst.ctypes.data, self.cap, C.byref(ft),
                           self.loss if loss is None else loss)
        return bytes(self.dst[:n]), ft.value

    def decompress(self, data: bytes, ftype: int, pitch: int | None = None) -> np.ndarray:
        pitch = self.pitch if pitch is None else pitch
        out = np.zeros(self.h * pitch, dtype=np.uint8)
        src = np.frombuffer(data, dtype=np.uint8).copy()
        # the decoder's refill may read a few bytes past the end of the frame
        src = np.concatenate([src, np.zeros(16, dtype=np.uint8)])
        r = self._decompress(self.h_, src.ctypes.data, len(data), out.ctypes.data, pitch, ftype)
        if r != 1:
            raise RuntimeError(f"decompress returned {r}")
        return out


class RefCodec(_Base):
    prefix = "ref_"
    so = REF_SO


class OracleCodec(_Base):
    prefix = "orc_"
    so = ORACLE_SO
