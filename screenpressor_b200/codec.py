"""Host-side mirror of the reference's `ScreenCodec` object (screencap.h:519-541) over the C ABI of
`libscpr_b200.so` (include/scpr_c.h).  Same method names, argument meaning and error behaviour as
the reference class so parity tests read like code written against the reference:

    sc = ScreenCodec(); sc.Init(CodecParameters(1920, 1080, 32))
    data, ftype = sc.CompressFrame(frame, ftype=1)        # bytes, actual frame type
    out = sc.DecompressFrame(data, pitch, ftype)          # ndarray, raises BadVersionException
    sc.Deinit()

plus the throughput calls `CompressClip` / `DecompressClip` (many frames per call).  Every call runs
CUDA kernels; importing this module fails loudly if the extension has not been built, and creating a
codec fails loudly without a GPU -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SCPR_LIB") or os.path.join(_HERE, "libscpr_b200.so")  # SCPR_LIB: A/B runs of two builds (tools/ab.sh)

SCPR_E_CUDA, SCPR_E_PARAM, SCPR_E_UNSUPPORTED, SCPR_E_DSTSIZE, SCPR_E_NODEVICE = -1000, -1001, -1002, -1003, -1004


class ScprError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__(f"libscpr_b200 error {code}: {text}")
        self.code = code


class BadVersionException(Exception):
    """Mirror of the reference's BadVersionException (screencap.h:86-90)."""

    def __init__(self, version: int):
        super().__init__(f"cannot decode stream version {version}")
        self.version = version


class _Params(C.Structure):
    _fields_ = [
        ("width", C.c_uint32), ("height", C.c_uint32), ("bits_per_pixel", C.c_uint8),
        ("redmask", C.c_uint16), ("greenmask", C.c_uint16), ("bluemask", C.c_uint16),
        ("high_range_x", C.c_uint32), ("high_range_y", C.c_uint32),
        ("low_range_x", C.c_uint32), ("low_range_y", C.c_uint32), ("loss", C.c_uint32),
    ]


class _Clip(C.Structure):
    _fields_ = [("stream", C.c_void_p), ("sizes", C.c_void_p), ("ftypes", C.c_void_p), ("n", C.c_int), ("frames", C.c_void_p),
                ("result", C.c_int)]


class _Policy(C.Structure):
    _fields_ = [("force_interval", C.c_int), ("kf_interval", C.c_int), ("force_loss", C.c_int), ("conf_loss", C.c_int)]


class _AviInfo(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("bits_per_pixel", C.c_uint32), ("redmask", C.c_uint32),
                ("greenmask", C.c_uint32), ("bluemask", C.c_uint32), ("fps_num", C.c_uint32), ("fps_den", C.c_uint32),
                ("frames", C.c_uint32), ("fourcc", C.c_uint32)]


@dataclass
class CodecParameters:
    """reference screencap.h:49-55; defaults from screenpressor.cpp:374-379"""
    width: int
    height: int
    bits_per_pixel: int = 32
    redmask: int = 0x7C00
    greenmask: int = 0x3E0
    bluemask: int = 0x1F
    high_range_x: int = 256
    high_range_y: int = 256
    low_range_x: int = 8
    low_range_y: int = 8
    loss: int = 0

    def _c(self) -> "_Params":
        return _Params(self.width, self.height, self.bits_per_pixel, self.redmask, self.greenmask, self.bluemask, self.high_range_x,
                       self.high_range_y, self.low_range_x, self.low_range_y, self.loss)


_lib = None


def load_library() -> C.CDLL:
    """Load libscpr_b200.so; raises if it was not built (run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build the CUDA extension first (__graft_entry__.build()); "
                          "screenpressor_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64
    lib.scpr_create.restype = i32
    lib.scpr_create.argtypes = [C.POINTER(_Params), i32, C.POINTER(vp)]
    lib.scpr_destroy.restype = None
    lib.scpr_destroy.argtypes = [vp]
    lib.scpr_compress_frame.restype = i32
    lib.scpr_compress_frame.argtypes = [vp, vp, vp, i32, C.POINTER(i32), i32]
    lib.scpr_decompress_frame.restype = i32
    lib.scpr_decompress_frame.argtypes = [vp, vp, i32, vp, i32, i32]
    for name in ("scpr_compress_clip", "scpr_compress_clip_dev"):
        fn = getattr(lib, name)
        fn.restype = i64
        fn.argtypes = [vp, vp, i32, vp, vp, C.c_size_t, vp, vp]
    for name in ("scpr_decompress_clip", "scpr_decompress_clip_dev"):
        fn = getattr(lib, name)
        fn.restype = i32
        fn.argtypes = [vp, vp, vp, vp, i32, vp, i32]
    lib.scpr_reset.restype = i32
    lib.scpr_decompress_clips.restype = i32
    lib.scpr_decompress_clips.argtypes = [vp, C.POINTER(_Clip), i32, i32]
    lib.scpr_decompress_clips_dev.restype = i32
    lib.scpr_decompress_clips_dev.argtypes = [vp, C.POINTER(_Clip), i32, vp, i32]
    lib.scpr_compress_clip_multi.restype = C.c_int64
    lib.scpr_compress_clip_multi.argtypes = [C.POINTER(_Params), vp, i32, vp, i32, vp, vp, C.c_size_t, vp, vp, vp, vp]
    lib.scpr_decompress_clip_multi.restype = i32
    lib.scpr_decompress_clip_multi.argtypes = [C.POINTER(_Params), vp, i32, vp, vp, vp, i32, vp, i32]
    lib.scpr_set_threads_layout.restype = i32
    lib.scpr_set_threads_layout.argtypes = [vp, i32]
    lib.scpr_multi_create.restype = i32
    lib.scpr_multi_create.argtypes = [C.POINTER(_Params), vp, i32, C.POINTER(vp)]
    lib.scpr_multi_destroy.argtypes = [vp]
    lib.scpr_multi_compress_clip.restype = C.c_int64
    lib.scpr_multi_compress_clip.argtypes = [vp, vp, i32, vp, vp, C.c_size_t, vp, vp, vp, vp]
    lib.scpr_multi_decompress_clip.restype = i32
    lib.scpr_multi_decompress_clip.argtypes = [vp, vp, vp, vp, i32, vp, i32]
    lib.scpr_reset.argtypes = [vp]
    lib.scpr_set_stream.restype = i32
    lib.scpr_set_stream.argtypes = [vp, vp]
    lib.scpr_last_error.restype = C.c_char_p
    lib.scpr_kernel_launches.restype = u64
    lib.scpr_kernel_launches.argtypes = [vp]
    lib.scpr_max_compressed_size.restype = C.c_size_t
    lib.scpr_max_compressed_size.argtypes = [C.POINTER(_Params)]
    lib.scpr_debug_events.restype = i64
    lib.scpr_debug_events.argtypes = [vp, i32, vp, vp, C.c_size_t]
    lib.scpr_debug_blocks.restype = i32
    lib.scpr_debug_blocks.argtypes = [vp, i32, vp, vp, vp]
    lib.scpr_range_state_size.restype = C.c_size_t
    lib.scpr_range_state_size.argtypes = [vp, i32]
    lib.scpr_export_range_state.restype = i64
    lib.scpr_export_range_state.argtypes = [vp, vp, C.c_size_t, i32]
    lib.scpr_import_range_state.restype = i32
    lib.scpr_import_range_state.argtypes = [vp, vp, C.c_size_t]
    lib.scpr_set_mvs_hooks.restype = i32
    lib.scpr_set_mvs_hooks.argtypes = [vp, vp, vp, vp]
    lib.scpr_quality_to_loss.restype = i32
    lib.scpr_quality_to_loss.argtypes = [C.c_uint32]
    lib.scpr_infer_frame_type.restype = i32
    lib.scpr_infer_frame_type.argtypes = [C.c_uint8, C.c_uint32]
    lib.scpr_policy_default.restype = None
    lib.scpr_policy_default.argtypes = [C.POINTER(_Policy)]
    lib.scpr_session_create.restype = i32
    lib.scpr_session_create.argtypes = [C.POINTER(_Params), i32, C.POINTER(_Policy), C.POINTER(vp)]
    lib.scpr_session_destroy.restype = None
    lib.scpr_session_destroy.argtypes = [vp]
    lib.scpr_session_compress.restype = i32
    lib.scpr_session_compress.argtypes = [vp, vp, vp, i32, i32, C.c_uint32, C.POINTER(i32)]
    lib.scpr_session_decompress.restype = i32
    lib.scpr_session_decompress.argtypes = [vp, vp, i32, vp, i32, i32]
    lib.scpr_avi_create.restype = i32
    lib.scpr_avi_create.argtypes = [C.c_char_p, C.POINTER(_AviInfo), C.POINTER(vp)]
    lib.scpr_avi_write_frame.restype = i32
    lib.scpr_avi_write_frame.argtypes = [vp, vp, C.c_uint32, i32]
    lib.scpr_avi_open.restype = i32
    lib.scpr_avi_open.argtypes = [C.c_char_p, C.POINTER(vp), C.POINTER(_AviInfo)]
    lib.scpr_avi_read_frame.restype = i64
    lib.scpr_avi_read_frame.argtypes = [vp, C.c_uint32, vp, C.c_size_t, C.POINTER(i32)]
    lib.scpr_avi_close.restype = i32
    lib.scpr_avi_close.argtypes = [vp]
    lib.scpr_debug_frame_scan.restype = C.c_float
    lib.scpr_debug_frame_scan.argtypes = [vp, i32, vp, vp, i32, i32, vp, vp]
    lib.scpr_bench_frame_scan.restype = C.c_float
    lib.scpr_bench_frame_scan.argtypes = [vp, vp, i32, i32]
    _lib = lib
    return lib


def _ptr(a) -> int:
    """host ndarray / bytes-like -> address"""
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    raise TypeError(type(a))


class ScreenCodec:
    def __init__(self, device: int = 0):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.device = device
        self.params: CodecParameters | None = None

    # ---- reference interface -----------------------------------------------------------------
    def Init(self, params: CodecParameters) -> None:
        self.Deinit()
        p = _Params(params.width, params.height, params.bits_per_pixel, params.redmask, params.greenmask,
                    params.bluemask, params.high_range_x, params.high_range_y, params.low_range_x,
                    params.low_range_y, params.loss)
        self._check(self._lib.scpr_create(C.byref(p), self.device, C.byref(self._h)))
        self.params = params
        # bytes per row of a frame handed to CompressFrame: implicit in the reference (screencap.cpp:1655, 1668)
        self.pitch = {32: params.width * 4, 16: params.width * 2}.get(params.bits_per_pixel, (params.width * 3 + 3) & ~3)
        self.frame_bytes = self.pitch * params.height
        self.max_size = params.width * params.height * 6  # CompressGetSize, screenpressor.cpp:386-388
        self._dst = np.empty(self.max_size + 64, dtype=np.uint8)

    def Reset(self) -> None:
        """Deinit() + Init() with the same parameters, keeping device workspaces (start of a new clip)."""
        self._check(self._lib.scpr_reset(self._h))

    def Deinit(self) -> None:
        if self._h:
            self._lib.scpr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.Deinit()
        except Exception:
            pass

    def CompressFrame(self, src: np.ndarray, ftype: int, loss: int = 0):
        """-> (bytes, actual ftype).  ftype: 0 = I requested, 1 = P requested."""
        src = np.ascontiguousarray(src).reshape(-1)
        assert src.dtype == np.uint8 and src.size == self.frame_bytes, (src.size, self.frame_bytes)
        ft = C.c_int(ftype)
        n = self._check(self._lib.scpr_compress_frame(self._h, _ptr(src), _ptr(self._dst), self._dst.size, C.byref(ft), loss))
        return bytes(self._dst[:n]), ft.value

    def DecompressFrame(self, data: bytes, pitch: int | None = None, ftype: int = 0) -> np.ndarray:
        """-> decoded frame as a flat uint8 array of height*pitch bytes."""
        pitch = self.pitch if pitch is None else pitch
        src = np.frombuffer(data, dtype=np.uint8)
        out = np.zeros(self.params.height * pitch, dtype=np.uint8)
        r = self._lib.scpr_decompress_frame(self._h, _ptr(np.ascontiguousarray(src)), len(data), _ptr(out), pitch, ftype)
        if -16 <= r < 0:
            raise BadVersionException(-r)
        if r == 0:
            raise ScprError(0, "P frame before any I frame")
        self._check(r)
        return out

    # ---- throughput interface ------------------------------------------------------------------
    def CompressClip(self, frames, keyflags: np.ndarray, device_ptr: int | None = None, n: int | None = None):
        """frames: host ndarray of n frames back to back, or (device_ptr, n) for HBM-resident input.
        -> (stream bytes ndarray, sizes uint32[n], ftypes uint8[n])"""
        keyflags = np.ascontiguousarray(keyflags, dtype=np.uint8)
        if device_ptr is None:
            frames = np.ascontiguousarray(frames)
            n = frames.size // self.frame_bytes
            assert frames.size == n * self.frame_bytes
        assert keyflags.size == n
        cap = getattr(self, "_clip_cap", 0)
        sizes = np.zeros(n, dtype=np.uint32)
        ftypes = np.zeros(n, dtype=np.uint8)
        while True:
            if cap < n * self.max_size:
                cap = n * self.max_size  # W*H*6 per frame (CompressGetSize); np.empty commits pages lazily
            if getattr(self, "_clip_dst", None) is None or self._clip_dst.size < cap:
                self._clip_dst = np.empty(cap, dtype=np.uint8)
            self._clip_cap = cap
            if device_ptr is None:
                r = self._lib.scpr_compress_clip(self._h, _ptr(frames), n, _ptr(keyflags), _ptr(self._clip_dst), cap,
                                                 _ptr(sizes), _ptr(ftypes))
            else:
                r = self._lib.scpr_compress_clip_dev(self._h, device_ptr, n, _ptr(keyflags), _ptr(self._clip_dst), cap,
                                                     _ptr(sizes), _ptr(ftypes))
            if r == SCPR_E_DSTSIZE:
                raise ScprError(r, "destination too small: the frames of this call were dropped and the next frame will be coded as an I "
                                   "frame; pass a larger capacity via reserve_clip_output()")
            self._check(r)
            return self._clip_dst[:r], sizes, ftypes

    def set_threads_layout(self, n_threads: int) -> None:
        """I frames in the layout of the reference running with n worker threads (scpr_set_threads_layout); 1 = canonical"""
        self._check(self._lib.scpr_set_threads_layout(self._h, int(n_threads)))

    def reserve_clip_output(self, nbytes: int) -> None:
        self._clip_cap = int(nbytes)
        self._clip_dst = np.empty(self._clip_cap, dtype=np.uint8)

    def DecompressClip(self, stream: np.ndarray, sizes: np.ndarray, ftypes: np.ndarray, pitch: int | None = None,
                       device_ptr: int | None = None):
        """-> ndarray (n, height*pitch) of decoded frames, or None when decoding into device_ptr."""
        pitch = self.pitch if pitch is None else pitch
        n = int(sizes.size)
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint32)
        ftypes = np.ascontiguousarray(ftypes, dtype=np.uint8)
        if device_ptr is None:
            out = np.zeros((n, self.params.height * pitch), dtype=np.uint8)
            r = self._lib.scpr_decompress_clip(self._h, _ptr(stream), _ptr(sizes), _ptr(ftypes), n, _ptr(out), pitch)
        else:
            out = None
            r = self._lib.scpr_decompress_clip_dev(self._h, _ptr(stream), _ptr(sizes), _ptr(ftypes), n, device_ptr, pitch)
        if -16 <= r < 0:
            raise BadVersionException(-r)
        if r == 0:
            raise ScprError(0, "P frame before any I frame")
        self._check(r)
        return out

    def DecompressClips(self, clips, pitch: int | None = None, device_ptr: int | None = None, out: list | None = None):
        """Many independent clips in one call: every GOP of every clip is one thread block of the same launch
        (scpr_decompress_clips).  clips: [(stream, sizes, ftypes), ...], each starting with an I frame.
        -> (results, frames): results[k] = 1 / 0 / error code of clip k; frames[k] = ndarray (n_k, height*pitch), or None when
        decoding into device_ptr (all clips back to back).  `out`: optional preallocated host arrays, one per clip."""
        pitch = self.pitch if pitch is None else pitch
        arr = (_Clip * len(clips))()
        keep, frames = [], []
        for k, (stream, sizes, ftypes) in enumerate(clips):
            stream = np.ascontiguousarray(stream, dtype=np.uint8)
            sizes = np.ascontiguousarray(sizes, dtype=np.uint32)
            ftypes = np.ascontiguousarray(ftypes, dtype=np.uint8)
            keep.append((stream, sizes, ftypes))
            arr[k].stream, arr[k].sizes, arr[k].ftypes, arr[k].n = _ptr(stream), _ptr(sizes), _ptr(ftypes), int(sizes.size)
            if device_ptr is None:
                o = out[k] if out is not None else np.zeros((int(sizes.size), self.params.height * pitch), dtype=np.uint8)
                frames.append(o)
                arr[k].frames = _ptr(o)
        if device_ptr is None:
            r = self._lib.scpr_decompress_clips(self._h, arr, len(clips), pitch)
        else:
            r = self._lib.scpr_decompress_clips_dev(self._h, arr, len(clips), device_ptr, pitch)
        results = [int(arr[k].result) for k in range(len(clips))]
        if r in (SCPR_E_PARAM, SCPR_E_CUDA):
            self._check(r)
        return results, (frames if device_ptr is None else None)

    # ---- frame-range sharding / checkpoint (include/scpr_c.h) ----------------------------------------
    def ExportRangeState(self, full: bool = False) -> np.ndarray:
        """Encoder state the next frame range needs: mvs[] + counters (full=False, ranges that start on a keyframe)
        or everything including the previous frame and the adaptive models (full=True, any cut / checkpoint)."""
        blob = np.empty(self._lib.scpr_range_state_size(self._h, int(full)), dtype=np.uint8)
        n = self._check(self._lib.scpr_export_range_state(self._h, _ptr(blob), blob.size, int(full)))
        return blob[:n]

    def ImportRangeState(self, blob: np.ndarray) -> None:
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        self._check(self._lib.scpr_import_range_state(self._h, _ptr(blob), blob.size))

    def set_mvs_hooks(self, wait=None, ready=None) -> None:
        """Python callables invoked around the in-order motion-vector resolve of every compress call (scpr_set_mvs_hooks):
        `wait()` before it (import the previous range's vectors there), `ready()` after it (export and pass them on)."""
        HOOK = C.CFUNCTYPE(None, C.c_void_p)
        self._hooks = (HOOK(lambda _u: wait()) if wait else None, HOOK(lambda _u: ready()) if ready else None)  # keep alive
        self._check(self._lib.scpr_set_mvs_hooks(self._h, C.cast(self._hooks[0], C.c_void_p) if wait else None,
                                                 C.cast(self._hooks[1], C.c_void_p) if ready else None, None))

    # ---- plumbing / test hooks -------------------------------------------------------------------
    def set_stream(self, cuda_stream: int) -> None:
        self._check(self._lib.scpr_set_stream(self._h, cuda_stream))

    def kernel_launches(self) -> int:
        return int(self._lib.scpr_kernel_launches(self._h))

    def debug_events(self, frame: int):
        n = self._check(self._lib.scpr_debug_events(self._h, frame, None, None, 0))
        ev = np.zeros(n, dtype=np.uint32)
        iv = np.zeros(n, dtype=np.uint32)
        if n:
            self._check(self._lib.scpr_debug_events(self._h, frame, _ptr(ev), _ptr(iv), n))
        return ev, iv

    def debug_blocks(self, frame: int):
        nb = ((self.params.width + 15) // 16) * ((self.params.height + 15) // 16)
        bts = np.zeros(nb, dtype=np.uint8)
        sxy = np.zeros((nb, 4), dtype=np.int32)
        mv = np.zeros((nb, 2), dtype=np.int32)
        self._check(self._lib.scpr_debug_blocks(self._h, frame, _ptr(bts), _ptr(sxy), _ptr(mv)))
        return bts, sxy, mv

    def bench_frame_scan(self, device_ptr: int, n: int, reps: int) -> float:
        ms = float(self._lib.scpr_bench_frame_scan(self._h, device_ptr, n, reps))
        if ms < 0:
            raise ScprError(int(ms), self._lib.scpr_last_error().decode())
        return ms

    def debug_frame_scan(self, mode: int, device_ptr: int, prev_ptr: int, n: int, reps: int = 1, fetch: bool = True):
        """scpr_debug_frame_scan: one chosen frame-scan kernel over device frames -> (ms, blkinfo (n, nb) u32, summary (n, 4) u32)"""
        nb = ((self.params.width + 15) // 16) * ((self.params.height + 15) // 16)
        bi = np.zeros((n, nb), dtype=np.uint32) if fetch else None
        sm = np.zeros((n, 4), dtype=np.uint32) if fetch else None
        ms = float(self._lib.scpr_debug_frame_scan(self._h, mode, device_ptr, prev_ptr or None, n, reps, _ptr(bi) if fetch else None,
                                                   _ptr(sm) if fetch else None))
        if ms < 0:
            raise ScprError(int(ms), self._lib.scpr_last_error().decode())
        return ms, bi, sm

    def _check(self, r: int) -> int:
        if r < 0:
            raise ScprError(int(r), self._lib.scpr_last_error().decode())
        return int(r)


# ---- host layer: the VfW policy around the codec object, and the AVI container (csrc/vfw_host.cpp) -----------------
FOURCC_SCPR = 0x52504353  # mmioFOURCC('S','C','P','R'), screenpressor.h:6


def quality_to_loss(quality: int) -> int:
    return int(load_library().scpr_quality_to_loss(quality))


def infer_frame_type(first_byte: int, data_size: int) -> int:
    return int(load_library().scpr_infer_frame_type(first_byte, data_size))


class CodecInst:
    """Mirror of the reference's CodecInst::Compress / Decompress (screenpressor.cpp:392-437, 592-620): keyframe interval
    policy, quality -> loss, frame type inferred from the data on decode."""

    def __init__(self, params: CodecParameters, device: int = 0, force_interval: bool = True, kf_interval: int = 500,
                 force_loss: bool = True, conf_loss: int = 0):
        self._lib = load_library()
        self._h = C.c_void_p()
        p = _Params(params.width, params.height, params.bits_per_pixel, params.redmask, params.greenmask, params.bluemask,
                    params.high_range_x, params.high_range_y, params.low_range_x, params.low_range_y, params.loss)
        pol = _Policy(int(force_interval), kf_interval, int(force_loss), conf_loss)
        r = self._lib.scpr_session_create(C.byref(p), device, C.byref(pol), C.byref(self._h))
        if r < 0:
            raise ScprError(r, self._lib.scpr_last_error().decode())
        self.params = params
        self._dst = np.empty(params.width * params.height * 6 + 64, dtype=np.uint8)

    def close(self):
        if self._h:
            self._lib.scpr_session_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def Compress(self, frame: np.ndarray, host_keyframe: bool = False, quality: int = 10000):
        """-> (bytes, is_keyframe)"""
        frame = np.ascontiguousarray(frame).reshape(-1).view(np.uint8)
        key = C.c_int(0)
        n = self._lib.scpr_session_compress(self._h, _ptr(frame), _ptr(self._dst), self._dst.size, int(host_keyframe), quality, C.byref(key))
        if n < 0:
            raise ScprError(n, self._lib.scpr_last_error().decode())
        return bytes(self._dst[:n]), bool(key.value)

    def Decompress(self, data: bytes, pitch: int, not_keyframe: bool) -> np.ndarray:
        src = np.frombuffer(data, dtype=np.uint8)
        out = np.zeros(self.params.height * pitch, dtype=np.uint8)
        r = self._lib.scpr_session_decompress(self._h, _ptr(np.ascontiguousarray(src)), len(data), _ptr(out), pitch, int(not_keyframe))
        if -16 <= r < 0:
            raise BadVersionException(-r)
        if r != 1:
            raise ScprError(r, self._lib.scpr_last_error().decode() if r < 0 else "P frame before any I frame")
        return out


def compress_clip_multi(params: "CodecParameters", devices, frames: np.ndarray, keyflags: np.ndarray):
    """One clip cut by GOP-aligned frame ranges across `devices` (CUDA ordinals, one range each) from this process
    (scpr_compress_clip_multi).  -> (stream, sizes, ftypes, [first frame of every range])"""
    lib = load_library()
    p = params._c()
    frames = np.ascontiguousarray(frames)
    keyflags = np.ascontiguousarray(keyflags, dtype=np.uint8)
    n = int(keyflags.size)
    dev = np.ascontiguousarray(devices, dtype=np.int32)
    cap = min(n * params.width * params.height * 6, 1 << 31)
    dst = np.empty(cap, dtype=np.uint8)
    sizes, ftypes = np.zeros(n, np.uint32), np.zeros(n, np.uint8)
    first, nr = np.zeros(len(dev), np.int32), C.c_int(0)
    r = lib.scpr_compress_clip_multi(C.byref(p), _ptr(dev), len(dev), _ptr(frames), n, _ptr(keyflags), _ptr(dst), cap, _ptr(sizes), _ptr(ftypes),
                                     _ptr(first), C.addressof(nr))
    if r < 0:
        raise ScprError(int(r), lib.scpr_last_error().decode())
    return dst[:r].copy(), sizes, ftypes, [int(x) for x in first[:nr.value]]


def decompress_clip_multi(params: "CodecParameters", devices, stream, sizes, ftypes, pitch: int | None = None) -> np.ndarray:
    """The decoding counterpart (scpr_decompress_clip_multi): ranges start at coded I frames.  -> ndarray (n, height*pitch)"""
    lib = load_library()
    p = params._c()
    bpp = params.bits_per_pixel // 8
    pitch = pitch or (((params.width * 3 + 3) & ~3) if bpp == 3 else params.width * bpp)
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    sizes = np.ascontiguousarray(sizes, dtype=np.uint32)
    ftypes = np.ascontiguousarray(ftypes, dtype=np.uint8)
    dev = np.ascontiguousarray(devices, dtype=np.int32)
    out = np.zeros((int(sizes.size), params.height * pitch), dtype=np.uint8)
    r = lib.scpr_decompress_clip_multi(C.byref(p), _ptr(dev), len(dev), _ptr(stream), _ptr(sizes), _ptr(ftypes), int(sizes.size), _ptr(out), pitch)
    if r != 1:
        raise ScprError(int(r), lib.scpr_last_error().decode())
    return out


class MultiCodec:
    """A standing set of codec objects over several GPUs driven from this process (scpr_multi_*): one clip per call, cut by
    GOP-aligned frame ranges, byte-identical to one codec."""

    def __init__(self, params: "CodecParameters", devices):
        self._lib = load_library()
        self.params = params
        self.devices = np.ascontiguousarray(devices, dtype=np.int32)
        h = C.c_void_p()
        p = params._c()
        r = self._lib.scpr_multi_create(C.byref(p), _ptr(self.devices), len(self.devices), C.byref(h))
        if r < 0:
            raise ScprError(int(r), self._lib.scpr_last_error().decode())
        self._h = h
        bpp = params.bits_per_pixel // 8
        self.pitch = ((params.width * 3 + 3) & ~3) if bpp == 3 else params.width * bpp
        self._dst = None

    def close(self) -> None:
        if self._h:
            self._lib.scpr_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def compress_clip(self, frames_ptr: int, keyflags: np.ndarray, cap: int = 256 << 20):
        """frames_ptr: host address of n frames back to back (pinned memory uploads fastest) -> (stream view, sizes, ftypes, firsts)"""
        keyflags = np.ascontiguousarray(keyflags, dtype=np.uint8)
        n = int(keyflags.size)
        if self._dst is None or self._dst.size < cap:
            self._dst = np.empty(cap, dtype=np.uint8)
        sizes, ftypes = np.zeros(n, np.uint32), np.zeros(n, np.uint8)
        first, nr = np.zeros(len(self.devices), np.int32), C.c_int(0)
        r = self._lib.scpr_multi_compress_clip(self._h, frames_ptr, n, _ptr(keyflags), _ptr(self._dst), self._dst.size, _ptr(sizes), _ptr(ftypes),
                                               _ptr(first), C.addressof(nr))
        if r < 0:
            raise ScprError(int(r), self._lib.scpr_last_error().decode())
        return self._dst[:r], sizes, ftypes, [int(x) for x in first[:nr.value]]

    def decompress_clip(self, stream: np.ndarray, sizes: np.ndarray, ftypes: np.ndarray, out_ptr: int, pitch: int | None = None) -> None:
        """out_ptr: host address of the destination frames (n * height * pitch bytes)"""
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        sizes = np.ascontiguousarray(sizes, dtype=np.uint32)
        ftypes = np.ascontiguousarray(ftypes, dtype=np.uint8)
        r = self._lib.scpr_multi_decompress_clip(self._h, _ptr(stream), _ptr(sizes), _ptr(ftypes), int(sizes.size), out_ptr, pitch or self.pitch)
        if r != 1:
            raise ScprError(int(r), self._lib.scpr_last_error().decode())


class AviWriter:
    def __init__(self, path: str, width: int, height: int, bits_per_pixel: int = 32, fps: tuple[int, int] = (30, 1),
                 masks: tuple[int, int, int] = (0x7C00, 0x3E0, 0x1F)):
        self._lib = load_library()
        self._h = C.c_void_p()
        info = _AviInfo(width, height, bits_per_pixel, masks[0], masks[1], masks[2], fps[0], fps[1], 0, FOURCC_SCPR)
        r = self._lib.scpr_avi_create(path.encode(), C.byref(info), C.byref(self._h))
        if r < 0:
            raise ScprError(r, self._lib.scpr_last_error().decode())

    def write(self, data: bytes, is_key: bool) -> None:
        buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, np.uint8)
        r = self._lib.scpr_avi_write_frame(self._h, _ptr(np.ascontiguousarray(buf)), len(data), int(is_key))
        if r < 0:
            raise ScprError(r, self._lib.scpr_last_error().decode())

    def close(self) -> None:
        if self._h:
            self._lib.scpr_avi_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class AviReader:
    def __init__(self, path: str):
        self._lib = load_library()
        self._h = C.c_void_p()
        self.info = _AviInfo()
        r = self._lib.scpr_avi_open(path.encode(), C.byref(self._h), C.byref(self.info))
        if r < 0:
            raise ScprError(r, self._lib.scpr_last_error().decode())

    def __len__(self):
        return int(self.info.frames)

    def read(self, i: int):
        """-> (bytes, is_keyframe)"""
        key = C.c_int(0)
        n = self._lib.scpr_avi_read_frame(self._h, i, None, 0, C.byref(key))
        if n < 0:
            raise ScprError(int(n), "no such frame")
        buf = np.empty(max(int(n), 1), dtype=np.uint8)
        self._lib.scpr_avi_read_frame(self._h, i, _ptr(buf), buf.size, C.byref(key))
        return bytes(buf[:n]), bool(key.value)

    def close(self) -> None:
        if self._h:
            self._lib.scpr_avi_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
