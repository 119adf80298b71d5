/* include/scpr_c.h -- C ABI of the B200-native ScreenPressor v4 frame codec (libscpr_b200.so).
 *
 * Drop-in boundary: these entry points replace the reference's `ScreenCodec` object
 * (reference screencap.h:519-541, implementation screencap.cpp:1560-1743), which is what the
 * VfW layer `CodecInst` holds and calls (screenpressor.cpp:381, 425, 575, 620, 444, 647).
 * Plain pointers and sizes only; no exceptions cross this boundary; all work runs as CUDA
 * kernels on the selected device -- there is no CPU fallback and nothing here touches oracle/.
 *
 * Bitstream: byte-identical to the reference encoder run with one worker thread (its only
 * deterministic configuration); decoding reproduces frames bit-exactly.
 */
#ifndef SCPR_C_H
#define SCPR_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* replaces struct CodecParameters (reference screencap.h:49-55) */
typedef struct scpr_params {
    uint32_t width, height;
    uint8_t bits_per_pixel;                  /* 16, 24 or 32 */
    uint16_t redmask, greenmask, bluemask;   /* 16 bpp only: channel masks, e.g. 0x7C00 / 0x3E0 / 0x1F (screencap.cpp:1575-1583) */
    uint32_t high_range_x, high_range_y;     /* motion search ranges; v3/v4 clamp these to 256 (screencap.cpp:79); a v2 stream's
                                                motion vectors are offset by them (1..256) */
    uint32_t low_range_x, low_range_y;       /* 8, 8 (screenpressor.cpp:377-378) */
    uint32_t loss;                           /* bits of loss, 0..4 (quality -> loss, screenpressor.cpp:418-422) */
} scpr_params;

typedef struct scpr_codec scpr_codec;

enum {
    SCPR_OK = 0,
    SCPR_E_CUDA = -1000,        /* a CUDA call failed; scpr_last_error() has the text */
    SCPR_E_PARAM = -1001,       /* bad argument */
    SCPR_E_UNSUPPORTED = -1002, /* feature outside the built hot path (a v2 stream with a motion range above 256) */
    SCPR_E_DSTSIZE = -1003,     /* destination buffer too small */
    SCPR_E_NODEVICE = -1004     /* no CUDA device: this library never computes on the CPU */
};

/* replaces ScreenCodec::ScreenCodec + Init (screencap.cpp:1560-1584).  `device` = CUDA ordinal. */
int scpr_create(const scpr_params* p, int device, scpr_codec** out);
/* replaces ScreenCodec::Deinit / ~ScreenCodec (screencap.cpp:1619-1629) */
void scpr_destroy(scpr_codec* c);

/* Deinit() followed by Init() with the same parameters (what CodecInst does between clips,
 * screenpressor.cpp:441-447, 343-384), keeping the device workspaces: the next frame starts a new clip. */
int scpr_reset(scpr_codec* c);

/* replaces ScreenCodec::CompressFrame (screencap.cpp:1632-1692).
 * src: host frame, rows top to bottom, pitch width*4 (32 bpp), (width*3+3)&~3 (24 bpp) or width*2 (16 bpp).
 * *ftype in: 0 = I requested, 1 = P requested; out: type actually coded (first and flat frames
 * are always I, screencap.cpp:1488-1511).  loss = bits of loss for this frame (the clip entry points use
 * the value given at creation).  Returns the byte count written to dst, or < 0.
 * The source is never written to.  (The reference does write to a 24 bpp source: it zeroes the row padding when the width
 * is not a multiple of 4 and applies the loss mask in the caller's buffer, screencap.cpp:215-219, 857-859; a host that relied
 * on seeing the masked pixels must mask its own copy.  32 and 16 bpp sources are repacked first by the reference as well.)
 * Errors: after any error that is returned once device work has started (SCPR_E_CUDA, SCPR_E_DSTSIZE, "batch too large")
 * the frames of that call are not part of the stream and the next coded frame is forced to be an I frame with fresh
 * models, so the stream the caller holds stays decodable. */
int scpr_compress_frame(scpr_codec* c, const uint8_t* src, uint8_t* dst, int dst_cap, int* ftype, int loss);

/* replaces ScreenCodec::DecompressFrame (screencap.cpp:1695-1743).
 * Stream generations 2 (range coder, ScreenPressor 2.x), 3 and 4 (ANS) are decoded, as by the reference.
 * Returns 1 on success, 0 for "P frame before any I frame", -v (v = 1, 5..16) for a stream version
 * nobody can decode (the reference throws BadVersionException(v), screencap.cpp:1589-1590), or an SCPR_E_* code. */
int scpr_decompress_frame(scpr_codec* c, const uint8_t* src, int src_len, uint8_t* dst, int pitch, int ftype);

/* Bitstream layout of the multi-threaded reference, I frames only.  The reference splits pixel classification of
 * an I frame into one row band per worker thread (CSquadWorker::GetSegment, squad.cpp:16-31; CMD_CLASSIFYPIXELSI,
 * screencap.cpp:862-866) and every band starts a new run (ClassifyPixelsI :876-919, serialised band by band :365-388), so the bytes
 * of an I frame depend -- deterministically -- on its thread count.  n_threads = 1 (default) is the canonical stream; n > 1 writes
 * I frames byte-identical to the reference with n worker threads (n <= block rows of the frame, the limit of the reference's own
 * per-thread table).  P frames of the multi-threaded reference depend on thread timing and cannot be reproduced by anyone
 * (SURVEY.md 0.1): they are always written in the canonical 1-thread layout, which every reference decoder reads.  Decoding needs
 * no setting. */
int scpr_set_threads_layout(scpr_codec* c, int n_threads);

/* ---- throughput entry points (no reference equivalent): many frames per call -------------
 * The frames of one call are processed as a batch: frame differencing, block typing, motion
 * search, pixel typing and event generation run in parallel over all frames; the adaptive
 * models are replayed per GOP in the reference's update order; every rANS block is its own
 * stream.  Results are identical to n calls of scpr_compress_frame.
 *
 * frames   : n frames back to back (same layout as scpr_compress_frame's src).
 * keyflags : n bytes, 1 = the host requests an I frame (ftype 0), 0 = P requested.
 * dst      : concatenated frame bitstreams; sizes[i] / ftypes[i] describe frame i.
 * Returns total bytes written or < 0.  `*_dev` variants take device pointers for the frames
 * (inputs/outputs already resident in HBM, 16-byte aligned, never written to by the encoder); bitstreams stay host-side in both. */
int64_t scpr_compress_clip(scpr_codec* c, const uint8_t* frames, int n, const uint8_t* keyflags,
                           uint8_t* dst, size_t dst_cap, uint32_t* sizes, uint8_t* ftypes);
int64_t scpr_compress_clip_dev(scpr_codec* c, const uint8_t* d_frames, int n, const uint8_t* keyflags,
                               uint8_t* dst, size_t dst_cap, uint32_t* sizes, uint8_t* ftypes);
/* stream: concatenated frame bitstreams, sizes[i] bytes each, ftypes[i] = 0 I / 1 P.
 * frames: n decoded frames back to back with row pitch `pitch` (may differ from call to call; only width * bytes-per-pixel bytes
 * of every row are written, the rest of the pitch is the caller's).  Returns 1, 0 or < 0 as above. */
int scpr_decompress_clip(scpr_codec* c, const uint8_t* stream, const uint32_t* sizes, const uint8_t* ftypes,
                         int n, uint8_t* frames, int pitch);
int scpr_decompress_clip_dev(scpr_codec* c, const uint8_t* stream, const uint32_t* sizes, const uint8_t* ftypes,
                             int n, uint8_t* d_frames, int pitch);

/* Many independent clips (or GOPs cut out of a long clip) in ONE call: decoding is one serial dependency chain per GOP
 * (SURVEY.md 0.4), so what fills a GPU is the number of chains in flight -- every GOP of every clip becomes one thread block of
 * the same launch.  Each clip must start with an I frame and is decoded as by a fresh codec; `result` reports per clip what
 * scpr_decompress_clip would have returned (1, 0 = starts with a P frame, -v / SCPR_E_* = cannot be decoded); the call returns 1
 * when every clip decoded, else the first failure.  The decoder state of `c` is reset by the call.
 *   scpr_decompress_clips     : clips[k].frames = host destination of clip k (n frames back to back, row pitch `pitch`).
 *   scpr_decompress_clips_dev : all frames go to `d_frames` (device memory), clip after clip in the order given; clips[k].frames is ignored. */
typedef struct scpr_clip {
    const uint8_t* stream;   /* the clip's frame bitstreams back to back */
    const uint32_t* sizes;   /* n sizes */
    const uint8_t* ftypes;   /* n frame types, 0 = I, 1 = P */
    int n;
    uint8_t* frames;         /* host destination (scpr_decompress_clips) */
    int result;              /* out */
} scpr_clip;
int scpr_decompress_clips(scpr_codec* c, scpr_clip* clips, int n_clips, int pitch);
int scpr_decompress_clips_dev(scpr_codec* c, scpr_clip* clips, int n_clips, uint8_t* d_frames, int pitch);

/* ---- frame-range sharding and checkpoint / resume of the encoder (no reference equivalent) ------
 * A clip is cut into contiguous frame ranges that are encoded by different codec objects (one per GPU) and the
 * bitstreams concatenated on the host.  What the reference's encoder carries from one frame to
 * the next, and what a byte-identical continuation therefore needs:
 *   - always: the persistent motion-vector array mvs[] -- never cleared, not even by an I frame (screencap.cpp:96-97,
 *     715-735) --, the frame counter and the last flat colour (screencap.cpp:1488-1511);
 *   - for a range that starts on a P frame also the previous frame and the adaptive models (20 MB), i.e. full = 1.
 * export after the last frame of a range, import into the codec that encodes the next range (created with the same
 * parameters) before its first frame.  A range that starts on a keyframe only needs the small blob (full = 0,
 * 8 bytes per 16x16 block).  full = 1 doubles as checkpoint / resume of an encode. */
/* Pipelining across ranges: mvs[] of a range is final as soon as its in-order motion-vector resolve has run, long
 * before its entropy stages finish.  With hooks set, every compress call invokes wait(user) on the calling thread
 * right before that resolve is enqueued (frame scan, motion search and candidate matching are already running) and
 * ready(user) right after it completed.  wait typically receives the previous range's blob and calls
 * scpr_import_range_state (which then takes only the vectors); ready calls scpr_export_range_state(full = 0) and
 * sends it on.  Range k's resolve thus starts when range k-1's ends and everything else overlaps. */
int scpr_set_mvs_hooks(scpr_codec* c, void (*wait)(void* user), void (*ready)(void* user), void* user);
size_t scpr_range_state_size(const scpr_codec* c, int full);
int64_t scpr_export_range_state(scpr_codec* c, uint8_t* blob, size_t cap, int full);   /* bytes written or < 0 */
int scpr_import_range_state(scpr_codec* c, const uint8_t* blob, size_t len);

/* One clip cut across the GPUs of a box from ONE host process (csrc/multi.cu): GOP-aligned contiguous frame ranges, one per entry of
 * `devices` (an ordinal may repeat; fewer ranges when the clip has fewer GOPs), a codec object and a host thread per range, the
 * bitstreams concatenated on the host in frame order.  mvs[] travels device to device (cudaMemcpyPeerAsync) between the in-order
 * resolves of neighbouring ranges; a range never starts at a requested keyframe that is a single-colour frame (the reference does
 * not start a GOP there).  The result is byte-identical to scpr_compress_clip on one codec.  frames: host memory.
 * range_first / n_ranges (optional): the plan that was used. */
int64_t scpr_compress_clip_multi(const scpr_params* p, const int* devices, int n_dev, const uint8_t* frames, int n, const uint8_t* keyflags,
                                 uint8_t* dst, size_t dst_cap, uint32_t* sizes, uint8_t* ftypes, int* range_first, int* n_ranges);
/* The decoding counterpart: ranges start at coded I frames, no hand-off.  frames: host memory, row pitch `pitch`. */
int scpr_decompress_clip_multi(const scpr_params* p, const int* devices, int n_dev, const uint8_t* stream, const uint32_t* sizes,
                               const uint8_t* ftypes, int n, uint8_t* frames, int pitch);
/* The plan those entries use, by itself (host arithmetic only, no device): cuts[f] = 1 where a range may start at frame f (a coded,
 * non-flat keyframe; frame 0 always may).  Writes up to n_ranges_max GOP-aligned contiguous ranges balanced by frame count and
 * returns how many (fewer when the clip has fewer GOPs). */
int scpr_plan_ranges(const uint8_t* cuts, int n, int n_ranges_max, int* first, int* count);
/* The same with a standing set of codec objects (an encoder and a decoder per device ordinal, created once: workspaces and model
 * states stay on the devices between calls).  Every call codes a clip of its own (the objects are reset first). */
typedef struct scpr_multi scpr_multi;
int scpr_multi_create(const scpr_params* p, const int* devices, int n_dev, scpr_multi** out);
void scpr_multi_destroy(scpr_multi* m);
int64_t scpr_multi_compress_clip(scpr_multi* m, const uint8_t* frames, int n, const uint8_t* keyflags, uint8_t* dst, size_t dst_cap,
                                 uint32_t* sizes, uint8_t* ftypes, int* range_first, int* n_ranges);
int scpr_multi_decompress_clip(scpr_multi* m, const uint8_t* stream, const uint32_t* sizes, const uint8_t* ftypes, int n, uint8_t* frames, int pitch);

/* ---- host layer above the codec object: VfW policy and the AVI container (csrc/vfw_host.cpp) -------------------
 * What CodecInst does around ScreenCodec (screenpressor.cpp:343-437, 579-620), for hosts that are not VfW. */
typedef struct scpr_policy {
    int force_interval;   /* 1: a keyframe every kf_interval frames, host requests ignored (Configuration::ForceInterval, default 1) */
    int kf_interval;      /* default 500 (conf.h:7) */
    int force_loss;       /* 1: always conf_loss; 0: derived from the host's quality value (default 1) */
    int conf_loss;        /* bits of loss, default 0 */
} scpr_policy;
typedef struct scpr_session scpr_session;
void scpr_policy_default(scpr_policy* p);
int scpr_quality_to_loss(uint32_t quality);                          /* 0..10000 -> 4..0 bits (screenpressor.cpp:410-422) */
int scpr_infer_frame_type(uint8_t first_byte, uint32_t data_size);   /* 0 = I, 1 = P, -1 = cannot tell (screenpressor.cpp:579-589) */
int scpr_session_create(const scpr_params* p, int device, const scpr_policy* policy /* NULL = defaults */, scpr_session** out);
void scpr_session_destroy(scpr_session* s);
/* CodecInst::Compress: frame type from the policy (and host_keyframe = ICCOMPRESS_KEYFRAME when the interval is not forced),
 * loss from the policy or quality, then scpr_compress_frame.  *is_key = the AVIIF_KEYFRAME flag of the chunk. */
int scpr_session_compress(scpr_session* s, const uint8_t* src, uint8_t* dst, int dst_cap, int host_keyframe, uint32_t quality, int* is_key);
/* CodecInst::Decompress: not_keyframe = ICDECOMPRESS_NOTKEYFRAME; overridden by what the data itself says. */
int scpr_session_decompress(scpr_session* s, const uint8_t* src, int src_len, uint8_t* dst, int pitch, int not_keyframe);

/* RIFF AVI (1.0, < 4 GB) with one video stream, handler / biCompression 'SCPR' (screenpressor.h:6), '00dc' chunks and an
 * 'idx1' index carrying AVIIF_KEYFRAME: the files VfW hosts write with the reference codec, and read with it. */
typedef struct scpr_avi_info {
    uint32_t width, height, bits_per_pixel;      /* of the uncompressed frames (16 / 24 / 32) */
    uint32_t redmask, greenmask, bluemask;       /* 16 bpp: stored after the BITMAPINFOHEADER (screenpressor.cpp:318-336) */
    uint32_t fps_num, fps_den;
    uint32_t frames;                             /* reader: number of video chunks */
    uint32_t fourcc;                             /* reader: biCompression of the stream ('SCPR' = 0x52504353) */
} scpr_avi_info;
typedef struct scpr_avi scpr_avi;
int scpr_avi_create(const char* path, const scpr_avi_info* info, scpr_avi** out);
int scpr_avi_write_frame(scpr_avi* a, const uint8_t* data, uint32_t len, int is_key);
int scpr_avi_open(const char* path, scpr_avi** out, scpr_avi_info* info);
/* buf = NULL: returns the chunk size.  *is_key from the index (0 when the file has none). */
int64_t scpr_avi_read_frame(scpr_avi* a, uint32_t i, uint8_t* buf, size_t cap, int* is_key);
int scpr_avi_close(scpr_avi* a);   /* writer: appends the index and patches the headers */

/* ---- plumbing ------------------------------------------------------------------------------ */
/* CUDA stream (cudaStream_t) all kernels of this codec are launched on; default: the legacy stream. */
int scpr_set_stream(scpr_codec* c, void* cuda_stream);
/* Text of the last error on this thread. */
const char* scpr_last_error(void);
/* Number of kernels this codec has launched so far (bench.py reports it as gpu_launches). */
uint64_t scpr_kernel_launches(const scpr_codec* c);
/* W*H*6, the output capacity the VfW layer promises (CompressGetSize, screenpressor.cpp:386-388). */
size_t scpr_max_compressed_size(const scpr_params* p);

/* ---- stage hooks for parity tests and profiling (not needed by a codec user) -------------- */
/* Events ((context id << 16) | symbol, DESIGN.md "event format") and intervals
 * ((cum << 16) | freq; freq 0 = raw byte) of frame `i` of the most recent compress call.
 * Copies up to cap entries to host memory; returns the frame's event count. */
int64_t scpr_debug_events(scpr_codec* c, int frame, uint32_t* ev, uint32_t* iv, size_t cap);
/* Stage-A result of frame `i` of the most recent compress call: per 16x16 block, block type
 * (bts, 0..4), changed sub-rect {x1,y1,x2,y2} and motion vector {mx,my}.  nb = nbx*nby entries. */
int scpr_debug_blocks(scpr_codec* c, int frame, uint8_t* bts, int32_t* sxy4, int32_t* mv2);
/* Runs only the frame-scan kernel (differencing + changed-block detection + flat test) over n
 * device-resident frames `reps` times; returns the mean kernel time in milliseconds measured with
 * CUDA events on the codec's stream (bench.py's roofline leg), or < 0. */
float scpr_bench_frame_scan(scpr_codec* c, const uint8_t* d_frames, int n, int reps);
/* The same with the kernel chosen by the caller: mode 0 = what the encoder uses (the TMA tile stream, csrc/frame_scan_tma.cu, for
 * 32 bpp frames at least 128 x 16 pixels), 1 = the plain-load kernels (csrc/frame_scan.cu), 2 = TMA or SCPR_E_UNSUPPORTED.
 * d_prev: device frame that precedes frame 0 (NULL = the codec's own previous frame).  When blkinfo / summary are not NULL the
 * outputs of the last run are copied to the host: n * nb block words (bit 0 changed, bit 1 partial, 4 x 4 bits sub-rect) and
 * n x {notflat, changed, pixel0, pad} u32 -- the tests compare the kernels with each other on ragged geometries. */
float scpr_debug_frame_scan(scpr_codec* c, int mode, const uint8_t* d_frames, const uint8_t* d_prev, int n, int reps, uint32_t* blkinfo, uint32_t* summary);

#ifdef __cplusplus
}
#endif
#endif
