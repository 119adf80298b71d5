// scpr_dev.cuh -- shared device-side definitions of the B200 ScreenPressor hot path.
//
// Data layout in HBM:
//   frames      native input pixels, n frames back to back; 32 bpp: pitch = 4*X, one u32 per pixel
//               (B,G,R,A in memory order -> value & 0x00FFFFFF is the RGB triple the codec sees).
//   blkinfo     one u32 per 16x16 block per frame (frame scan output, see BI_* below).
//   events      one u32 per coded symbol in bitstream order: (context id << 16) | symbol.
//   intervals   one u32 per event: (cum << 16) | freq, freq == 0 -> raw byte `cum` (ransmt.h:125-128).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace scpr {

// ---- context ids (identical to oracle/scpr_oracle.h so event lists can be diffed) ----------
constexpr int CX_COLOR = 0;      // 0..12287: cntab[ch][cx]      (reference screencap.h:436)
constexpr int CX_NTAB = 12288;   // +ptype: run lengths, 256     (screencap.h:437)
constexpr int CX_NTAB2 = 12294;  // block-type run lengths, 256  (screencap.h:438)
constexpr int CX_XX = 12295;     // changed-range bytes, 256     (screencap.h:443)
constexpr int CX_BT = 12296;     // block types, 5               (screencap.h:439)
constexpr int CX_SXY = 12297;    // +0..3 sub-rect paddings, 16  (screencap.h:440)
constexpr int CX_MV = 12301;     // +0..1 motion vectors, 512    (screencap.h:441)
constexpr int CX_PTYPE = 12303;  // +lastptype: pixel types, 6   (screencap.h:442)
constexpr int CX_BOOL = 12309;   // fixed p = 1/2 flag           (screencap.h:407-410)
constexpr int NUM_CX = 12310;
constexpr int NUM_COLOR_CX = 12288;
constexpr int NUM_FIXED_CX = CX_BOOL - CX_NTAB;  // 21

constexpr int PROB_BITS = 12;
constexpr int PROB_SCALE = 4096;
constexpr uint32_t RANS_L = 1u << 23;
constexpr int RANS_BLOCK = 131072;  // ransmt.h:38

__host__ __device__ inline int fixed_nsym(int id) {
    if (id < CX_BT) return 256;
    if (id == CX_BT) return 5;
    if (id < CX_MV) return 16;
    if (id < CX_PTYPE) return 512;
    return 6;
}

__host__ __device__ inline uint32_t make_ev(int ctx, int sym) { return ((uint32_t)ctx << 16) | (uint32_t)sym; }
__host__ __device__ inline uint32_t make_iv(uint32_t freq, uint32_t cum) { return (cum << 16) | freq; }

// ---- blkinfo word ----------------------------------------------------------------------------
// bit 0      block differs from the previous frame
// bit 1      changed sub-rect is smaller than the block ("partial", reference bts 2/4)
// bits 4-7   sx1 - bx*16      bits 8-11  sy1 - by*16
// bits 12-15 sx2-1 - bx*16    bits 16-19 sy2-1 - by*16      (exact bbox of differing pixels,
//                                                            reference screencap.cpp:985-1039)
constexpr uint32_t BI_CHANGED = 1u, BI_PARTIAL = 2u;
__host__ __device__ inline uint32_t bi_pack(int sx1, int sy1, int sx2m1, int sy2m1, bool partial) {
    return BI_CHANGED | (partial ? BI_PARTIAL : 0u) | ((uint32_t)sx1 << 4) | ((uint32_t)sy1 << 8) |
           ((uint32_t)sx2m1 << 12) | ((uint32_t)sy2m1 << 16);
}

// per-frame summary written by the frame scan
struct FrameSummary {
    uint32_t notflat;   // some pixel differs from pixel 0 (IsFlat, screencap.cpp:1436-1444)
    uint32_t changed;   // some pixel differs from the previous frame (CMD_CMPPREV, screencap.cpp:845-851)
    uint32_t pixel0;    // RGB of pixel 0
    uint32_t pad;
};

// frame geometry shared by all kernels
struct Geo {
    int X, Y;          // pixels
    int bpp;           // bytes per pixel of the resident frames (3 or 4)
    int pitch;         // bytes per row of the resident frames
    int nbx, nby, nb;  // 16x16 block grid
    size_t frame_bytes;
};

// RGB triple of pixel (x, y) as 0x00RRGGBB-style u32 (memory bytes 0,1,2 in bits 0-7, 8-15, 16-23)
__device__ __forceinline__ uint32_t load_px(const uint8_t* f, const Geo& g, int x, int y) {
    const uint8_t* p = f + (uint32_t)(y * g.pitch + x * g.bpp);  // a frame is < 4 GB (<= 65536 blocks)
    if (g.bpp == 4) return *reinterpret_cast<const uint32_t*>(p) & 0x00FFFFFFu;
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
}

// Byte `o` of the reference's padded RGB24 image (stride (3X+3)&~3, padding bytes are zero,
// screencap.cpp:215-219); o may address padding.  Used for the I-frame "top-left" neighbour at
// x == 0, which reaches back into the tail of row y-2.
__device__ __forceinline__ uint32_t rgb24_triple_at(const uint8_t* f, const Geo& g, int stride24, long o) {
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        long ok = o + k;
        int row = (int)(ok / stride24), col = (int)(ok % stride24);
        uint32_t b = 0;
        if (col < 3 * g.X) b = f[(size_t)row * g.pitch + (size_t)(col / 3) * g.bpp + (col % 3)];
        v |= b << (8 * k);
    }
    return v;
}

#define SCPR_CUDA_CHECK(call)                                                                  \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            scpr::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return SCPR_E_CUDA;                                                                \
        }                                                                                      \
    } while (0)

void set_error(const char* fmt, ...);

}  // namespace scpr
