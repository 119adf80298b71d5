// frame_scan.cu -- stage A, pass 1: frame differencing, changed-block detection and flat test.
//
// Replaces, fused into one streaming pass over the current and previous frame in their native
// pixel format (no RGB32->RGB24 repack, no separate memcmp, no memcpy(prev)):
//   ScreenCodec::CompressFrame's repack loop   reference screencap.cpp:1652-1664
//   IsFlat                                     screencap.cpp:1436-1444
//   CMD_CMPPREV whole-frame compare            screencap.cpp:845-851
//   DecideBlockTypes' change test and exact    screencap.cpp:985-1039
//   bounding box of the changed pixels
//
// HBM roofline kernel: algorithmic bytes = W*H*(bpp+bpp) read per frame; output is
// one u32 per 16x16 block.  One warp owns a strip of 8 blocks (128 pixels x 16 rows): each lane
// streams 4 pixels (one 128-bit load) of 16 rows from both frames, 16 loads in flight per half
// strip; the 4 lanes of a block merge their row/column difference masks with shuffles, so an
// unchanged block costs no further work anywhere in the pipeline (its blkinfo word is 0).
#include <stdlib.h>
#include <string.h>

#include "kernels.cuh"

namespace scpr {

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// 32 bpp, X % 4 == 0.  grid.x * 8 warps cover n * nby * ceil(nbx/8) strips.
__global__ void __launch_bounds__(256) k_frame_scan32(const uint8_t* __restrict__ frames, const uint8_t* __restrict__ prev0,
                                                      int n, Geo g, uint32_t* __restrict__ blkinfo,
                                                      FrameSummary* __restrict__ summary) {
    const int lane = threadIdx.x & 31;
    const int strips_x = (g.nbx + 7) >> 3;
    const int strips_per_frame = strips_x * g.nby;
    const long strip = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (strip >= (long)n * strips_per_frame) return;
    const int f = (int)(strip / strips_per_frame);
    const int s = (int)(strip - (long)f * strips_per_frame);
    const int by = s / strips_x, sx = s - by * strips_x;

    const uint8_t* cur = frames + (size_t)f * g.frame_bytes;
    const uint8_t* prv = f > 0 ? cur - g.frame_bytes : prev0;
    const uint32_t pixel0 = *reinterpret_cast<const uint32_t*>(cur) & 0x00FFFFFFu;

    const int x0 = sx * 128 + lane * 4;
    const bool lane_ok = x0 < g.X;
    const int y0 = by * 16;
    const int rows = min(16, g.Y - y0);

    uint32_t rowmask = 0, colmask = 0, flatdiff = 0;
    if (lane_ok) {
        const size_t base = (size_t)y0 * g.pitch + (size_t)x0 * 4;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            uint4 c[8], p[8];
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int rr = half * 8 + r;
                if (rr < rows) {
                    c[r] = ld_stream(reinterpret_cast<const uint4*>(cur + base + (size_t)rr * g.pitch));
                    p[r] = ld_stream(reinterpret_cast<const uint4*>(prv + base + (size_t)rr * g.pitch));
                }
            }
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const int rr = half * 8 + r;
                if (rr < rows) {
                    const uint32_t d0 = (c[r].x ^ p[r].x) & 0x00FFFFFFu, d1 = (c[r].y ^ p[r].y) & 0x00FFFFFFu;
                    const uint32_t d2 = (c[r].z ^ p[r].z) & 0x00FFFFFFu, d3 = (c[r].w ^ p[r].w) & 0x00FFFFFFu;
                    const uint32_t m = (d0 ? 1u : 0u) | (d1 ? 2u : 0u) | (d2 ? 4u : 0u) | (d3 ? 8u : 0u);
                    colmask |= m;
                    rowmask |= (m ? 1u : 0u) << rr;
                    flatdiff |= ((c[r].x ^ pixel0) | (c[r].y ^ pixel0) | (c[r].z ^ pixel0) | (c[r].w ^ pixel0)) & 0x00FFFFFFu;
                }
            }
        }
    }
    // merge the 4 lanes of each 16-pixel block
    uint32_t col16 = colmask << (4 * (lane & 3));
    col16 |= __shfl_xor_sync(0xFFFFFFFFu, col16, 1);
    col16 |= __shfl_xor_sync(0xFFFFFFFFu, col16, 2);
    rowmask |= __shfl_xor_sync(0xFFFFFFFFu, rowmask, 1);
    rowmask |= __shfl_xor_sync(0xFFFFFFFFu, rowmask, 2);

    const int bx = sx * 8 + (lane >> 2);
    if ((lane & 3) == 0 && bx < g.nbx) {
        uint32_t info = 0;
        if (col16) {
            const int sx1 = __ffs(col16) - 1, sx2m1 = 31 - __clz(col16);
            const int sy1 = __ffs(rowmask) - 1, sy2m1 = 31 - __clz(rowmask);
            const int bw = min(16, g.X - bx * 16);
            const bool partial = sx1 > 0 || sy1 > 0 || sx2m1 < bw - 1 || sy2m1 < rows - 1;
            info = bi_pack(sx1, sy1, sx2m1, sy2m1, partial);
        }
        blkinfo[(size_t)f * g.nb + (size_t)by * g.nbx + bx] = info;
    }
    // frame-level flags: benign races, every writer stores the same value
    if (__any_sync(0xFFFFFFFFu, col16 != 0) && lane == 0) summary[f].changed = 1;
    if (__any_sync(0xFFFFFFFFu, flatdiff != 0) && lane == 0) summary[f].notflat = 1;
    if (s == 0 && lane == 0) summary[f].pixel0 = pixel0;
}

// Any pixel format / width: one warp per 16x16 block, lane handles 8 pixels.
__global__ void __launch_bounds__(256) k_frame_scan_generic(const uint8_t* __restrict__ frames, const uint8_t* __restrict__ prev0,
                                                            int n, Geo g, uint32_t* __restrict__ blkinfo,
                                                            FrameSummary* __restrict__ summary) {
    const int lane = threadIdx.x & 31;
    const long blk = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (blk >= (long)n * g.nb) return;
    const int f = (int)(blk / g.nb);
    const int bi = (int)(blk - (long)f * g.nb);
    const int by = bi / g.nbx, bx = bi - by * g.nbx;
    const uint8_t* cur = frames + (size_t)f * g.frame_bytes;
    const uint8_t* prv = f > 0 ? cur - g.frame_bytes : prev0;
    const uint32_t pixel0 = load_px(cur, g, 0, 0);
    const int bw = min(16, g.X - bx * 16), bh = min(16, g.Y - by * 16);
    uint32_t rowmask = 0, colmask = 0, flatdiff = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int p = lane + 32 * k, px = p & 15, py = p >> 4;
        if (px < bw && py < bh) {
            const uint32_t c = load_px(cur, g, bx * 16 + px, by * 16 + py);
            const uint32_t q = load_px(prv, g, bx * 16 + px, by * 16 + py);
            if (c != q) {
                colmask |= 1u << px;
                rowmask |= 1u << py;
            }
            flatdiff |= c ^ pixel0;
        }
    }
    colmask = __reduce_or_sync(0xFFFFFFFFu, colmask);
    rowmask = __reduce_or_sync(0xFFFFFFFFu, rowmask);
    flatdiff = __reduce_or_sync(0xFFFFFFFFu, flatdiff);
    if (lane == 0) {
        uint32_t info = 0;
        if (colmask) {
            const int sx1 = __ffs(colmask) - 1, sx2m1 = 31 - __clz(colmask);
            const int sy1 = __ffs(rowmask) - 1, sy2m1 = 31 - __clz(rowmask);
            const bool partial = sx1 > 0 || sy1 > 0 || sx2m1 < bw - 1 || sy2m1 < bh - 1;
            info = bi_pack(sx1, sy1, sx2m1, sy2m1, partial);
            summary[f].changed = 1;
        }
        blkinfo[blk] = info;
        if (flatdiff) summary[f].notflat = 1;
        if (bi == 0) summary[f].pixel0 = pixel0;
    }
}

// 24 bpp with X % 4 != 0: IsFlat also compares the caller's row padding (screencap.cpp:1439-1440).
__global__ void k_flat_padding24(const uint8_t* __restrict__ frames, int n, Geo g, FrameSummary* __restrict__ summary) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long)n * g.Y) return;
    const int f = (int)(t / g.Y), y = (int)(t - (long)f * g.Y);
    const uint8_t* cur = frames + (size_t)f * g.frame_bytes;
    for (int k = 3 * g.X; k < g.pitch; k++)
        if (cur[(size_t)y * g.pitch + k] != cur[k]) summary[f].notflat = 1;
}

// Ordered list of changed blocks per P frame (raster order) + bounding box of changed blocks
// (xx1, xx2 of CompressP, screencap.cpp:1132-1150).  One CTA per frame.
__global__ void __launch_bounds__(256) k_compact_changed(const uint32_t* __restrict__ blkinfo, const uint8_t* __restrict__ ftype,
                                                         Geo g, uint32_t* __restrict__ chg_list, PFrameHdr* __restrict__ hdr) {
    const int f = blockIdx.x;
    if (ftype[f] != FT_P) return;
    __shared__ int warp_cnt[8];
    __shared__ int base;
    __shared__ int bb[4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        base = 0;
        bb[0] = g.nbx; bb[1] = g.nby; bb[2] = -1; bb[3] = -1;
    }
    __syncthreads();
    int bx1 = g.nbx, by1 = g.nby, bx2 = -1, by2 = -1;
    for (int b0 = 0; b0 < g.nb; b0 += 256) {
        const int bi = b0 + threadIdx.x;
        const bool ch = bi < g.nb && (blkinfo[(size_t)f * g.nb + bi] & BI_CHANGED);
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, ch);
        if (lane == 0) warp_cnt[w] = __popc(m);
        __syncthreads();
        int off = base;
        for (int k = 0; k < w; k++) off += warp_cnt[k];
        if (ch) {
            chg_list[(size_t)f * g.nb + off + __popc(m & ((1u << lane) - 1))] = bi;
            const int by = bi / g.nbx, bx = bi - by * g.nbx;
            bx1 = min(bx1, bx); bx2 = max(bx2, bx); by1 = min(by1, by); by2 = max(by2, by);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int k = 0; k < 8; k++) t += warp_cnt[k];
            base += t;
        }
        __syncthreads();
    }
    atomicMin(&bb[0], bx1); atomicMin(&bb[1], by1); atomicMax(&bb[2], bx2); atomicMax(&bb[3], by2);
    __syncthreads();
    if (threadIdx.x == 0) {
        hdr[f].n_changed = base;
        hdr[f].xx1 = bb[1] * g.nbx + bb[0];
        hdr[f].xx2 = bb[3] * g.nbx + bb[2];
    }
}

// Lossy modes (SetupLossMask / CMD_DOLOSS, screencap.cpp:127-139, 852-861): every colour byte of a
// non-flat frame becomes (b & ~(2^loss - 1)) | (2^loss >> 1), in place, before differencing; the masked
// frame is what the next frame is compared against.  Flat frames are left untouched (the reference
// stores them in prev before DoLoss runs, screencap.cpp:1488-1499).
__global__ void __launch_bounds__(256) k_apply_loss(uint8_t* __restrict__ frames, int n, Geo g, const FrameSummary* __restrict__ summary,
                                                    uint32_t keep, uint32_t corr) {
    const size_t words_per_frame = g.frame_bytes / 4;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * words_per_frame) return;
    const int f = (int)(i / words_per_frame);
    if (!summary[f].notflat) return;
    uint32_t* p = reinterpret_cast<uint32_t*>(frames) + i;
    *p = (*p & keep) | corr;
}

void launch_apply_loss(uint8_t* frames, int n, const Geo& g, const FrameSummary* summary, int loss, cudaStream_t st, uint64_t* launches) {
    uint32_t m = (1u << loss) - 1;
    m |= m << 8;
    m |= m << 16;
    uint32_t cm = (1u << loss) >> 1;
    cm |= cm << 8;
    cm |= cm << 16;
    const size_t words = (size_t)n * (g.frame_bytes / 4);
    uint32_t keep = ~m;
    if (g.bpp == 4) {  // the alpha byte is not part of the picture: leave it alone
        keep |= 0xFF000000u;
        cm &= 0x00FFFFFFu;
    }
    k_apply_loss<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(frames, n, g, summary, keep, cm);
    ++*launches;
}

// mode: 0 = pick (TMA tile stream when the geometry allows, unless SCPR_FRAME_SCAN=ld), 1 = plain-load kernels, 2 = TMA or fail.
bool launch_frame_scan_mode(int mode, const uint8_t* frames, const uint8_t* prev0, int n, const Geo& g, uint32_t* blkinfo,
                            FrameSummary* summary, cudaStream_t st, uint64_t* launches) {
    if (g.bpp == 4 && (g.X & 3) == 0) {
        static const bool force_ld = []() { const char* e = getenv("SCPR_FRAME_SCAN"); return e && !strcmp(e, "ld"); }();
        if (mode == 2 || (mode == 0 && !force_ld)) {
            if (frame_scan_tma_usable(frames, prev0, g) && launch_frame_scan_tma(frames, prev0, n, g, blkinfo, summary, st, launches)) return true;
            if (mode == 2) return false;
        }
        const long strips = (long)n * g.nby * ((g.nbx + 7) >> 3);
        k_frame_scan32<<<(unsigned)((strips + 7) / 8), 256, 0, st>>>(frames, prev0, n, g, blkinfo, summary);
        ++*launches;
    } else {
        if (mode == 2) return false;
        const long blocks = (long)n * g.nb;
        k_frame_scan_generic<<<(unsigned)((blocks + 7) / 8), 256, 0, st>>>(frames, prev0, n, g, blkinfo, summary);
        ++*launches;
        if (g.bpp == 3 && (g.X & 3)) {
            const long rows = (long)n * g.Y;
            k_flat_padding24<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(frames, n, g, summary);
            ++*launches;
        }
    }
    return true;
}
void launch_frame_scan(const uint8_t* frames, const uint8_t* prev0, int n, const Geo& g, uint32_t* blkinfo,
                       FrameSummary* summary, cudaStream_t st, uint64_t* launches) {
    launch_frame_scan_mode(0, frames, prev0, n, g, blkinfo, summary, st, launches);
}

// ---- 16 bpp input / output (ScreenCodec::CompressFrame, screencap.cpp:1665-1678; DecompressFrame, :1726-1734) --------
// The codec proper only knows RGB24.  A 16-bit pixel is split with the caller's channel masks into three *unscaled*
// bytes (5- or 6-bit values), rows of the RGB24 image are padded with zeros to a multiple of four bytes; decoding puts
// the three bytes back with the same shifts.  One thread per pixel pair keeps the 16-bit side 32-bit coalesced.
__global__ void k_unpack16(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n, Geo g, Rgb16 m) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int per_row = g.pitch;  // one thread per output byte column group of 3 (pixel) + padding bytes
    const long rows = (long)n * g.Y;
    const long row = t / per_row;
    if (row >= rows) return;
    const int col = (int)(t - row * per_row);
    const int f = (int)(row / g.Y), y = (int)(row - (long)f * g.Y);
    uint8_t v = 0;
    if (col < 3 * g.X) {
        const int x = col / 3, ch = col - 3 * x;
        const uint16_t w = *reinterpret_cast<const uint16_t*>(src + ((size_t)f * g.Y + y) * (size_t)(2 * g.X) + 2 * (size_t)x);
        v = ch == 0 ? (uint8_t)((w & m.rmask) >> m.rshift) : ch == 1 ? (uint8_t)((w & m.gmask) >> m.gshift) : (uint8_t)((w & m.bmask) >> m.bshift);
    }
    dst[(size_t)f * g.frame_bytes + (size_t)y * g.pitch + col] = v;
}
__global__ void k_pack16(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int n, Geo g, int out_pitch, Rgb16 m) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long px = (long)n * g.Y * g.X;
    if (t >= px) return;
    const int x = (int)(t % g.X);
    const long row = t / g.X;
    const int f = (int)(row / g.Y), y = (int)(row - (long)f * g.Y);
    const uint8_t* p = src + (size_t)f * g.frame_bytes + (size_t)y * g.pitch + 3 * (size_t)x;
    const uint32_t w = ((uint32_t)p[0] << m.rshift) + ((uint32_t)p[1] << m.gshift) + ((uint32_t)p[2] << m.bshift);
    *reinterpret_cast<uint16_t*>(dst + ((size_t)f * g.Y + y) * (size_t)out_pitch + 2 * (size_t)x) = (uint16_t)w;
}
void launch_unpack16(const uint8_t* src16, uint8_t* dst24, int n, const Geo& g, const Rgb16& m, cudaStream_t st, uint64_t* launches) {
    const long total = (long)n * g.Y * g.pitch;
    k_unpack16<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src16, dst24, n, g, m);
    ++*launches;
}
void launch_pack16(const uint8_t* src24, uint8_t* dst16, int n, const Geo& g, int out_pitch, const Rgb16& m, cudaStream_t st, uint64_t* launches) {
    const long total = (long)n * g.Y * g.X;
    k_pack16<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src24, dst16, n, g, out_pitch, m);
    ++*launches;
}

void launch_compact_changed(const uint32_t* blkinfo, const uint8_t* ftype, int n, const Geo& g, uint32_t* chg_list,
                            PFrameHdr* hdr, cudaStream_t st, uint64_t* launches) {
    k_compact_changed<<<n, 256, 0, st>>>(blkinfo, ftype, g, chg_list, hdr);
    ++*launches;
}

}  // namespace scpr
