// frame_scan_tma.cu -- stage A, pass 1 for 32 bpp frames: the frame scan as a persistent, TMA-staged tile stream.
//
// Same outputs as k_frame_scan32 (frame_scan.cu): blkinfo words and the per-frame summary; same reference code replaced
// (repack loop screencap.cpp:1652-1664, IsFlat :1436-1444, CMD_CMPPREV :845-851, DecideBlockTypes' change test and exact
// bounding box :985-1039).  What changes is how the pixels travel:
//
//   * A tile is 128 pixels x 16 rows (8 blocks) of ONE frame = 8 KB, fetched by one cp.async.bulk.tensor.3d (tensor map over
//     {x, y, frame} of u32 pixels, box {128, 16, 1}; rows / columns outside the frame arrive as zeros) into a shared-memory ring,
//     completion signalled on an mbarrier.  No register staging, no address arithmetic per row, one instruction per 8 KB.
//   * A warp owns a tile POSITION and walks it through a run of consecutive frames: the tile of frame f stays in registers
//     (16 x 128 bit per lane) and is the "previous frame" for frame f + 1.  Every pixel therefore crosses HBM -> L2 -> SM once
//     per run instead of twice (k_frame_scan32 reads cur and prev and relies on L2 for the second read): algorithmic bytes stay
//     8 B/pixel, real traffic is 4 B/pixel * (1 + 1/run).
//   * Each warp keeps its own ring of STAGES tiles in flight (its lane 0 issues the next copy as soon as the warp has lifted a
//     tile into registers), so a CTA of WARPS warps has WARPS * STAGES * 8 KB outstanding -- sized to cover HBM latency at full
//     bandwidth with one persistent CTA per SM (148 * 192 KB = 28 MB in flight).
//   * Frame-level flags are collected per run in two 32-bit masks and stored once per run (a run is <= 32 frames), instead of one
//     store per tile on the same few cache lines.
#include <cuda.h>

#include "kernels.cuh"

namespace scpr {

namespace {

constexpr int TILE_W = 128, TILE_H = 16;
constexpr uint32_t TILE_BYTES = TILE_W * TILE_H * 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // (A spin limit with a trap in this loop was measured at +5 % on the whole kernel, 0.84 against 0.80 ms per 600 frames, and is
    // not needed: a copy with a bad descriptor or address faults the kernel, it does not stay silent.)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_tile(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(dst),
        "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
    return r;
}

struct ScanPlan {
    int n;               // frames
    int run;             // frames per run (<= 32)
    int runs;            // ceil(n / run)
    int strips_x;        // tiles per tile row
    int tiles;           // tile positions per frame
    long units;          // tiles * runs
};

// position in a warp's sequence of tile loads: unit u (tile position x run of frames), load k of the unit
// (k = 0: the frame before the run, k = 1..cnt: the frames of the run)
struct Cursor {
    long u;
    int k, cnt, f0, tx, ty;
    __device__ __forceinline__ void set(long unit, const ScanPlan& pl) {
        u = unit;
        k = 0;
        if (u < pl.units) {
            const int r = (int)(u / pl.tiles), t = (int)(u - (long)r * pl.tiles);
            ty = t / pl.strips_x;
            tx = t - ty * pl.strips_x;
            f0 = r * pl.run;
            cnt = min(pl.run, pl.n - f0);
        }
    }
    __device__ __forceinline__ void next(const ScanPlan& pl, long stride) {
        if (++k > cnt) set(u + stride, pl);
    }
};

template <int WARPS, int STAGES>
__global__ void __launch_bounds__(WARPS * 32, 1)
k_frame_scan_tma(const __grid_constant__ CUtensorMap tm_frames, const __grid_constant__ CUtensorMap tm_prev0, const uint8_t* __restrict__ frames,
                 ScanPlan pl, Geo g, uint32_t* __restrict__ blkinfo, FrameSummary* __restrict__ summary) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint8_t* ring = smem + (size_t)w * STAGES * TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * STAGES * TILE_BYTES) + w * STAGES;
    const uint32_t ring_s = smem_u32(ring), bar_s = smem_u32(bars);
    if (lane == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(bar_s + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));

    const long stride = (long)gridDim.x * WARPS;
    const long first = (long)blockIdx.x * WARPS + w;
    Cursor iss, con;
    iss.set(first, pl);
    con.set(first, pl);

    auto issue = [&](const Cursor& c, int stage) {
        if (lane == 0) {
            const uint32_t bar = bar_s + 8 * stage;
            mbar_expect_tx(bar, TILE_BYTES);
            const int f = c.f0 + c.k - 1;
            if (f < 0)
                tma_load_tile(ring_s + stage * TILE_BYTES, &tm_prev0, c.tx * TILE_W, c.ty * TILE_H, 0, bar, policy);
            else
                tma_load_tile(ring_s + stage * TILE_BYTES, &tm_frames, c.tx * TILE_W, c.ty * TILE_H, f, bar, policy);
        }
    };
#pragma unroll 1
    for (int s = 0; s < STAGES; s++)
        if (iss.u < pl.units) {
            issue(iss, s);
            iss.next(pl, stride);
        }

    uint4 p[TILE_H];  // this position's tile of the previous frame
    uint32_t phases = 0, chg_bits = 0, nf_bits = 0, pixel0_mine = 0;
    int stage = 0;
#pragma unroll 1
    while (con.u < pl.units) {
        if (con.k == 0) {
            // pixel 0 of every frame of the run, one frame per lane (used from k = 1 on; the load overlaps the wait below)
            chg_bits = nf_bits = 0;
            if (lane < con.cnt) pixel0_mine = __ldg(reinterpret_cast<const uint32_t*>(frames + (size_t)(con.f0 + lane) * g.frame_bytes)) & 0x00FFFFFFu;
        }
        mbar_wait(bar_s + 8 * stage, (phases >> stage) & 1u);
        phases ^= 1u << stage;
        uint4 c[TILE_H];
        const uint32_t src = ring_s + stage * TILE_BYTES + lane * 16;
#pragma unroll
        for (int r = 0; r < TILE_H; r++) c[r] = lds128(src + r * (TILE_W * 4));
        __syncwarp();  // every lane has its copy: the slot can be refilled
        if (iss.u < pl.units) {
            issue(iss, stage);
            iss.next(pl, stride);
        }
        if (con.k > 0) {
            const int f = con.f0 + con.k - 1;
            const uint32_t pixel0 = __shfl_sync(0xFFFFFFFFu, pixel0_mine, con.k - 1);
            const int y0 = con.ty * TILE_H, rows = min(TILE_H, g.Y - y0);
            const bool lane_ok = con.tx * TILE_W + lane * 4 < g.X;
            uint32_t rowmask = 0, colmask = 0, flatdiff = 0;
#pragma unroll
            for (int r = 0; r < TILE_H; r++) {
                const uint32_t d0 = (c[r].x ^ p[r].x) & 0x00FFFFFFu, d1 = (c[r].y ^ p[r].y) & 0x00FFFFFFu;
                const uint32_t d2 = (c[r].z ^ p[r].z) & 0x00FFFFFFu, d3 = (c[r].w ^ p[r].w) & 0x00FFFFFFu;
                const uint32_t m = (d0 ? 1u : 0u) | (d1 ? 2u : 0u) | (d2 ? 4u : 0u) | (d3 ? 8u : 0u);
                colmask |= m;
                rowmask |= (m ? 1u : 0u) << r;
                if (r < rows) flatdiff |= ((c[r].x ^ pixel0) | (c[r].y ^ pixel0) | (c[r].z ^ pixel0) | (c[r].w ^ pixel0)) & 0x00FFFFFFu;
            }
            if (!lane_ok) flatdiff = 0;
            // merge the 4 lanes of each 16-pixel block
            uint32_t col16 = colmask << (4 * (lane & 3));
            col16 |= __shfl_xor_sync(0xFFFFFFFFu, col16, 1);
            col16 |= __shfl_xor_sync(0xFFFFFFFFu, col16, 2);
            rowmask |= __shfl_xor_sync(0xFFFFFFFFu, rowmask, 1);
            rowmask |= __shfl_xor_sync(0xFFFFFFFFu, rowmask, 2);
            const int bx = con.tx * 8 + (lane >> 2);
            if ((lane & 3) == 0 && bx < g.nbx) {
                uint32_t info = 0;
                if (col16) {
                    const int sx1 = __ffs(col16) - 1, sx2m1 = 31 - __clz(col16);
                    const int sy1 = __ffs(rowmask) - 1, sy2m1 = 31 - __clz(rowmask);
                    const int bw = min(16, g.X - bx * 16);
                    const bool partial = sx1 > 0 || sy1 > 0 || sx2m1 < bw - 1 || sy2m1 < rows - 1;
                    info = bi_pack(sx1, sy1, sx2m1, sy2m1, partial);
                }
                blkinfo[(size_t)f * g.nb + (size_t)con.ty * g.nbx + bx] = info;
            }
            if (__any_sync(0xFFFFFFFFu, col16 != 0)) chg_bits |= 1u << (con.k - 1);
            if (__any_sync(0xFFFFFFFFu, flatdiff != 0)) nf_bits |= 1u << (con.k - 1);
            if (con.k == con.cnt && lane < con.cnt) {
                // frame-level flags of the whole run: one store per frame and flag (benign races, every writer stores the same value)
                FrameSummary* sm = summary + con.f0 + lane;
                if ((chg_bits >> lane) & 1u) sm->changed = 1;
                if ((nf_bits >> lane) & 1u) sm->notflat = 1;
                if (con.tx == 0 && con.ty == 0) sm->pixel0 = pixel0_mine;
            }
        }
#pragma unroll
        for (int r = 0; r < TILE_H; r++) p[r] = c[r];
        con.next(pl, stride);
        stage = stage + 1 == STAGES ? 0 : stage + 1;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

bool make_map(CUtensorMap* m, const uint8_t* base, int n, const Geo& g) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)g.X, (cuuint64_t)g.Y, (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)g.pitch, (cuuint64_t)g.frame_bytes};
    const cuuint32_t box[3] = {TILE_W, TILE_H, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

constexpr int SCAN_WARPS = 6, SCAN_STAGES = 4;

}  // namespace

// true when the TMA tile stream can take this geometry (else the caller uses k_frame_scan32 / the generic kernel)
bool frame_scan_tma_usable(const uint8_t* frames, const uint8_t* prev0, const Geo& g) {
    return g.bpp == 4 && (g.X & 3) == 0 && g.X >= TILE_W && g.Y >= TILE_H && ((uintptr_t)frames & 15) == 0 && ((uintptr_t)prev0 & 15) == 0 &&
           (g.frame_bytes & 15) == 0 && encode_tiled() != nullptr;
}

bool launch_frame_scan_tma(const uint8_t* frames, const uint8_t* prev0, int n, const Geo& g, uint32_t* blkinfo, FrameSummary* summary,
                           cudaStream_t st, uint64_t* launches) {
    CUtensorMap tm_frames, tm_prev0;
    if (!make_map(&tm_frames, frames, n, g) || !make_map(&tm_prev0, prev0, 1, g)) return false;
    static int sm_count[64] = {0};  // per device ordinal; 0 = the kernel's shared-memory limit has not been raised there yet
    constexpr size_t smem = (size_t)SCAN_WARPS * SCAN_STAGES * TILE_BYTES + SCAN_WARPS * SCAN_STAGES * 8;
    auto kern = k_frame_scan_tma<SCAN_WARPS, SCAN_STAGES>;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
    if (!sm_count[dev]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return false;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
        sm_count[dev] = v;
    }
    const int sms = sm_count[dev];
    ScanPlan pl;
    pl.n = n;
    pl.strips_x = (g.X + TILE_W - 1) / TILE_W;
    pl.tiles = pl.strips_x * ((g.Y + TILE_H - 1) / TILE_H);
    const long gw = (long)sms * SCAN_WARPS;
    long run = ((long)n * pl.tiles) / (8 * gw);  // at least ~8 units per warp, runs as long as that allows
    pl.run = (int)(run < 1 ? 1 : run > 32 ? 32 : run);
    pl.runs = (n + pl.run - 1) / pl.run;
    pl.units = (long)pl.tiles * pl.runs;
    const long ctas = (pl.units + SCAN_WARPS - 1) / SCAN_WARPS;
    kern<<<(unsigned)(ctas < sms ? ctas : sms), SCAN_WARPS * 32, smem, st>>>(tm_frames, tm_prev0, frames, pl, g, blkinfo, summary);
    ++*launches;
    return true;
}

}  // namespace scpr
