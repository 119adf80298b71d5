// pframe.cu -- stage A for P frames: motion search, in-order MV resolve, pixel typing + run-length
// coding of the changed sub-rects, and event emission.
//
// Replaces (reference, 1-thread canonical order):
//   FindMV / SameBlocks            screencap.cpp:684-825
//   DecideBlockTypes (typing part) screencap.cpp:1041-1069
//   GetPixelTypeP/P0, PixelTypeFitsP/P0  screencap.cpp:525-556, 578-604
//   CompressP's serialisation      screencap.cpp:1144-1248, WritePixel :609-627
//
// The reference's motion search is serial through `last_mv` and the persistent mvs[] array
// (SURVEY.md A.3).  Here F(b) = "first hit of the fixed candidate order" is computed for every
// changed block of every frame in parallel (k_mv_search, 32 candidates per warp step, first hit by
// ballot), and only the two state-dependent shortcut candidates are evaluated in frame/raster order
// (k_mv_resolve); a shortcut whose vector equals F(b) needs no compare at all.
#include <stdio.h>

#include "codec.h"

namespace scpr {

struct SubRect {
    int bx, by, x1, y1, x2, y2, w, h;
    bool partial;
};

__device__ __forceinline__ SubRect subrect_of(uint32_t bi, uint32_t info, const Geo& g) {
    SubRect r;
    r.by = (int)bi / g.nbx;
    r.bx = (int)bi - r.by * g.nbx;
    r.x1 = r.bx * 16 + (int)((info >> 4) & 15);
    r.y1 = r.by * 16 + (int)((info >> 8) & 15);
    r.x2 = r.bx * 16 + (int)((info >> 12) & 15) + 1;
    r.y2 = r.by * 16 + (int)((info >> 16) & 15) + 1;
    r.w = r.x2 - r.x1;
    r.h = r.y2 - r.y1;
    r.partial = (info & BI_PARTIAL) != 0;
    return r;
}

// search windows of FindMV (screencap.cpp:691-709)
struct Windows {
    int fx1, fx2, fy1, fy2, rx1, rx2, ry1, ry2;
};
__device__ __forceinline__ Windows windows_of(const SubRect& r, const Geo& g) {
    Windows w;
    w.rx1 = max(0, r.x1 - 8);
    w.ry1 = max(0, r.y1 - 8);
    w.rx2 = r.x1 + 8;
    w.ry2 = r.y1 + 8;
    if (w.rx2 + r.w > g.X) w.rx2 = g.X - r.w + 1;
    if (w.ry2 + r.h > g.Y) w.ry2 = g.Y - r.h + 1;
    w.fx1 = max(0, r.x1 - 256);
    w.fy1 = max(0, r.y1 - 256);
    w.fx2 = r.x1 + 256;
    w.fy2 = r.y1 + 256;
    if (w.fx2 + r.w > g.X) w.fx2 = g.X - r.w + 1;
    if (w.fy2 + r.h > g.Y) w.fy2 = g.Y - r.h + 1;
    return w;
}

__device__ __forceinline__ int find_pframe(const PFrameHdr* hdr, const int* pframes, int n_pframes, int slot) {
    int lo = 0, hi = n_pframes - 1;
    while (lo < hi) {  // largest i with chg_off <= slot
        const int mid = (lo + hi + 1) >> 1;
        if (hdr[pframes[mid]].chg_off <= slot) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ------------------------------------------------------------------------------------------------
// k_mv_search: F(b) for every changed block.  One warp per block.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_mv_search(PWork w) {
    __shared__ uint32_t s_cur[4][256];
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int slot = blockIdx.x * 4 + wi;
    if (slot >= w.total_blocks) return;
    const int pi = find_pframe(w.hdr, w.pframes, w.n_pframes, slot);
    const int f = w.pframes[pi];
    const int k = slot - w.hdr[f].chg_off;
    const Geo& g = w.g;
    const uint32_t bi = w.chg_list[(size_t)f * g.nb + k];
    const uint32_t info = w.blkinfo[(size_t)f * g.nb + bi];
    const SubRect r = subrect_of(bi, info, g);
    const Windows win = windows_of(r, g);
    const uint8_t* cur = w.frames + (size_t)f * g.frame_bytes;
    const uint8_t* prv = f > 0 ? cur - g.frame_bytes : w.prev0;

    const int npx = r.w * r.h;
    for (int p = lane; p < npx; p += 32) s_cur[wi][p] = load_px(cur, g, r.x1 + p % r.w, r.y1 + p / r.w);
    __syncwarp();

    // candidate segments in the reference's order (screencap.cpp:737-811)
    const int common = min(r.y1 - win.fy1, win.fy2 - r.y1 - 1);
    const int nA = 2 * common;
    const int nB = max(0, (r.y1 - 1 - common) - win.fy1 + 1);
    const int nC = max(0, win.fy2 - (r.y1 + 1 + common));
    const int nD = r.x1 - win.fx1 + 1;
    const int nE = win.fx2 - r.x1;
    const int nyu = r.y1 - win.ry1 + 1, nyd = win.ry2 - r.y1 - 1, ny = nyu + nyd;
    const int nxl = r.x1 - win.rx1 + 1, nxr = max(0, win.rx2 - r.x1 - 1);
    const int total = nA + nB + nC + nD + nE + (nxl + nxr) * ny;

    int hit = -1, hx = 0, hy = 0;
    for (int t0 = 0; t0 < total; t0 += 32) {
        int t = t0 + lane;
        bool ok = t < total;
        int cx = r.x1, cy = r.y1;
        if (ok) {
            if (t < nA) {
                const int j = t >> 1;
                cy = (t & 1) ? r.y1 + 1 + j : r.y1 - 1 - j;
            } else if ((t -= nA) < nB) {
                cy = r.y1 - 1 - common - t;
            } else if ((t -= nB) < nC) {
                cy = r.y1 + 1 + common + t;
            } else if ((t -= nC) < nD) {
                cx = r.x1 - t;
            } else if ((t -= nD) < nE) {
                cx = r.x1 + t;
            } else {
                t -= nE;
                const int xi = t / ny, yi = t - xi * ny;
                cx = xi < nxl ? r.x1 - xi : r.x1 + 1 + (xi - nxl);
                cy = yi < nyu ? r.y1 - yi : r.y1 + 1 + (yi - nyu);
            }
            // SameBlocks (screencap.cpp:817-825), early exit on the first differing pixel
            for (int yy = 0; yy < r.h && ok; yy++)
                for (int xx = 0; xx < r.w; xx++)
                    if (load_px(prv, g, cx + xx, cy + yy) != s_cur[wi][yy * r.w + xx]) {
                        ok = false;
                        break;
                    }
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, ok);
        if (m) {
            const int src = __ffs(m) - 1;
            hit = t0 + src;
            hx = __shfl_sync(0xFFFFFFFFu, cx, src);
            hy = __shfl_sync(0xFFFFFFFFu, cy, src);
            break;
        }
    }
    if (lane == 0) {
        ChgBlock& b = w.blocks[slot];
        b.bi = bi;
        b.info = info;
        b.has_f = hit >= 0;
        b.fmx = (int16_t)(hx - r.x1);
        b.fmy = (int16_t)(hy - r.y1);
    }
}

// warp-cooperative SameBlocks: sub-rect of cur at (x1,y1) vs prev at (x1+mx, y1+my); 32 pixels per
// step, stops at the first step that differs
__device__ __forceinline__ bool warp_match(const uint8_t* cur, const uint8_t* prv, const Geo& g, const SubRect& r, int mx,
                                           int my, int lane) {
    const int npx = r.w * r.h;
    for (int p0 = 0; p0 < npx; p0 += 32) {
        const int p = p0 + lane;
        bool same = true;
        if (p < npx) {
            const int xx = p % r.w, yy = p / r.w;
            same = load_px(cur, g, r.x1 + xx, r.y1 + yy) == load_px(prv, g, r.x1 + mx + xx, r.y1 + my + yy);
        }
        if (!__all_sync(0xFFFFFFFFu, same)) return false;
    }
    return true;
}
// the same test for the serial resolve, where the latency of a compare is on the critical path: the first 32 pixels
// decide most mismatches (one round trip, few instructions -- a lone warp pays ~6 cycles per instruction); only when
// they agree are the remaining pixels fetched, all loads in flight at once
__device__ __forceinline__ bool warp_match_full(const uint8_t* cur, const uint8_t* prv, const Geo& g, const SubRect& r, int mx,
                                                int my, int lane) {
    const int npx = r.w * r.h;  // <= 256
    const uint32_t winv = (65536u + (uint32_t)r.w - 1) / (uint32_t)r.w;  // p / w == (p * winv) >> 16 for p < 256
    {
        bool same = true;
        if (lane < npx) {
            const int yy = (int)(((uint32_t)lane * winv) >> 16), xx = lane - yy * r.w;
            same = load_px(cur, g, r.x1 + xx, r.y1 + yy) == load_px(prv, g, r.x1 + mx + xx, r.y1 + my + yy);
        }
        if (!__all_sync(0xFFFFFFFFu, same)) return false;
        if (npx <= 32) return true;
    }
    uint32_t a[7], b[7];
#pragma unroll
    for (int u = 0; u < 7; u++) {
        const int p = lane + 32 * (u + 1);
        a[u] = b[u] = 0;
        if (p < npx) {
            const int yy = (int)(((uint32_t)p * winv) >> 16), xx = p - yy * r.w;
            a[u] = load_px(cur, g, r.x1 + xx, r.y1 + yy);
            b[u] = load_px(prv, g, r.x1 + mx + xx, r.y1 + my + yy);
        }
    }
    bool same = true;
#pragma unroll
    for (int u = 0; u < 7; u++) same = same && a[u] == b[u];
    return __all_sync(0xFFFFFFFFu, same);
}
__device__ __forceinline__ bool in_far_window(const SubRect& r, const Windows& win, int mx, int my) {
    const int sx = r.x1 + mx, sy = r.y1 + my;
    return sx >= win.fx1 && sx < win.fx2 && sy >= win.fy1 && sy < win.fy2;
}

// ------------------------------------------------------------------------------------------------
// k_mv_cands: per P frame, the distinct F vectors of its blocks (first MAXC of them, in order of
// appearance) -- the only values last_mv can take in this frame besides (0,0).  One warp per frame;
// lane k keeps candidate k.  Also records each block's index into that list (fidx, 0xFF = none).
// ------------------------------------------------------------------------------------------------
constexpr int MAXC = MV_MAXC;
static_assert(MAXC == 64, "two candidates per lane");
// a frame's candidate list across the warp: slot k in lane k & 31, `a` for k < 32, `b` for k >= 32
struct Cands {
    int a, b;
    __device__ __forceinline__ int find(int v, int n, int lane) const {  // index of v among the first n slots, -1 if absent
        const uint32_t ha = __ballot_sync(0xFFFFFFFFu, lane < n && a == v);
        if (ha) return __ffs(ha) - 1;
        const uint32_t hb = __ballot_sync(0xFFFFFFFFu, lane + 32 < n && b == v);
        return hb ? 31 + __ffs(hb) : -1;
    }
    __device__ __forceinline__ void put(int v, int n, int lane) {  // slot n := v
        if (n < 32) {
            if (lane == n) a = v;
        } else if (lane == n - 32)
            b = v;
    }
    __device__ __forceinline__ int get(int k) const {  // slot k, broadcast
        const int va = __shfl_sync(0xFFFFFFFFu, a, k & 31), vb = __shfl_sync(0xFFFFFFFFu, b, k & 31);
        return k < 32 ? va : vb;
    }
    __device__ __forceinline__ void load(const int* p, int lane) {
        a = p[lane];
        b = p[32 + lane];
    }
    __device__ __forceinline__ void store(int* p, int n, int lane) const {
        p[lane] = lane < n ? a : 0x7FFFFFFF;
        p[32 + lane] = lane + 32 < n ? b : 0x7FFFFFFF;
    }
};
__device__ unsigned long long g_mv_stats2[4];  // frames whose own list is full, blocks with an unlisted F, lidx == -2 steps, -
__device__ unsigned long long g_mv_stats[4];  // steps, c1 direct compares, c2 direct compares, blocks (-DSCPR_MVSTATS builds only)
#ifdef SCPR_MVSTATS
#define MV_STAT(...) __VA_ARGS__
#else
#define MV_STAT(...)
#endif
__global__ void __launch_bounds__(32) k_mv_cands(PWork w) {
    const int lane = threadIdx.x;
    const int f = w.pframes[blockIdx.x];
    const int nchg = w.hdr[f].n_changed, off = w.hdr[f].chg_off;
    Cands cand;  // packed (mx & 0xFFFF) | (my << 16)
    cand.a = cand.b = 0x7FFFFFFF;
    int ncand = 0;
    for (int k0 = 0; k0 < nchg; k0 += 32) {
        const int k = k0 + lane;
        int fv = 0x7FFFFFFF;
        if (k < nchg && w.blocks[off + k].has_f)
            fv = ((int)w.blocks[off + k].fmx & 0xFFFF) | ((int)w.blocks[off + k].fmy << 16);
        uint32_t todo = __ballot_sync(0xFFFFFFFFu, fv != 0x7FFFFFFF);
        int myidx = 0xFF;
        while (todo) {
            const int src = __ffs(todo) - 1;
            const int v = __shfl_sync(0xFFFFFFFFu, fv, src);
            int idx = cand.find(v, ncand, lane);
            if (idx < 0) {
                if (ncand < MAXC) {
                    cand.put(v, ncand, lane);
                    idx = ncand++;
                } else
                    idx = 0xFF;
            }
            // every block of this step with the same vector gets the same answer
            const uint32_t same = __ballot_sync(0xFFFFFFFFu, fv == v);
            if (fv == v) myidx = idx;
            todo &= ~same;
        }
        if (k < nchg) w.blocks[off + k].fidx = (uint8_t)myidx;
        MV_STAT({
            const uint32_t un = __ballot_sync(0xFFFFFFFFu, k < nchg && fv != 0x7FFFFFFF && myidx == 0xFF);
            if (lane == 0 && un) atomicAdd(&g_mv_stats2[1], (unsigned long long)__popc(un));
        })
    }
    MV_STAT(if (lane == 0 && ncand == MAXC) atomicAdd(&g_mv_stats2[0], 1ull);)
    cand.store(w.cands0 + (size_t)blockIdx.x * MAXC, ncand, lane);
    if (lane == 0) w.ncands0[blockIdx.x] = ncand;
}

// k_mv_cands_merge: the vector stored for the block above (candidate 2) is often a leftover of an
// earlier frame, so each frame's list is topped up with the lists of the preceding P frames.
constexpr int MERGE_BACK = 192;  // a drag or scroll session ends, the vectors it left in mvs[] stay for a long time
__global__ void __launch_bounds__(32) k_mv_cands_merge(PWork w) {
    const int lane = threadIdx.x;
    const int pi = blockIdx.x;
    Cands cand;
    cand.load(w.cands0 + (size_t)pi * MAXC, lane);
    int ncand = w.ncands0[pi];
    for (int back = 1; back <= MERGE_BACK && pi - back >= 0 && ncand < MAXC; back++) {
        Cands pv;
        pv.load(w.cands0 + (size_t)(pi - back) * MAXC, lane);
        const int pn = w.ncands0[pi - back];
        for (int k = 0; k < pn && ncand < MAXC; k++) {
            const int v = pv.get(k);
            if (cand.find(v, ncand, lane) < 0) {
                cand.put(v, ncand, lane);
                ncand++;
            }
        }
    }
    cand.store(w.cands + (size_t)pi * MAXC, ncand, lane);
    if (lane == 0) w.ncands[pi] = ncand;
}

// ------------------------------------------------------------------------------------------------
// k_mv_prematch: for every changed block, which of its frame's candidate vectors reproduce the
// block from the previous frame (window test included).  One warp per block -> 64-bit mask.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_mv_prematch(PWork w) {
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int slot = blockIdx.x * 4 + wi;
    if (slot >= w.total_blocks) return;
    const int pi = find_pframe(w.hdr, w.pframes, w.n_pframes, slot);
    const int f = w.pframes[pi];
    const Geo& g = w.g;
    ChgBlock& b = w.blocks[slot];
    const SubRect r = subrect_of(b.bi, b.info, g);
    const Windows win = windows_of(r, g);
    const uint8_t* cur = w.frames + (size_t)f * g.frame_bytes;
    const uint8_t* prv = f > 0 ? cur - g.frame_bytes : w.prev0;
    const int nc = w.ncands[pi];
    const int fidx = b.fidx;
    unsigned long long mask = 0;
    for (int k = 0; k < nc; k++) {
        if (k == fidx) {  // F(b) matches by construction and lies inside the far window
            mask |= 1ull << k;
            continue;
        }
        const int cv = w.cands[(size_t)pi * MAXC + k];
        const int mx = (int)(int16_t)(cv & 0xFFFF), my = cv >> 16;
        if (in_far_window(r, win, mx, my) && warp_match(cur, prv, g, r, mx, my, lane)) mask |= 1ull << k;
    }
    if (lane == 0) {
        b.mmask = (uint32_t)mask;
        b.mmask2 = (uint32_t)(mask >> 32);
    }
}

// ------------------------------------------------------------------------------------------------
// k_mv_resolve: the serial part of FindMV (screencap.cpp:715-735), one warp walking P frames and
// their changed blocks in order.  Candidate 1 = last_mv, candidate 2 = persistent MV of the block
// above; search hits (= F) update last_mv, shortcut hits do not.
// A step takes up to 32 consecutive changed blocks, one per lane, and resolves them together under
// the assumption that last_mv does not change inside the step: with the prematch masks that is a
// table look-up per lane.  last_mv only changes at a search hit, so the step is cut right after the
// first lane that had one; everything before it is final.  The MV-repeat flag (screencap.cpp:1202)
// and the "previous pixel-coded block" link are prefix operations over the step (ballots).
// A step never contains a block together with its upper neighbour (their indices differ by nbx),
// so the upper MV can be read from mvs[] when the step is loaded.
// ------------------------------------------------------------------------------------------------
// Two things keep memory latency off the serial path: the persistent mvs[] array is held in shared memory while
// the kernel runs (SMV; written through to global memory), and the block records of a frame are read 64 ahead into
// two register windows, a step picking its 32 records out of them with shuffles.
struct MvRec {
    uint32_t bi, info, mmask, mmask2;
    int fidx, fv;
    int sp;   // speculated compare (helper warps): vector tried, packed; valid when spf & 1, its outcome in spf bit 1
    int spf;
};
__device__ __forceinline__ MvRec load_mvrec(const ChgBlock* blocks, int k, int nchg) {
    MvRec r;
    r.bi = r.info = r.mmask = r.mmask2 = 0;
    r.fidx = 0x100;
    r.fv = 0;
    r.sp = r.spf = 0;
    if (k < nchg) {
        const ChgBlock& b = blocks[k];
        r.bi = b.bi; r.info = b.info; r.mmask = b.mmask; r.mmask2 = b.mmask2; r.fidx = b.has_f ? b.fidx : 0x100;
        r.fv = ((int)b.fmx & 0xFFFF) | ((int)b.fmy << 16);
        r.sp = ((int)b.mx & 0xFFFF) | ((int)b.my << 16);
        r.spf = b.pad[0];
    }
    return r;
}
// Helper warps (every warp that does not share warp 0's scheduler) answer the expensive question ahead of time: while warp 0
// resolves frame pi, they take the changed blocks of frame pi + 1, read the vector currently stored for the block above, and --
// when it is not one of that frame's prematched candidates -- run the direct compare, leaving {vector, outcome} in the block
// record (mx / my / pad[0], which the resolve overwrites).  The stored vectors only change where a block is motion-coded, so the
// guess is almost always the vector warp 0 will find there one frame later; it checks (vector equal?) and falls back to its own
// compare otherwise.  The 25 000 serial compares of a 600-frame desktop clip (1.1 us each) become a few parallel rounds per frame.
constexpr int MVR_WARPS = 16;  // MVR_RES resolver warps (below) + the speculating warps
template <bool SMV>
__device__ void mv_speculate(const PWork& w, const uint32_t* s_mvs, int pi, int hw, int nh, int lane) {
    const Geo& g = w.g;
    const int f = w.pframes[pi];
    const int nchg = w.hdr[f].n_changed, off = w.hdr[f].chg_off;
    const uint8_t* cur = w.frames + (size_t)f * g.frame_bytes;
    const uint8_t* prv = f > 0 ? cur - g.frame_bytes : w.prev0;
    const int nc = w.ncands[pi];
    Cands cand;
    cand.load(w.cands + (size_t)pi * MAXC, lane);
    for (int k = hw; k < nchg; k += nh) {
        ChgBlock& b = w.blocks[off + k];
        const uint32_t bi = b.bi, info = b.info;
        int flag = 0, uv = 0;
        if (bi >= (uint32_t)g.nbx) {
            if (SMV)
                uv = (int)((const volatile uint32_t*)s_mvs)[bi - g.nbx];
            else {
                const volatile int* u = (const volatile int*)(w.mvs + (bi - g.nbx));
                uv = (u[0] & 0xFFFF) | (u[1] << 16);
            }
            uv = __shfl_sync(0xFFFFFFFFu, uv, 0);  // one value for the whole warp even while warp 0 is writing
            if (uv != 0 && cand.find(uv, nc, lane) < 0) {
                const SubRect r = subrect_of(bi, info, g);
                const Windows win = windows_of(r, g);
                const int mx = (int)(int16_t)(uv & 0xFFFF), my = uv >> 16;
                const bool hit = in_far_window(r, win, mx, my) && warp_match_full(cur, prv, g, r, mx, my, lane);
                flag = 1 | (hit ? 2 : 0);
            }
        }
        if (lane == 0) {
            b.mx = (int16_t)(uv & 0xFFFF);
            b.my = (int16_t)(uv >> 16);
            b.pad[0] = (uint8_t)flag;
        }
    }
}
// Frames in a pipeline.  The only things one frame's resolve hands to the next are the stored vectors: block b of frame p reads
// mvs[b - nbx] and may write mvs[b].  Frame p + 1 can therefore run while frame p is still going, as long as it stays more than one
// block row behind it -- then everything it reads has been settled by the earlier frames, and nothing it writes is still to be read
// by them.  MVR_RES resolver warps (one per scheduler) take the frames round robin; each publishes {frame, first block of the step
// it is about to do} in a shared word before every step and waits until every earlier frame still in flight is past its own step's
// last block + one row.  The remaining warps speculate a few frames ahead (mv_speculate) and count the frames they have finished.
constexpr int MVR_RES = 4;
constexpr int MVR_LOOK = MVR_RES + 2;   // the speculating warps stay at most this many frames ahead of the slowest resolver
constexpr uint32_t MVR_DONE = 0x1FFFFu; // block position "frame complete" (block indices are below 65536)
__device__ __forceinline__ uint32_t ldv_u32(const volatile uint32_t* p) { return *p; }
__device__ __forceinline__ unsigned long long ldv_u64(const volatile unsigned long long* p) { return *p; }
constexpr unsigned long long MVR_IDLE = ~0ull;  // a resolver warp without (more) frames
// The waits below cannot deadlock (a frame only waits for earlier frames and for speculation that does not depend on it), but a
// kernel that spins for ever takes the device with it: after ~10 s of spinning on one word the kernel traps instead.
constexpr uint32_t MVR_SPIN_LIMIT = 1u << 28;
template <bool SMV>
__global__ void __launch_bounds__(32 * MVR_WARPS, 1) k_mv_resolve(PWork w) {
    extern __shared__ uint32_t s_mvs[];  // SMV: packed vector of every block
    __shared__ unsigned long long s_key[MVR_RES];  // resolver r: frame << 17 | blocks below this index are done (MVR_DONE: the whole frame)
    __shared__ uint32_t s_spec;          // frames the speculating warps have finished
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1;
    const Geo& g = w.g;
    if (SMV) {
        for (int i = threadIdx.x; i < g.nb; i += blockDim.x) {
            const int2 u = w.mvs[i];
            s_mvs[i] = ((uint32_t)u.x & 0xFFFFu) | ((uint32_t)u.y << 16);
        }
    }
    if (threadIdx.x < MVR_RES) s_key[threadIdx.x] = (int)threadIdx.x < w.n_pframes ? 0ull : MVR_IDLE;
    if (threadIdx.x == 0) s_spec = 0;
    __syncthreads();
    const int nh = MVR_WARPS - MVR_RES;
    if (warp >= MVR_RES) {
        // ---- speculating warps: frame after frame, a bounded distance ahead of the resolvers
        const int hw = warp - MVR_RES;
        for (int q = 0; q < w.n_pframes; q++) {
            if (q >= MVR_LOOK) {
                const unsigned long long need = (unsigned long long)(q - MVR_LOOK) << 17;
                for (uint32_t spins = 0;; spins++) {
                    unsigned long long lo = MVR_IDLE;
#pragma unroll
                    for (int r = 0; r < MVR_RES; r++) lo = min(lo, ldv_u64(&s_key[r]));
                    if (lo >= need) break;
                    if (spins > MVR_SPIN_LIMIT / 8) __trap();  // (see MVR_SPIN_LIMIT)
                    __nanosleep(200);
                }
            }
            mv_speculate<SMV>(w, s_mvs, q, hw, nh, lane);
            __threadfence_block();
            asm volatile("bar.sync 1, %0;" ::"r"(nh * 32) : "memory");  // the speculating warps only: the frame is finished by all of them
            if (hw == 0 && lane == 0) *(volatile uint32_t*)&s_spec = (uint32_t)(q + 1);
        }
        return;
    }
    // ---- resolver warps.  The frame's header words (frame index, block count, offset, candidates) are fetched one frame ahead:
    // three dependent global round trips per frame otherwise sit in front of the first step
    int nf = 0, nnchg = 0, noff = 0, nnc = 0;
    Cands ncand;
    ncand.a = ncand.b = 0x7FFFFFFF;
    if (warp < w.n_pframes) {
        nf = w.pframes[warp];
        nnchg = w.hdr[nf].n_changed;
        noff = w.hdr[nf].chg_off;
        nnc = w.ncands[warp];
        ncand.load(w.cands + (size_t)warp * MAXC, lane);
    }
    for (int pi = warp; pi < w.n_pframes; pi += MVR_RES) {
        const int f = nf, nchg = nnchg, off = noff, nc = nnc;
        const Cands cand = ncand;  // lane k: candidates k and 32 + k
        if (pi + MVR_RES < w.n_pframes) {
            nf = w.pframes[pi + MVR_RES];
            nnchg = w.hdr[nf].n_changed;
            noff = w.hdr[nf].chg_off;
            nnc = w.ncands[pi + MVR_RES];
            ncand.load(w.cands + (size_t)(pi + MVR_RES) * MAXC, lane);
        }
        // the speculating warps have answered this frame's questions (and their writes to the block records are visible)
        for (uint32_t spins = 0; ldv_u32(&s_spec) < (uint32_t)(pi + 1); spins++)
            if (spins > MVR_SPIN_LIMIT) __trap();
        __threadfence_block();
        const uint8_t* cur = w.frames + (size_t)f * g.frame_bytes;
        const uint8_t* prv = f > 0 ? cur - g.frame_bytes : w.prev0;
        int lv = 0, lidx = -1;   // last_mv (packed) and its index in the candidate list (-1: (0,0), -2: not listed)
        int cv = 0;              // last coded MV (lastmx/lastmy of CompressP, screencap.cpp:1177)
        int prev_nonmv = -1;
        int k0 = 0;
        int w0 = 0;              // records [w0, w0 + 32) are in r0, [w0 + 32, w0 + 64) in r1
        MvRec r0 = load_mvrec(w.blocks + off, lane, nchg), r1 = load_mvrec(w.blocks + off, 32 + lane, nchg);
        while (k0 < nchg) {
            if (k0 - w0 >= 32) {
                r0 = r1;
                w0 += 32;
                r1 = load_mvrec(w.blocks + off, w0 + 32 + lane, nchg);
            }
            const int k = k0 + lane;
            const int jw = k0 - w0 + lane, js = jw & 31;
            const bool lo = jw < 32;
            uint32_t bi, info, mmask, mmask2;
            int fidx, fv, uv = 0, sp, spf;
            {
                const int s0 = __shfl_sync(0xFFFFFFFFu, r0.sp, js), s1 = __shfl_sync(0xFFFFFFFFu, r1.sp, js);
                const int t0 = __shfl_sync(0xFFFFFFFFu, r0.spf, js), t1 = __shfl_sync(0xFFFFFFFFu, r1.spf, js);
                sp = lo ? s0 : s1; spf = lo ? t0 : t1;
                const uint32_t a0 = __shfl_sync(0xFFFFFFFFu, r0.bi, js), a1 = __shfl_sync(0xFFFFFFFFu, r1.bi, js);
                const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, r0.info, js), b1 = __shfl_sync(0xFFFFFFFFu, r1.info, js);
                const uint32_t c0 = __shfl_sync(0xFFFFFFFFu, r0.mmask, js), c1 = __shfl_sync(0xFFFFFFFFu, r1.mmask, js);
                const int d0 = __shfl_sync(0xFFFFFFFFu, r0.fidx, js), d1 = __shfl_sync(0xFFFFFFFFu, r1.fidx, js);
                const int e0 = __shfl_sync(0xFFFFFFFFu, r0.fv, js), e1 = __shfl_sync(0xFFFFFFFFu, r1.fv, js);
                const uint32_t g0 = __shfl_sync(0xFFFFFFFFu, r0.mmask2, js), g1 = __shfl_sync(0xFFFFFFFFu, r1.mmask2, js);
                bi = lo ? a0 : a1; info = lo ? b0 : b1; mmask = lo ? c0 : c1; mmask2 = lo ? g0 : g1; fidx = lo ? d0 : d1; fv = lo ? e0 : e1;
            }
            const uint32_t bi0 = __shfl_sync(0xFFFFFFFFu, bi, 0);
            int cnt = __popc(__ballot_sync(0xFFFFFFFFu, k < nchg && bi - bi0 < (uint32_t)g.nbx));
            {
                // every block below bi0 is done (the stores of the previous step are ordered in front of this word) ...
                if (SMV) __threadfence_block(); else __threadfence();
                if (lane == 0) *(volatile unsigned long long*)&s_key[warp] = ((unsigned long long)pi << 17) | bi0;
                // ... and every earlier frame still in flight must be more than a row past this step's last block
                const uint32_t bmax = __shfl_sync(0xFFFFFFFFu, bi, cnt - 1);
                const uint32_t past = min(bmax + (uint32_t)g.nbx + 1u, MVR_DONE);
#pragma unroll
                for (int d = 1; d < MVR_RES; d++) {
                    const int q = pi - d;
                    if (q >= 0) {
                        const unsigned long long need = ((unsigned long long)q << 17) | past;
                        const volatile unsigned long long* kp = &s_key[(warp - d + MVR_RES) % MVR_RES];
                        for (uint32_t spins = 0; ldv_u64(kp) < need; spins++)
                            if (spins > MVR_SPIN_LIMIT) __trap();
                    }
                }
                if (SMV) __threadfence_block(); else __threadfence();
            }
            if (k < nchg && bi >= (uint32_t)g.nbx) {
                if (SMV)
                    uv = (int)((const volatile uint32_t*)s_mvs)[bi - g.nbx];
                else {
                    const volatile int* u = (const volatile int*)(w.mvs + (bi - g.nbx));
                    uv = (u[0] & 0xFFFF) | (u[1] << 16);
                }
            }
            const bool in = lane < cnt;
            // ---- candidate 1: last_mv (same for every lane of the step) ----
            bool found = false;
            int mv = 0;
            if (lidx >= 0)
                found = in && (((lidx < 32 ? mmask : mmask2) >> (lidx & 31)) & 1);
            else if (lidx == -2) {  // last_mv is not in the candidate list: direct compares, lane by lane
                MV_STAT(if (lane == 0) { atomicAdd(&g_mv_stats[1], (unsigned long long)cnt); atomicAdd(&g_mv_stats2[2], 1ull); })
                for (int i = 0; i < cnt; i++) {
                    const SubRect r = subrect_of(__shfl_sync(0xFFFFFFFFu, bi, i), __shfl_sync(0xFFFFFFFFu, info, i), g);
                    const Windows win = windows_of(r, g);
                    const int mx = (int)(int16_t)(lv & 0xFFFF), my = lv >> 16;
                    const bool hit = in_far_window(r, win, mx, my) && warp_match_full(cur, prv, g, r, mx, my, lane);
                    if (lane == i) found = hit;
                }
            }
            if (found) mv = lv;
            // ---- candidate 2: the vector stored for the block above, if it differs from last_mv ----
            const bool try2 = in && !found && bi >= (uint32_t)g.nbx && uv != lv && uv != 0;
            int uidx = -1;  // its index in the candidate list: one ballot per distinct vector among the lanes that need it
            for (uint32_t todo = __ballot_sync(0xFFFFFFFFu, try2); todo;) {
                const int v = __shfl_sync(0xFFFFFFFFu, uv, __ffs(todo) - 1);
                const int at = cand.find(v, nc, lane);
                const bool mine = uv == v;
                if (mine) uidx = at;
                todo &= ~__ballot_sync(0xFFFFFFFFu, mine);
            }
            bool f2 = try2 && uidx >= 0 && (((uidx < 32 ? mmask : mmask2) >> (uidx & 31)) & 1);
            const bool spec_ok = (spf & 1) && sp == uv;  // a helper warp has already compared this block with this vector
            if (try2 && uidx < 0 && spec_ok) f2 = (spf & 2) != 0;
            uint32_t slow = __ballot_sync(0xFFFFFFFFu, try2 && uidx < 0 && !spec_ok);
            MV_STAT(if (lane == 0) {
                atomicAdd(&g_mv_stats[0], 1ull);
                atomicAdd(&g_mv_stats[2], (unsigned long long)__popc(slow));
            })
            while (slow) {  // a stale vector that is not one of this frame's candidates: compare directly
                const int i = __ffs(slow) - 1;
                slow &= slow - 1;
                const SubRect r = subrect_of(__shfl_sync(0xFFFFFFFFu, bi, i), __shfl_sync(0xFFFFFFFFu, info, i), g);
                const Windows win = windows_of(r, g);
                const int u = __shfl_sync(0xFFFFFFFFu, uv, i);
                const int mx = (int)(int16_t)(u & 0xFFFF), my = u >> 16;
                const bool hit = in_far_window(r, win, mx, my) && warp_match_full(cur, prv, g, r, mx, my, lane);
                if (lane == i) f2 = hit;
            }
            if (f2) {
                found = true;
                mv = uv;
            }
            // ---- the fixed-order search: first hit wins and becomes last_mv; cut the step after it ----
            const bool hit3 = in && !found && fidx != 0x100;
            const uint32_t h3 = __ballot_sync(0xFFFFFFFFu, hit3);
            if (h3) {
                const int tH = __ffs(h3) - 1;
                cnt = tH + 1;
                if (lane == tH) {
                    found = true;
                    mv = fv;
                }
                lv = __shfl_sync(0xFFFFFFFFu, fv, tH);
                const int fi = __shfl_sync(0xFFFFFFFFu, fidx, tH);
                lidx = fi == 0xFF ? -2 : fi;
            }
            const bool inn = lane < cnt;
            // ---- repeat flag and links: prefix operations over the lanes of the step ----
            const uint32_t fm = __ballot_sync(0xFFFFFFFFu, inn && found);
            const uint32_t before = fm & lt;
            const int pf_lane = before ? 31 - __clz(before) : 0;
            const int pmv = __shfl_sync(0xFFFFFFFFu, mv, pf_lane);
            const int prev_coded = before ? pmv : cv;  // a repeated MV leaves lastmx/lastmy unchanged = same value
            const bool rep = found && bi > 0 && mv == prev_coded;
            const uint32_t nm = __ballot_sync(0xFFFFFFFFu, inn && !found);
            const uint32_t nbefore = nm & lt;
            const int my_prev_nonmv = nbefore ? k0 + 31 - __clz(nbefore) : prev_nonmv;
            if (inn) {
                ChgBlock& b = w.blocks[off + k];
                b.bt = (uint8_t)(((info & BI_PARTIAL) ? 2 : 1) + (found ? 2 : 0));
                b.rep = rep;
                b.mx = (int16_t)(mv & 0xFFFF);
                b.my = (int16_t)(mv >> 16);
                b.prev_nonmv = found ? -1 : my_prev_nonmv;
                if (found) {
                    if (SMV) s_mvs[bi] = (uint32_t)mv;
                    w.mvs[bi] = make_int2((int)(int16_t)(mv & 0xFFFF), mv >> 16);
                }
            }
            if (fm) cv = __shfl_sync(0xFFFFFFFFu, mv, 31 - __clz(fm));
            if (nm) prev_nonmv = k0 + 31 - __clz(nm);
            __syncwarp();
            MV_STAT(if (lane == 0) atomicAdd(&g_mv_stats[3], (unsigned long long)cnt);)
            k0 += cnt;
        }
        __threadfence();  // mvs[] of this frame visible ...
        if (lane == 0) *(volatile unsigned long long*)&s_key[warp] = ((unsigned long long)pi << 17) | MVR_DONE;  // ... before the frame is declared complete
    }
    __threadfence();
    if (lane == 0) *(volatile unsigned long long*)&s_key[warp] = MVR_IDLE;  // no more frames on this warp: nobody waits for it
}

// ------------------------------------------------------------------------------------------------
// k_p_runs: pixel typing + run segmentation of pixel-coded blocks; per-block event counts.
// One warp per changed block.  runs[slot*256 + r] = (ptype << 8) | n.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool grad_ok(uint32_t p, uint32_t l, uint32_t t, uint32_t tl) {
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int sh = 8 * c;
        const int v = (int)((l >> sh) & 255) + (int)((t >> sh) & 255) - (int)((tl >> sh) & 255);
        ok = ok && ((int)((p >> sh) & 255) == v);
    }
    return ok;
}

__global__ void __launch_bounds__(128) k_p_runs(PWork w) {
    __shared__ uint16_t s_px[4][256];  // per pixel: best type (bits 0-2) | fit mask for types 0..5 (bits 4-9)
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int slot = blockIdx.x * 4 + wi;
    if (slot >= w.total_blocks) return;
    ChgBlock& b = w.blocks[slot];
    const Geo& g = w.g;
    const SubRect r = subrect_of(b.bi, b.info, g);
    const uint32_t sxy_ev = r.partial ? 4u : 0u;
    if (b.bt >= 3) {  // MV block: [SXY x4] BOOL [MX MY]
        if (lane == 0) {
            b.n_runs = 0;
            b.n_ev = sxy_ev + (b.rep ? 1u : 3u);
        }
        return;
    }
    const int pi = find_pframe(w.hdr, w.pframes, w.n_pframes, slot);
    const int f = w.pframes[pi];
    const uint8_t* cur = w.frames + (size_t)f * g.frame_bytes;
    const uint8_t* prv = f > 0 ? cur - g.frame_bytes : w.prev0;
    const int npx = r.w * r.h;
    for (int p = lane; p < npx; p += 32) {
        const int xx = p % r.w, yy = p / r.w, x = r.x1 + xx, y = r.y1 + yy;
        const uint32_t c = load_px(cur, g, x, y);
        const uint32_t pr = load_px(prv, g, x, y);
        // previous pixel in the sub-rect's raster order (lasti, screencap.cpp:1045-1061)
        uint32_t last = 0;
        if (p > 0) {
            const int q = p - 1;
            last = load_px(cur, g, r.x1 + q % r.w, r.y1 + q / r.w);
        }
        uint32_t fit = (p > 0 && c == last) ? 1u : 0u;  // type 0 continues on "equals last pixel"
        int best;
        if (x > 0 && y > 0) {
            const uint32_t l = load_px(cur, g, x - 1, y), t = load_px(cur, g, x, y - 1), tl = load_px(cur, g, x - 1, y - 1);
            const bool e1 = c == l, e3 = c == pr, e5 = c == tl, e2 = c == t, e4 = grad_ok(c, l, t, tl);
            fit |= (e1 ? 2u : 0u) | (e2 ? 4u : 0u) | (e3 ? 8u : 0u) | (e4 ? 16u : 0u) | (e5 ? 32u : 0u);
            best = e1 ? 1 : e3 ? 3 : e5 ? 5 : e2 ? 2 : e4 ? 4 : 0;  // GetPixelTypeP priority
        } else {
            const bool e3 = c == pr;
            fit |= e3 ? 8u : 0u;
            best = e3 ? 3 : 0;  // GetPixelTypeP0
        }
        s_px[wi][p] = (uint16_t)(best | (fit << 4));
    }
    __syncwarp();
    if (lane == 0) {
        uint16_t* runs = w.runs + (size_t)slot * 256;
        int nr = 0, type = s_px[wi][0] & 7, n = 1;
        uint32_t nev = sxy_ev;
        for (int p = 1; p < npx; p++) {
            const uint32_t v = s_px[wi][p];
            if (n < 255 && ((v >> (4 + type)) & 1))
                n++;
            else {
                runs[nr++] = (uint16_t)((type << 8) | n);
                nev += type ? 2u : 5u;
                type = v & 7;
                n = 1;
            }
        }
        runs[nr++] = (uint16_t)((type << 8) | n);
        nev += type ? 2u : 5u;
        b.n_runs = nr;
        b.n_ev = nev;
    }
}

// ------------------------------------------------------------------------------------------------
// k_p_count: per P frame -- block-type RLE (screencap.cpp:1155-1169) into scratch, exclusive scan
// of the per-block event counts, frame totals.  One CTA per P frame.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_p_count(PWork w) {
    const int f = w.pframes[blockIdx.x];
    const Geo& g = w.g;
    PFrameHdr& h = w.hdr[f];
    ChgBlock* blocks = w.blocks + h.chg_off;
    const int nchg = h.n_changed;
    __shared__ uint32_t s_hdr_ev;
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base;
    // dense block types of the coded range [xx1, xx2] live in the upper half of the scratch
    uint32_t* rle = w.bts_rle + (size_t)blockIdx.x * 2 * g.nb;
    uint8_t* dense = reinterpret_cast<uint8_t*>(rle + g.nb);
    for (int i = threadIdx.x; i < g.nb; i += 256) dense[i] = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < nchg; k += 256) dense[blocks[k].bi] = blocks[k].bt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int oldt = -1, n = -1, np = 0;
        for (int x = h.xx1; x <= h.xx2; x++) {
            const int t = dense[x];
            if (t == oldt && n < 255)
                n++;
            else {
                if (n > 0) rle[np++] = ((uint32_t)oldt << 8) | (uint32_t)n;
                oldt = t;
                n = 1;
            }
        }
        rle[np++] = ((uint32_t)oldt << 8) | (uint32_t)n;
        s_hdr_ev = 4u + 2u * (uint32_t)np;
        s_base = 4u + 2u * (uint32_t)np;
        h.n_hdr_ev = s_hdr_ev;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    for (int k0 = 0; k0 < nchg; k0 += 256) {
        const int k = k0 + threadIdx.x;
        const uint32_t v = k < nchg ? blocks[k].n_ev : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) s_warp[wi] = inc;
        __syncthreads();
        uint32_t woff = s_base;
        for (int j = 0; j < wi; j++) woff += s_warp[j];
        if (k < nchg) blocks[k].ev_off = woff + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int j = 0; j < 8; j++) t += s_warp[j];
            s_base += t;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) h.n_ev = s_base;
}

// ------------------------------------------------------------------------------------------------
// k_p_emit_hdr: XX bytes + (BT, BN) pairs.  One CTA per P frame.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_p_emit_hdr(PWork w) {
    const int f = w.pframes[blockIdx.x];
    const Geo& g = w.g;
    const PFrameHdr& h = w.hdr[f];
    uint32_t* ev = w.events + w.frame_ev_off[f];
    const uint32_t* rle = w.bts_rle + (size_t)blockIdx.x * 2 * g.nb;
    if (threadIdx.x == 0) {
        ev[0] = make_ev(CX_XX, h.xx1 & 255);
        ev[1] = make_ev(CX_XX, (h.xx1 >> 8) & 255);
        ev[2] = make_ev(CX_XX, h.xx2 & 255);
        ev[3] = make_ev(CX_XX, (h.xx2 >> 8) & 255);
    }
    const int np = (int)(h.n_hdr_ev - 4) / 2;
    for (int i = threadIdx.x; i < np; i += 256) {
        ev[4 + 2 * i] = make_ev(CX_BT, rle[i] >> 8);
        ev[5 + 2 * i] = make_ev(CX_NTAB2, rle[i] & 255);
    }
}

// colour context ids of a literal pixel `c` that follows pixel `last` (WritePixel / MAKECX1,
// screencap.cpp:609-627, screencap.h:36); has_last == false -> cx = cx1 = 0 (frame start)
__device__ __forceinline__ void emit_literal(uint32_t* ev, uint32_t c, uint32_t last, bool has_last) {
    const uint32_t r = c & 255, gg = (c >> 8) & 255, bb = (c >> 16) & 255;
    const uint32_t lg = has_last ? ((last >> 8) & 255) >> 2 : 0, lb = has_last ? ((last >> 16) & 255) >> 2 : 0;
    ev[0] = make_ev(CX_COLOR + 0 * 4096 + (int)(lb + (lg << 6)), r);
    ev[1] = make_ev(CX_COLOR + 1 * 4096 + (int)((r >> 2) + (lb << 6)), gg);
    ev[2] = make_ev(CX_COLOR + 2 * 4096 + (int)((gg >> 2) + ((r >> 2) << 6)), bb);
}

// ------------------------------------------------------------------------------------------------
// k_p_emit: events of every changed block.  One warp per block.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_p_emit(PWork w) {
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int slot = blockIdx.x * 4 + wi;
    if (slot >= w.total_blocks) return;
    const ChgBlock& b = w.blocks[slot];
    const Geo& g = w.g;
    const int pi = find_pframe(w.hdr, w.pframes, w.n_pframes, slot);
    const int f = w.pframes[pi];
    const SubRect r = subrect_of(b.bi, b.info, g);
    uint32_t* ev = w.events + w.frame_ev_off[f] + b.ev_off;
    if (r.partial) {
        if (lane == 0) {
            ev[0] = make_ev(CX_SXY + 0, r.x1 - r.bx * 16);
            ev[1] = make_ev(CX_SXY + 1, r.y1 - r.by * 16);
            ev[2] = make_ev(CX_SXY + 2, r.x2 - 1 - r.bx * 16);
            ev[3] = make_ev(CX_SXY + 3, r.y2 - 1 - r.by * 16);
        }
        ev += 4;
    }
    if (b.bt >= 3) {
        if (lane == 0) {
            uint32_t* iv = w.intervals + (ev - w.events);
            ev[0] = make_ev(CX_BOOL, b.rep);
            iv[0] = make_iv(PROB_SCALE / 2, b.rep ? PROB_SCALE / 2 : 0);  // encodeBool, screencap.h:407-410
            if (!b.rep) {
                ev[1] = make_ev(CX_MV + 0, b.mx + 256);
                ev[2] = make_ev(CX_MV + 1, b.my + 256);
            }
        }
        return;
    }
    const uint8_t* cur = w.frames + (size_t)f * g.frame_bytes;
    const uint16_t* runs = w.runs + (size_t)slot * 256;
    const int nr = (int)b.n_runs;
    // context pixel for the first run: bottom-right pixel of the previous pixel-coded block
    uint32_t first_last = 0;
    bool first_has = false;
    if (b.prev_nonmv >= 0) {
        const ChgBlock& pb = w.blocks[w.hdr[f].chg_off + b.prev_nonmv];
        const SubRect pr = subrect_of(pb.bi, pb.info, g);
        first_last = load_px(cur, g, pr.x2 - 1, pr.y2 - 1);
        first_has = true;
    }
    // each lane owns 8 consecutive runs; prefix sums of pixel counts and event counts
    uint32_t rv[8];
    uint32_t npx = 0, nev = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int ri = lane * 8 + j;
        rv[j] = ri < nr ? runs[ri] : 0u;
        if (ri < nr) {
            npx += rv[j] & 255;
            nev += (rv[j] >> 8) ? 2u : 5u;
        }
    }
    uint32_t ipx = npx, iev = nev;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t a = __shfl_up_sync(0xFFFFFFFFu, ipx, d), c = __shfl_up_sync(0xFFFFFFFFu, iev, d);
        if (lane >= d) { ipx += a; iev += c; }
    }
    uint32_t px = ipx - npx, eo = iev - nev;
    // type of the run just before this lane's first run
    uint32_t prev_type = __shfl_up_sync(0xFFFFFFFFu, rv[7] >> 8, 1);
    if (lane == 0) prev_type = 0;  // lastptype restarts at 0 per block (screencap.cpp:1218)
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int ri = lane * 8 + j;
        if (ri < nr) {
            const uint32_t type = rv[j] >> 8, n = rv[j] & 255;
            uint32_t* e = ev + eo;
            *e++ = make_ev(CX_PTYPE + (int)prev_type, type);
            if (type == 0) {
                const uint32_t c = load_px(cur, g, r.x1 + (int)(px % r.w), r.y1 + (int)(px / r.w));
                uint32_t last = first_last;
                bool has = first_has;
                if (px > 0) {
                    const uint32_t q = px - 1;
                    last = load_px(cur, g, r.x1 + (int)(q % r.w), r.y1 + (int)(q / r.w));
                    has = true;
                }
                emit_literal(e, c, last, has);
                e += 3;
            }
            *e = make_ev(CX_NTAB + (int)type, n);
            eo += type ? 2u : 5u;
            px += n;
            prev_type = type;
        }
    }
}

void mv_stats_report() {
#ifndef SCPR_MVSTATS
    return;
#endif
    unsigned long long h[4] = {0, 0, 0, 0}, z[4] = {0, 0, 0, 0};
    cudaMemcpyFromSymbol(h, g_mv_stats, sizeof(h));
    cudaMemcpyToSymbol(g_mv_stats, z, sizeof(z));
    unsigned long long h2[4] = {0, 0, 0, 0};
    cudaMemcpyFromSymbol(h2, g_mv_stats2, sizeof(h2));
    cudaMemcpyToSymbol(g_mv_stats2, z, sizeof(z));
    fprintf(stderr, "[scpr timing] mv_resolve: %llu blocks in %llu steps, direct compares: last_mv %llu (in %llu steps), upper %llu | frames with a full own candidate list %llu, blocks with an unlisted F %llu\n",
            h[3], h[0], h[1], h2[2], h[2], h2[0], h2[1]);
}

void launch_p_stage_a(const PWork& w, cudaStream_t st, uint64_t* launches) {
    if (w.total_blocks > 0) {
        k_mv_search<<<(w.total_blocks + 3) / 4, 128, 0, st>>>(w);
        if (w.tm) w.tm->mark("mv_search");
        k_mv_cands<<<w.n_pframes, 32, 0, st>>>(w);
        k_mv_cands_merge<<<w.n_pframes, 32, 0, st>>>(w);
        ++*launches;
        k_mv_prematch<<<(w.total_blocks + 3) / 4, 128, 0, st>>>(w);
        *launches += 2;
        if (w.tm) w.tm->mark("mv_prematch");
        if (w.pre_resolve) {  // everything enqueued so far is independent of mvs[]: it runs while the host waits for the hand-off
            *w.in_hook = true;
            w.pre_resolve(w.hook_user);
            *w.in_hook = false;
        }
        const size_t mv_smem = (size_t)w.g.nb * 4;
        if (mv_smem <= 200 * 1024) {
            cudaFuncSetAttribute(k_mv_resolve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mv_smem);
            k_mv_resolve<true><<<1, 32 * MVR_WARPS, mv_smem, st>>>(w);
        } else
            k_mv_resolve<false><<<1, 32 * MVR_WARPS, 0, st>>>(w);
        if (w.tm) w.tm->mark("mv_resolve");
        if (w.post_resolve) {  // mvs[] is final for this call: the next range may start its resolve
            cudaStreamSynchronize(st);
            *w.in_hook = true;
            w.post_resolve(w.hook_user);
            *w.in_hook = false;
        }
        k_p_runs<<<(w.total_blocks + 3) / 4, 128, 0, st>>>(w);
        if (w.tm) w.tm->mark("p_runs");
        *launches += 3;
    }
    else {  // nothing to resolve: mvs[] passes through
        if (w.pre_resolve || w.post_resolve) *w.in_hook = true;
        if (w.pre_resolve) w.pre_resolve(w.hook_user);
        if (w.post_resolve) w.post_resolve(w.hook_user);
        if (w.pre_resolve || w.post_resolve) *w.in_hook = false;
    }
    if (w.n_pframes > 0) {
        k_p_count<<<w.n_pframes, 256, 0, st>>>(w);
        ++*launches;
    }
}

void launch_p_emit(const PWork& w, cudaStream_t st, uint64_t* launches) {
    if (w.n_pframes > 0) {
        k_p_emit_hdr<<<w.n_pframes, 256, 0, st>>>(w);
        ++*launches;
    }
    if (w.total_blocks > 0) {
        k_p_emit<<<(w.total_blocks + 3) / 4, 128, 0, st>>>(w);
        ++*launches;
    }
}

}  // namespace scpr
