// iframe.cu -- stage A for I frames: per-pixel predictor classification, run segmentation and event
// emission.
//
// Replaces (reference, one band = the canonical 1-thread stream):
//   ClassifyPixelsI / GetPixelType / PixelTypeFits   screencap.cpp:876-919, 502-521, 560-574
//   CompressI's serialisation                        screencap.cpp:344-388
//
// The reference walks pixels serially: a run takes the best predictor of its first pixel and
// extends while that predictor still fits (cap 255).  Where a run starts therefore depends on
// where the previous one ended.  Here every pixel q gets the descriptor of the run that WOULD start
// at q (type, length) in parallel -- lengths come from ctz-style scans over per-predictor bit
// arrays in shared memory -- and the actual run starts are the orbit of pixel X+1 under
// q -> q + len(q).  That orbit is resolved per 2048-pixel chunk: pointer doubling in shared memory
// gives each chunk a 255-entry table "entry offset -> entry offset of the next chunk", a short
// serial pass chains the tables, and then chunks mark their run starts and emit independently.
#include "codec.h"

namespace scpr {

constexpr int HALO = 256;
constexpr int CWORDS = (ICHUNK + HALO) / 32;  // 72

__device__ __forceinline__ bool grad_ok_i(uint32_t p, uint32_t l, uint32_t t, uint32_t tl) {
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int sh = 8 * c;
        const int v = (int)((l >> sh) & 255) + (int)((t >> sh) & 255) - (int)((tl >> sh) & 255);
        ok = ok && ((int)((p >> sh) & 255) == v);
    }
    return ok;
}

__device__ __forceinline__ uint32_t px_lin(const uint8_t* f, const Geo& g, long q) {
    const int y = (int)(q / g.X), x = (int)(q - (long)y * g.X);
    return load_px(f, g, x, y);
}

// The multi-threaded reference classifies an I frame in `bands` row bands (CSquadWorker::GetSegment, squad.cpp:16-31: band b covers
// rows [Y*b/bands, Y*(b+1)/bands)), one worker each, and every band starts a new run at its first pixel (ClassifyPixelsI,
// screencap.cpp:876-919; the bands are serialised one after the other, :365-388).  That is the only way the thread count shows
// in an I frame's bytes, and in this design it is one clamp: a run that would start at pixel q may not reach past the first
// pixel of the next band.  The orbit q -> q + len(q) then lands on every band start by itself.  -> pixels left before that limit
__device__ __forceinline__ long band_room(const Geo& g, int bands, long q) {
    if (bands <= 1) return 1L << 40;
    const int y = (int)(q / g.X);
    int b = (int)(((long)y * bands) / g.Y);
    while (b + 1 < bands && (long)g.Y * (b + 1) / bands <= y) b++;
    if (b + 1 >= bands) return 1L << 40;
    return ((long)g.Y * (b + 1) / bands) * g.X - q;
}

// number of consecutive set bits starting at bit position `pos`, at most `maxn`
__device__ __forceinline__ int ones_from(const uint32_t* bits, int pos, int maxn) {
    int n = 0;
    while (n < maxn) {
        const int wd = (pos + n) >> 5, bo = (pos + n) & 31;
        const uint32_t inv = ~(bits[wd] >> bo);  // zero bits shifted in at the top read as "stop"
        const int run = inv ? __ffs(inv) - 1 : 32;
        const int avail = 32 - bo;
        if (run < avail) return min(maxn, n + run);
        n += avail;
    }
    return maxn;
}

// ------------------------------------------------------------------------------------------------
// k_i_classify: descriptors + chunk exit tables.  grid = (nchunks, n_iframes), 256 threads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_i_classify(IWork w) {
    __shared__ uint32_t s_bits[4][CWORDS];  // fit bits for predictor classes: 0 = last(1/0), 1 = top(2), 2 = grad(4), 3 = topleft(5)
    __shared__ uint8_t s_type[ICHUNK];
    __shared__ uint16_t s_jump[2][ICHUNK];
    const Geo& g = w.g;
    const int c = blockIdx.x, fi = blockIdx.y;
    const uint8_t* f = w.frames + (size_t)w.hdr[fi].frame * g.frame_bytes;
    const long total = (long)g.X * g.Y;
    const long q0 = (long)g.X + 1 + (long)c * ICHUNK;
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const int stride24 = (g.X * 3 + 3) & ~3;
    const bool padded = stride24 != g.X * 3;

    for (int i = 0; i < (ICHUNK + HALO) / 256; i++) {
        const int p = i * 256 + threadIdx.x;
        const long q = q0 + p;
        bool f1 = false, f2 = false, f4 = false, f5 = false;
        if (q < total) {
            const int y = (int)(q / g.X), x = (int)(q - (long)y * g.X);
            const uint32_t cpx = load_px(f, g, x, y);
            const uint32_t l = px_lin(f, g, q - 1);          // previous pixel in raster order (lasti)
            const uint32_t t = load_px(f, g, x, y - 1);
            uint32_t tl;
            if (!padded || x > 0)
                tl = px_lin(f, g, q - g.X - 1);              // byte offset -stride-3 without padding
            else
                tl = rgb24_triple_at(f, g, stride24, (long)y * stride24 - stride24 - 3);
            f1 = cpx == l; f5 = cpx == tl; f2 = cpx == t; f4 = grad_ok_i(cpx, l, t, tl);
            if (p < ICHUNK) s_type[p] = f1 ? 1 : f5 ? 5 : f2 ? 2 : f4 ? 4 : 0;  // GetPixelType priority
        } else if (p < ICHUNK)
            s_type[p] = 0;
        const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, f1), b1 = __ballot_sync(0xFFFFFFFFu, f2);
        const uint32_t b2 = __ballot_sync(0xFFFFFFFFu, f4), b3 = __ballot_sync(0xFFFFFFFFu, f5);
        if (lane == 0) {
            const int wd = i * 8 + wi;
            s_bits[0][wd] = b0; s_bits[1][wd] = b1; s_bits[2][wd] = b2; s_bits[3][wd] = b3;
        }
    }
    __syncthreads();
    uint16_t* desc = w.desc + (size_t)fi * total;
    for (int i = 0; i < ICHUNK / 256; i++) {
        const int p = i * 256 + threadIdx.x;
        const long q = q0 + p;
        int len = 1;
        const int type = s_type[p];
        if (q < total) {
            const int cls = type <= 1 ? 0 : type == 2 ? 1 : type == 4 ? 2 : 3;
            len = 1 + ones_from(s_bits[cls], p + 1, 254);
            len = (int)min((long)len, band_room(g, w.bands, q));
            desc[q] = (uint16_t)((type << 8) | len);
        }
        s_jump[0][p] = (uint16_t)min(p + len, 2 * ICHUNK);
    }
    __syncthreads();
    int cur = 0;
    for (int round = 0; round < 11; round++) {  // 2^11 = ICHUNK, every run advances >= 1
        for (int i = 0; i < ICHUNK / 256; i++) {
            const int p = i * 256 + threadIdx.x;
            const uint16_t j = s_jump[cur][p];
            s_jump[cur ^ 1][p] = j < ICHUNK ? s_jump[cur][j] : j;
        }
        cur ^= 1;
        __syncthreads();
    }
    uint8_t* tab = w.exit_tab + ((size_t)fi * w.nchunks + c) * 256;
    if (threadIdx.x < 255) tab[threadIdx.x] = (uint8_t)(s_jump[cur][threadIdx.x] - ICHUNK);
}

// k_i_entries: chain the exit tables.  One thread per I frame.
__global__ void k_i_entries(IWork w) {
    const int fi = blockIdx.x * blockDim.x + threadIdx.x;
    if (fi >= w.n_iframes) return;
    int e = 0;
    for (int c = 0; c < w.nchunks; c++) {
        w.entry[(size_t)fi * w.nchunks + c] = (uint16_t)e;
        e = w.exit_tab[((size_t)fi * w.nchunks + c) * 256 + e];
    }
}

// k_i_mark: run starts of each chunk, run / event counts.  grid = (nchunks, n_iframes).
__global__ void __launch_bounds__(256) k_i_mark(IWork w) {
    __shared__ uint16_t s_desc[ICHUNK];
    __shared__ int s_n;
    __shared__ uint32_t s_ev;
    const Geo& g = w.g;
    const int c = blockIdx.x, fi = blockIdx.y;
    const long total = (long)g.X * g.Y;
    const long q0 = (long)g.X + 1 + (long)c * ICHUNK;
    const int limit = (int)min((long)ICHUNK, total - q0);
    const uint16_t* desc = w.desc + (size_t)fi * total;
    for (int p = threadIdx.x; p < limit; p += 256) s_desc[p] = desc[q0 + p];
    if (threadIdx.x == 0) s_ev = 0;
    __syncthreads();
    const size_t ch = (size_t)fi * w.nchunks + c;
    uint16_t* starts = w.starts + ch * ICHUNK;
    if (threadIdx.x == 0) {
        int p = w.entry[ch], k = 0;
        while (p < limit) {
            starts[k++] = (uint16_t)p;
            p += s_desc[p] & 255;
        }
        s_n = k;
    }
    __syncthreads();
    const int n = s_n;
    uint32_t ev = 0;
    for (int k = threadIdx.x; k < n; k += 256) ev += (s_desc[starts[k]] >> 8) ? 2u : 5u;
    ev = __reduce_add_sync(0xFFFFFFFFu, ev);
    if ((threadIdx.x & 31) == 0 && ev) atomicAdd(&s_ev, ev);
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t* cc = w.chunk_cnt + ch * 4;
        cc[0] = (uint32_t)n;
        cc[1] = s_ev;
        cc[2] = n ? (uint32_t)(s_desc[starts[n - 1]] >> 8) : 0u;
    }
}

// k_i_offsets: header event count (first row + pixel (0,1), screencap.cpp:344-362), exclusive scan
// of the chunk event counts, frame total.  One warp per I frame.
__global__ void __launch_bounds__(32) k_i_offsets(IWork w) {
    const int fi = blockIdx.x, lane = threadIdx.x;
    const Geo& g = w.g;
    const uint8_t* f = w.frames + (size_t)w.hdr[fi].frame * g.frame_bytes;
    // equality runs over pixels 0..X, cap 255: a break at k when pixel k differs from k-1 or the run is full
    uint32_t breaks = 0;
    int run = 1;  // length of the run ending at the previous pixel, carried serially per 32-pixel step
    for (int k0 = 1; k0 <= g.X; k0 += 32) {
        const int k = k0 + lane;
        bool same = false;
        if (k <= g.X) same = px_lin(f, g, k) == px_lin(f, g, k - 1);
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, same);
        if (lane == 0) {
            const int cnt = min(32, g.X - k0 + 1);
            for (int j = 0; j < cnt; j++) {
                if (((m >> j) & 1) && run < 255)
                    run++;
                else {
                    breaks++;
                    run = 1;
                }
            }
        }
    }
    breaks = __shfl_sync(0xFFFFFFFFu, breaks, 0);
    const uint32_t n_hdr = 3u + 4u * breaks + 1u;
    uint32_t base = n_hdr;
    for (int c0 = 0; c0 < w.nchunks; c0 += 32) {
        const int c = c0 + lane;
        uint32_t* cc = w.chunk_cnt + ((size_t)fi * w.nchunks + c) * 4;
        const uint32_t v = c < w.nchunks ? cc[1] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += t;
        }
        if (c < w.nchunks) cc[3] = base + inc - v;
        base += __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
    if (lane == 0) {
        w.hdr[fi].n_hdr_ev = n_hdr;
        w.hdr[fi].n_ev = base;
    }
}

__device__ __forceinline__ void emit_literal_i(uint32_t* ev, uint32_t c, uint32_t last, bool has_last) {
    const uint32_t r = c & 255, gg = (c >> 8) & 255, bb = (c >> 16) & 255;
    const uint32_t lg = has_last ? ((last >> 8) & 255) >> 2 : 0, lb = has_last ? ((last >> 16) & 255) >> 2 : 0;
    ev[0] = make_ev(CX_COLOR + 0 * 4096 + (int)(lb + (lg << 6)), r);
    ev[1] = make_ev(CX_COLOR + 1 * 4096 + (int)((r >> 2) + (lb << 6)), gg);
    ev[2] = make_ev(CX_COLOR + 2 * 4096 + (int)((gg >> 2) + ((r >> 2) << 6)), bb);
}

// k_i_emit_hdr: RGB(pixel 0), then (N, RGB) per run break, final N -- all lengths in ntab[0].
// The walk over pixels 1..X is serial (the 255 cap), but its loads need not be: the warp stages 2048 pixels at a time in shared
// memory (coalesced), lane 0 walks them there (one global round trip per pixel before: 318 us per 1080p I frame, 88 us now).
__global__ void __launch_bounds__(32) k_i_emit_hdr(IWork w) {
    __shared__ uint32_t s_px[2048];
    const int fi = blockIdx.x, lane = threadIdx.x;
    const Geo& g = w.g;
    const int frame = w.hdr[fi].frame;
    const uint8_t* f = w.frames + (size_t)frame * g.frame_bytes;
    uint32_t* ev = w.events + w.frame_ev_off[frame];
    uint32_t prev = px_lin(f, g, 0);
    if (lane == 0) emit_literal_i(ev, prev, 0, false);
    ev += 3;
    int n = 1;
    for (int k0 = 1; k0 <= g.X; k0 += 2048) {
        const int cnt = min(2048, g.X + 1 - k0);
        for (int i = lane; i < cnt; i += 32) s_px[i] = px_lin(f, g, k0 + i);
        __syncwarp();
        if (lane == 0) {
            for (int i = 0; i < cnt; i++) {
                const uint32_t c = s_px[i];
                if (c == prev && n < 255)
                    n++;
                else {
                    *ev++ = make_ev(CX_NTAB + 0, n);
                    emit_literal_i(ev, c, prev, true);
                    ev += 3;
                    n = 1;
                }
                prev = c;
            }
        }
        __syncwarp();
    }
    if (lane == 0) *ev = make_ev(CX_NTAB + 0, n);
}

// k_i_emit: events of the runs of one chunk.  grid = (nchunks, n_iframes), 256 threads, 8 runs each.
__global__ void __launch_bounds__(256) k_i_emit(IWork w) {
    __shared__ uint32_t s_warp[8];
    const Geo& g = w.g;
    const int c = blockIdx.x, fi = blockIdx.y;
    const int frame = w.hdr[fi].frame;
    const uint8_t* f = w.frames + (size_t)frame * g.frame_bytes;
    const long total = (long)g.X * g.Y;
    const long q0 = (long)g.X + 1 + (long)c * ICHUNK;
    const size_t ch = (size_t)fi * w.nchunks + c;
    const uint32_t* cc = w.chunk_cnt + ch * 4;
    const int n = (int)cc[0];
    if (n == 0) return;
    const uint16_t* starts = w.starts + ch * ICHUNK;
    const uint16_t* desc = w.desc + (size_t)fi * total;
    uint32_t* ev = w.events + w.frame_ev_off[frame] + cc[3];
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    const uint32_t carry_type = c > 0 ? w.chunk_cnt[(ch - 1) * 4 + 2] : 0u;  // lastptype starts at 0 (screencap.cpp:346)
    uint32_t base = 0;
    for (int k0 = 0; k0 < n; k0 += 2048) {
        uint32_t d[8];
        uint32_t cnt = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = k0 + threadIdx.x * 8 + j;
            d[j] = k < n ? desc[q0 + starts[k]] : 0xFFFFu;
            if (k < n) cnt += (d[j] >> 8) ? 2u : 5u;
        }
        uint32_t inc = cnt;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, dd);
            if (lane >= dd) inc += t;
        }
        if (lane == 31) s_warp[wi] = inc;
        __syncthreads();
        uint32_t off = base + inc - cnt;
        for (int j = 0; j < wi; j++) off += s_warp[j];
        uint32_t tot = 0;
        for (int j = 0; j < 8; j++) tot += s_warp[j];
        // type of the run before this thread's first run
        const int kfirst = k0 + threadIdx.x * 8;
        uint32_t prev_type = carry_type;
        if (kfirst > 0 && kfirst <= n) prev_type = desc[q0 + starts[kfirst - 1]] >> 8;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int k = kfirst + j;
            if (k < n) {
                const uint32_t type = d[j] >> 8, len = d[j] & 255;
                uint32_t* e = ev + off;
                *e++ = make_ev(CX_PTYPE + (int)prev_type, type);
                if (type == 0) {
                    const long q = q0 + starts[k];
                    emit_literal_i(e, px_lin(f, g, q), px_lin(f, g, q - 1), true);
                    e += 3;
                }
                *e = make_ev(CX_NTAB + (int)type, len);
                off += type ? 2u : 5u;
                prev_type = type;
            }
        }
        base += tot;
        __syncthreads();
    }
}

void launch_i_stage_a(const IWork& w, cudaStream_t st, uint64_t* launches) {
    if (w.n_iframes == 0) return;
    dim3 grid(w.nchunks, w.n_iframes);
    k_i_classify<<<grid, 256, 0, st>>>(w);
    if (w.tm) w.tm->mark("i_classify");
    k_i_entries<<<(w.n_iframes + 31) / 32, 32, 0, st>>>(w);
    if (w.tm) w.tm->mark("i_entries");
    k_i_mark<<<grid, 256, 0, st>>>(w);
    if (w.tm) w.tm->mark("i_mark");
    k_i_offsets<<<w.n_iframes, 32, 0, st>>>(w);
    *launches += 4;
}

void launch_i_emit(const IWork& w, cudaStream_t st, uint64_t* launches) {
    if (w.n_iframes == 0) return;
    dim3 grid(w.nchunks, w.n_iframes);
    k_i_emit_hdr<<<w.n_iframes, 32, 0, st>>>(w);
    k_i_emit<<<grid, 256, 0, st>>>(w);
    *launches += 2;
}

}  // namespace scpr
