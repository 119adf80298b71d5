// decode.cu -- the decoder: one warp per GOP chain for the serial part, a streaming fill kernel for
// everything that is a copy.
//
// Replaces (reference):
//   ScreenCodec::DecompressFrame                 screencap.cpp:1695-1743 (RGB24 -> RGB32 repack fused away)
//   CScreenCapt::DecompressFrame / I / P         screencap.cpp:1522-1557, 414-498, 1275-1432
//   UseANS::decodeC / decodeF / decodeBool       screencap.h:318-359, 411-421
//   Context::decode / update                     ans_contexts.cpp:52-74
//   RansDecInit / Get / Advance                  rans_byte.h:105-146
//
// A GOP is one dependency chain (model state, rANS state and reconstructed pixels interleave,
// SURVEY.md 0.4), so the entropy walk of a chain is serial and runs on one warp; chains of a clip
// run concurrently.  What the reference spends most of a sparse P frame on -- copying unchanged
// blocks from the previous frame and repacking RGB24 to RGB32 -- is not done by that warp at all:
// the chain only writes the blocks a frame changes (decoded in a shared-memory tile), tracks for
// every block which frame holds its current pixels, and k_dec_fill afterwards gathers every
// untouched block of every frame from that source in one HBM-bound pass.
#include <stdio.h>
#include <string.h>

#include <vector>

#include "codec.h"
#include "models.cuh"

namespace scpr {

enum : uint8_t { DK_FLAT = 0, DK_I = 1, DK_PSAME = 2, DK_P = 3 };

struct DecFrame {
    uint32_t src_off;   // offset of the frame's bytes in the device stream copy
    uint32_t size;
    uint8_t kind;       // DK_*
    uint8_t renew;      // flat frame that resets the models (screencap.cpp:1547-1550)
    uint8_t pad[2];
    uint32_t flat_clr;  // colour of a flat frame
};

struct DecChain {
    int first, count;   // frames [first, first+count) of the clip
    int state;          // model state slot
};

struct DecWork {
    Geo g;              // geometry of the OUTPUT frames (pitch = caller's pitch)
    const uint8_t* stream;
    const DecFrame* frames;
    const DecChain* chains;
    uint8_t* states;
    int f0;
    uint8_t* out;       // n output frames
    const uint8_t* prev0;   // frame decoded last by the previous call (output format, pitch g.pitch)
    int* src_cur;       // per chain: nb ints, frame that holds each block's current pixels (-1 = prev0)
    int* src_prev;      // per chain: same, as of the previous frame
    int* stamp;         // per chain: frame in which src_cur was last changed
    uint8_t* upd;       // n * nb flags: block written by the chain kernel in this frame
    int16_t* fill_src;  // n * nb: source frame of every block (k_dec_sources)
    int n;
};

// ---- pixel access through the block-source map ----------------------------------------------------
struct ChainCtx {
    const DecWork* w;
    int* src_cur; int* src_prev; int* stamp;
    int f;              // frame being decoded
};

__device__ __forceinline__ uint32_t frame_px(const DecWork& w, int src, int x, int y) {
    if (src >= 0 && w.frames[src].kind == DK_FLAT) return w.frames[src].flat_clr;
    const uint8_t* base = src >= 0 ? w.out + (size_t)src * w.g.frame_bytes : w.prev0;
    return load_px(base, w.g, x, y);
}
struct PixSrc {
    const uint8_t* base;
    uint32_t clr;
    bool flat;
};
__device__ __forceinline__ PixSrc resolve_src(const DecWork& w, int src) {
    PixSrc s;
    s.flat = src >= 0 && w.frames[src].kind == DK_FLAT;
    s.clr = s.flat ? w.frames[src].flat_clr : 0u;
    s.base = src >= 0 ? w.out + (size_t)src * w.g.frame_bytes : w.prev0;
    return s;
}
__device__ __forceinline__ uint32_t src_px(const PixSrc& s, const Geo& g, int x, int y) { return s.flat ? s.clr : load_px(s.base, g, x, y); }
// pixel of the frame being decoded (only valid for pixels already reconstructed or unchanged)
__device__ __forceinline__ uint32_t cur_px(const ChainCtx& c, int x, int y) {
    const int b = (y >> 4) * c.w->g.nbx + (x >> 4);
    return frame_px(*c.w, c.src_cur[b], x, y);
}
// pixel of the previous frame
__device__ __forceinline__ uint32_t prev_px(const ChainCtx& c, int x, int y) {
    const int b = (y >> 4) * c.w->g.nbx + (x >> 4);
    const int s = c.stamp[b] == c.f ? c.src_prev[b] : c.src_cur[b];
    return frame_px(*c.w, s, x, y);
}
__device__ __forceinline__ void store_px(uint8_t* frame, const Geo& g, int x, int y, uint32_t v) {
    uint8_t* p = frame + (uint32_t)(y * g.pitch + x * g.bpp);
    if (g.bpp == 4)
        *reinterpret_cast<uint32_t*>(p) = v | 0xFF000000u;  // alpha := 255, screencap.cpp:1721
    else {
        p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16);
    }
}

// ---- entropy state ------------------------------------------------------------------------------------
// The rANS state, the byte cursor and the colour-context registers are held identically by all 32
// lanes (every lane executes the same arithmetic on the same values), so no broadcast is needed
// between symbols; only model *stores* are done by one lane.
//
// With one resident warp per chain the decoder is bound by the latency of its dependent
// instruction chain, so the per-symbol path is kept to two shared-memory loads:
//   * the 21 fixed tables live in shared memory for the whole chain: fc[] = freq<<16|cum per symbol,
//     cnt[] the adaptive counters, and -- because a table's intervals are frozen between rescales
//     (ans_contexts.h:1070-1091) -- a direct slot -> symbol map lut[4096] per table that is rebuilt
//     together with the table (every ~128 symbols of that table, by the whole warp).  A symbol is
//     then  v = x & 4095;  sym = lut[v];  fc = fc[sym];  x = freq*(x>>12) + v - cum.
//   * "events until the next rescale" is a countdown per table (the rescale instant depends only on
//     the number of events, each adds 16 to cntsum).
constexpr int FX_TOTAL = 3192;
constexpr int LUT_ROW = 32 * 132;  // 128 slots per lane + 4 bytes padding: conflict-free fills
constexpr int N_LUT = 19;          // every table except the two 512-symbol MV tables
__host__ __device__ constexpr int fx_off(int t) {
    return t < 8 ? t * 256 : t == 8 ? 2048 : t < 13 ? 2056 + (t - 9) * 16 : t < 15 ? 2120 + (t - 13) * 512 : 3144 + (t - 15) * 8;
}
__host__ __device__ constexpr int fx_lut(int t) { return t < 13 ? t : t - 2; }
struct FixedSmem {
    uint32_t fc[FX_TOTAL];
    uint16_t cnt[FX_TOTAL];
    int left[24];
    uint8_t lut[N_LUT][LUT_ROW];
};

#ifdef SCPR_PROF
#define PROF_T0 const long long t0__ = clock64();
#define PROF_ADD(field) e.field += clock64() - t0__;
#define PROF_CNT(field) e.field++;
#else
#define PROF_T0
#define PROF_ADD(field)
#define PROF_CNT(field)
#endif

struct Ent {
#ifdef SCPR_PROF
    long long c_fixed = 0, n_fixed = 0, c_color = 0, n_color = 0, c_tile = 0, n_blocks = 0, c_blkwr = 0, c_mv = 0, c_runs = 0, c_ifill = 0,
              c_hdr = 0, c_total = 0;
#endif
    const uint8_t* p;
    uint32_t x;
    int ndec;
    uint32_t cx, cx1;
    ModelState* m;
    FixedSmem* fs;
    int f0;
    int lane;
};
__device__ __forceinline__ void rdec_init(Ent& e) {  // RansDecInit
    e.x = (uint32_t)e.p[0] | ((uint32_t)e.p[1] << 8) | ((uint32_t)e.p[2] << 16) | ((uint32_t)e.p[3] << 24);
    e.p += 4;
}
__device__ __forceinline__ void rdec_count(Ent& e) {  // re-init every 131072 symbols (screencap.h:327-331)
    if (++e.ndec == RANS_BLOCK) {
        rdec_init(e);
        e.ndec = 0;
    }
}
__device__ __forceinline__ void rdec_advance(Ent& e, uint32_t start, uint32_t freq) {  // RansDecAdvance
    uint32_t x = e.x;
    x = freq * (x >> PROB_BITS) + (x & (PROB_SCALE - 1)) - start;
    while (x < RANS_L) x = (x << 8) | *e.p++;
    e.x = x;
}

// (Re)build the derived data of fixed table t from its counters: `rescale` applies the
// FixedSizeRansCtx rescale (freq := cnt, cum := prefix, cnt -= freq>>1, ans_contexts.h:1075-1090);
// without it the intervals in fc[] are kept (table just loaded).  Then the countdown and the
// slot -> symbol map.  Whole warp.
__device__ __noinline__ void fixed_rebuild(FixedSmem& fs, int t, int lane, bool rescale) {
    const int off = fx_off(t), nsym = fixed_nsym(CX_NTAB + t);
    const int per = (nsym + 31) >> 5, b = lane * per;
    uint32_t ns = 0;
    if (rescale) {
        uint32_t sum = 0;
        for (int j = 0; j < per; j++)
            if (b + j < nsym) sum += fs.cnt[off + b + j];
        uint32_t inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += u;
        }
        uint32_t cf = inc - sum;
        for (int j = 0; j < per; j++)
            if (b + j < nsym) {
                const uint32_t fr = fs.cnt[off + b + j];
                fs.fc[off + b + j] = (fr << 16) | cf;
                cf += fr;
                const uint32_t nc = fr - (fr >> 1);
                fs.cnt[off + b + j] = (uint16_t)nc;
                ns += nc;
            }
    } else {
        for (int j = 0; j < per; j++)
            if (b + j < nsym) ns += fs.cnt[off + b + j];
    }
    const int cntsum = (int)__reduce_add_sync(0xFFFFFFFFu, ns);
    if (lane == 0) fs.left[t] = (PROB_SCALE - 16 - cntsum) / 16 + 1;  // events until cntsum + 16 > 4096
    __syncwarp();
    if (nsym > 256) return;
    // slot -> symbol map: lane owns slots [128*lane, 128*lane + 128)
    uint8_t* row = fs.lut[fx_lut(t)] + lane * 132;
    const int s0 = lane * 128;
    int lo = 0, hi = nsym - 1;  // last symbol whose cum <= s0
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if ((int)(fs.fc[off + mid] & 0xFFFFu) <= s0) lo = mid; else hi = mid - 1;
    }
    int j = lo;
    uint32_t fcj = fs.fc[off + j];
    int endj = (int)(fcj & 0xFFFFu) + (int)(fcj >> 16);
    for (int wd = 0; wd < 32; wd++) {
        const int s = s0 + 4 * wd;
        uint32_t word;
        if (s + 4 <= endj || j == nsym - 1)
            word = (uint32_t)j * 0x01010101u;
        else {
            word = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                while (s + k >= endj && j < nsym - 1) {
                    j++;
                    fcj = fs.fc[off + j];
                    endj = (int)(fcj & 0xFFFFu) + (int)(fcj >> 16);
                }
                word |= (uint32_t)j << (8 * k);
            }
        }
        *reinterpret_cast<uint32_t*>(row + 4 * wd) = word;
    }
    __syncwarp();
}

// decodeF (screencap.h:346-359) for table t (0..20) with NSYM symbols; off = fx_off(t), row = fx_lut(t)
template <int NSYM>
__device__ __forceinline__ int dec_fx(Ent& e, int t, int off, int row) {
    PROF_T0
    FixedSmem& fs = *e.fs;
    const uint32_t v = e.x & (PROB_SCALE - 1);
    int j;
    if (NSYM <= 256) {
        j = fs.lut[row][v + ((v >> 7) << 2)];
    } else {  // 512 symbols: two ballot levels over the cumulative frequencies
        const bool le = (fs.fc[off + e.lane * 16] & 0xFFFFu) <= v;
        const int L = 31 - __clz(__ballot_sync(0xFFFFFFFFu, le));
        const bool le2 = e.lane < 16 && (fs.fc[off + L * 16 + (e.lane & 15)] & 0xFFFFu) <= v;
        j = L * 16 + __popc(__ballot_sync(0xFFFFFFFFu, le2)) - 1;
    }
    const uint32_t fc = fs.fc[off + j];
    const int left = fs.left[t] - 1;
    // every lane of the (converged) warp performs the same read-modify-write with the same values:
    // no lane predicate, no divergence, no barrier on the per-symbol path
    fs.cnt[off + j] = (uint16_t)(fs.cnt[off + j] + 16);
    fs.left[t] = left;
    rdec_advance(e, fc & 0xFFFFu, fc >> 16);
    rdec_count(e);
    if (left == 0) {
        __syncwarp();
        fixed_rebuild(fs, t, e.lane, true);
    }
    PROF_ADD(c_fixed) PROF_CNT(n_fixed)
    return j;
}
template <int NSYM, int T>
__device__ __forceinline__ int dec_fxc(Ent& e) { return dec_fx<NSYM>(e, T, fx_off(T), fx_lut(T)); }
__device__ __forceinline__ int dec_n(Ent& e, int ptype) { return dec_fx<256>(e, ptype, ptype << 8, ptype); }
__device__ __forceinline__ int dec_ptype(Ent& e, int last) { return dec_fx<6>(e, 15 + last, 3144 + 8 * last, 13 + last); }

// ---- colour contexts (global memory, L1 resident working set) --------------------------------------------
// find: uniform (all lanes); update: lane 0.
__device__ __forceinline__ int dec_color(Ent& e, int id) {  // decodeC, screencap.h:318-333
    PROF_T0
    ColorState& x = e.m->color[id];
    const int kind = x.kind;
    int c;
    if (kind >= 6) {  // flat table: two ballot levels over cum[256]
        const uint32_t v = e.x & (PROB_SCALE - 1);
        const bool le = x.cum[e.lane * 8] <= v;
        const int L = 31 - __clz(__ballot_sync(0xFFFFFFFFu, le));
        const bool le2 = e.lane < 8 && x.cum[L * 8 + (e.lane & 7)] <= v;
        c = L * 8 + __popc(__ballot_sync(0xFFFFFFFFu, le2)) - 1;
        const uint32_t freq = x.freq[c], cum = x.cum[c];
        const int cn = x.cnt[c], cs = x.cntsum;
        const int step = kind == 7 ? 16 : (25 << x.fshift);
        if (cn != 0 && cs + 2 * step <= PROB_SCALE) {  // symbol already met, no rescale: count it (all lanes, same values)
            x.cnt[c] = (uint16_t)(cn + step);
            x.cntsum = cs + step;
        } else {
            if (e.lane == 0) cc_encode_counted(x, c);  // new symbol, promotion or rescale
            __syncwarp();
        }
        rdec_advance(e, cum, freq);
    } else if (kind >= 4) {  // SmallContext: walk the (<= 16) sorted symbols, ans_contexts.h:238-283
        const uint32_t v0 = e.x & (PROB_SCALE - 1);
        // header + symbols + frequencies: four independent 128-bit loads, then registers only
        const uint4 hd = *reinterpret_cast<const uint4*>(&x.kind);
        const uint4 sy4 = *reinterpret_cast<const uint4*>(x.ssym);
        const uint4 f0 = *reinterpret_cast<const uint4*>(x.sfreq);
        const uint4 f1 = *reinterpret_cast<const uint4*>(x.sfreq + 8);
        const int maxpos = (hd.x >> 16) & 255, d = hd.y & 0xFFFF;
        const uint32_t syw[4] = {sy4.x, sy4.y, sy4.z, sy4.w};
        const uint32_t frw[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
        int totFr = (int)hd.z;
        if (kind == 4) totFr = 256 - d + (int)((f0.x & 0xFFFF) + (f0.x >> 16) + (f0.y & 0xFFFF) + (f0.y >> 16));  // :303
        int shift = 0, tot = totFr;
        while (tot <= PROB_SCALE / 2) {
            tot <<= 1;
            shift++;
        }
        const int v = (int)(v0 >> shift);
        const int bonus = (PROB_SCALE - tot) >> shift;
        int cumFr = 0, lastSymb = 0, pos = 0, fr = 1, fpos = 0;
        bool found = false;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (k >= d) break;  // uniform: leaves the unrolled chain as soon as the symbol is placed
            const int fk = (int)((frw[k >> 1] >> (16 * (k & 1))) & 0xFFFF);
            const int sk = (int)((syw[k >> 2] >> (8 * (k & 3))) & 255);
            const int startFr = cumFr + sk - lastSymb;
            if (v < startFr) break;
            const int frk = (fk + (k == maxpos ? bonus : 0)) & 0xFFFF;
            if (startFr + frk > v) {
                c = sk;
                cumFr = startFr;
                fr = frk;
                fpos = fk;
                found = true;
                break;
            }
            cumFr = startFr + frk;
            lastSymb = sk + 1;
            pos = k + 1;
        }
        if (!found) {  // a symbol not met yet: width 1
            c = lastSymb + v - cumFr;
            cumFr = v;
            fr = 1;
        }
        if (found && totFr + 100 <= PROB_SCALE) {  // hot path: count, no rescale (ans_contexts.h:211-214); all lanes, same values
            const int nf = fpos + 50;
            x.sfreq[pos] = (uint16_t)nf;
            if (kind == 5) x.cntsum = totFr + 50;
            if (pos != maxpos && nf > x.sfreq[maxpos]) x.maxpos = (uint8_t)pos;  // sfreq[maxpos] is not the entry just written
        } else {
            if (e.lane == 0) cc_encode_counted(x, c);  // new symbol, promotion or rescale: the general path
            __syncwarp();
        }
        rdec_advance(e, (uint32_t)(cumFr << shift) & 0xFFFFu, (uint32_t)(fr << shift) & 0xFFFFu);
    } else {
        c = *e.p++;
        if (e.lane == 0) cc_update_raw(x, c, e.f0);
        __syncwarp();  // lane 0's model stores are visible to the whole warp before the next symbol
    }
    rdec_count(e);
    PROF_ADD(c_color) PROF_CNT(n_color)
    return c;
}
__device__ inline int dec_bool(Ent& e) {  // decodeBool
    const int flag = (e.x & (PROB_SCALE - 1)) >= PROB_SCALE / 2;
    rdec_advance(e, flag ? PROB_SCALE / 2 : 0, PROB_SCALE / 2);
    rdec_count(e);
    return flag;
}
__device__ __forceinline__ uint32_t dec_rgb(Ent& e) {  // DecodeRGB, screencap.cpp:662-679
    uint32_t px = 0;
#pragma unroll 1
    for (int ch = 0; ch < 3; ch++) {  // one copy of the colour decoder in the instruction stream
        const uint32_t v = (uint32_t)dec_color(e, ch * 4096 + (int)(e.cx + e.cx1));
        e.cx1 = (e.cx << 6) & 0xFC0;
        e.cx = v >> 2;
        px |= v << (8 * ch);
    }
    return px;
}
__device__ __forceinline__ void set_cx_from(Ent& e, uint32_t px) {  // screencap.cpp:488-493, 1417-1419
    e.cx = ((px >> 8) & 255) >> 2;
    e.cx1 = (e.cx << 6) & 0xFC0;
    e.cx = ((px >> 16) & 255) >> 2;
}

__device__ __forceinline__ uint32_t grad_px(uint32_t l, uint32_t t, uint32_t tl) {  // type 4, truncated to bytes
    uint32_t v = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const int sh = 8 * c;
        v |= (uint32_t)(((int)((l >> sh) & 255) + (int)((t >> sh) & 255) - (int)((tl >> sh) & 255)) & 255) << sh;
    }
    return v;
}

// ---- I frame (DecompressI, screencap.cpp:414-498) --------------------------------------------------
// Runs are laid along the raster; all neighbours are pixels of this frame: left = the previous
// pixel in raster order (lasti), top = (x, y-1), top-left = byte offset -stride-3, i.e. (x-1, y-1),
// which for x == 0 is the tail of row y-2 (with row padding it straddles padding bytes, A.2).
struct IPos {
    int x, y;
};
__device__ __forceinline__ IPos ipos_add(IPos p, int i, int X) {
    p.x += i;
    while (p.x >= X) {
        p.x -= X;
        p.y++;
    }
    return p;
}
__device__ __forceinline__ IPos ipos_prev(IPos p, int X) {
    if (p.x > 0) return IPos{p.x - 1, p.y};
    return IPos{X - 1, p.y - 1};
}
__device__ __noinline__ uint32_t tl_padded(const uint8_t* frame, const Geo& g, int y) {
    // x == 0 with row padding: bytes (y-1)*stride24 - 3 .. of the RGB24 view, padding reads as 0
    const int stride24 = (g.X * 3 + 3) & ~3;
    uint32_t v = 0;
    const int o = y * stride24 - stride24 - 3;
    for (int k = 0; k < 3; k++) {
        const int ok = o + k;
        const int row = ok / stride24, col = ok % stride24;
        uint32_t b = 0;
        if (col < 3 * g.X) b = (load_px(frame, g, col / 3, row) >> (8 * (col % 3))) & 255;
        v |= b << (8 * k);
    }
    return v;
}
__device__ __forceinline__ uint32_t tl_at(const uint8_t* frame, const Geo& g, IPos p, bool padded) {
    if (p.x > 0) return load_px(frame, g, p.x - 1, p.y - 1);
    if (!padded) return load_px(frame, g, g.X - 1, p.y - 2);
    return tl_padded(frame, g, p.y);
}

__device__ void decode_i(const DecWork& w, Ent& e, uint8_t* frame, int lane) {
    const Geo& g = w.g;
    const int X = g.X, Y = g.Y;
    const bool padded = ((X * 3 + 3) & ~3) != X * 3;
    IPos p{0, 0};
    int ptype = 0;
    int hdr = X + 1;
    // first row and one pixel: (rgb, n) pairs, lengths in ntab[0] (screencap.cpp:423-438)
    while (hdr > 0) {
        const uint32_t c = dec_rgb(e);
        const int n = dec_n(e, 0);
        if (n <= 0) return;
        for (int i = lane; i < n; i += 32) {
            const IPos q = ipos_add(p, i, X);
            if (q.y < Y) store_px(frame, g, q.x, q.y, c);
        }
        p = ipos_add(p, n, X);
        hdr -= n;
        __syncwarp();
    }
    while (p.y < Y) {
        uint32_t c = 0;
        ptype = dec_ptype(e, ptype);
        if (!ptype) c = dec_rgb(e);
        const int n = dec_n(e, ptype);
        if (n <= 0) return;
        PROF_T0
        if (ptype == 0 || ptype == 1) {
            if (ptype == 1) {
                const IPos l = ipos_prev(p, X);
                c = load_px(frame, g, l.x, l.y);
            }
            for (int i = lane; i < n; i += 32) {
                const IPos q = ipos_add(p, i, X);
                if (q.y < Y) store_px(frame, g, q.x, q.y, c);
            }
        } else if (n < X && (ptype == 2 || ptype == 5)) {  // sources lie strictly before the run
            for (int i = lane; i < n; i += 32) {
                const IPos q = ipos_add(p, i, X);
                if (q.y < Y) store_px(frame, g, q.x, q.y, ptype == 2 ? load_px(frame, g, q.x, q.y - 1) : tl_at(frame, g, q, padded));
            }
        } else {  // gradient chains through the left pixel; tiny frames may read their own run
            if (lane == 0) {
                IPos q = p;
                for (int i = 0; i < n && q.y < Y; i++) {
                    uint32_t v;
                    if (ptype == 2) v = load_px(frame, g, q.x, q.y - 1);
                    else if (ptype == 5) v = tl_at(frame, g, q, padded);
                    else {
                        const IPos l = ipos_prev(q, X);
                        v = grad_px(load_px(frame, g, l.x, l.y), load_px(frame, g, q.x, q.y - 1), tl_at(frame, g, q, padded));
                    }
                    store_px(frame, g, q.x, q.y, v);
                    q = ipos_add(q, 1, X);
                }
            }
        }
        __syncwarp();
        p = ipos_add(p, n, X);
        const IPos l = ipos_prev(p, X);
        set_cx_from(e, load_px(frame, g, l.x, l.y));
        PROF_ADD(c_ifill)
    }
}

// ---- P frame (DecompressP, screencap.cpp:1275-1432) ------------------------------------------------
__device__ void decode_p(const DecWork& w, ChainCtx& cc, Ent& e, uint8_t* frame, int f, uint8_t* s_bts, uint32_t (*tile)[17],
                         int lane) {
    const Geo& g = w.g;
    const long long thdr__ = clock64();
    int t0 = dec_fxc<256, CX_XX - CX_NTAB>(e);
    const int xx1 = (dec_fxc<256, CX_XX - CX_NTAB>(e) << 8) + t0;
    t0 = dec_fxc<256, CX_XX - CX_NTAB>(e);
    int xx2 = (dec_fxc<256, CX_XX - CX_NTAB>(e) << 8) + t0;
    if (xx2 >= g.nb) xx2 = g.nb - 1;  // corrupt input guard
    // block types of [xx1, xx2] as (type, run) pairs (screencap.cpp:1306-1313)
    for (int x = xx1; x <= xx2;) {
        const int c = dec_fxc<5, CX_BT - CX_NTAB>(e);
        const int n = dec_fxc<256, CX_NTAB2 - CX_NTAB>(e);
        if (n <= 0) break;
        for (int i = lane; i < n && x + i < g.nb; i += 32) s_bts[x + i] = (uint8_t)c;
        x += n;
    }
    __syncwarp();
#ifdef SCPR_PROF
    e.c_hdr += clock64() - thdr__;
#endif
    e.cx = e.cx1 = 0;
    int lastmx = 0, lastmy = 0;
    uint8_t* upd = w.upd + (size_t)f * g.nb;
    // visit changed blocks only: 32 block types per step, ballot, iterate the set bits
    for (int b0 = xx1 & ~31; b0 <= xx2; b0 += 32) {
      const int myb = b0 + lane;
      uint32_t chm = __ballot_sync(0xFFFFFFFFu, myb >= xx1 && myb <= xx2 && s_bts[myb] != 0);
      while (chm) {
        const int bi = b0 + __ffs(chm) - 1;
        chm &= chm - 1;
        const int bt = s_bts[bi];
        const int by = bi / g.nbx, bx = bi - by * g.nbx;
        const int bx0 = bx * 16, by0 = by * 16;
        const int bw = min(16, g.X - bx0), bh = min(16, g.Y - by0);
        int x1 = bx0, y1 = by0, x2 = bx0 + bw, y2 = by0 + bh;
        PROF_CNT(n_blocks)
        if ((bt - 1) & 2) {
            // ---- motion-vector block: a pure gather from the previous frame, written straight to the
            // frame (no tile, no neighbours): sub-rect from prev at (x+mx, y+my), the rest of a partial
            // block from prev at the same place (screencap.cpp:1333-1368)
            PROF_T0
            if ((bt - 1) & 1) {
                x1 = bx0 + dec_fxc<16, CX_SXY - CX_NTAB + 0>(e);
                y1 = by0 + dec_fxc<16, CX_SXY - CX_NTAB + 1>(e);
                x2 = bx0 + dec_fxc<16, CX_SXY - CX_NTAB + 2>(e) + 1;
                y2 = by0 + dec_fxc<16, CX_SXY - CX_NTAB + 3>(e) + 1;
                if (x2 > bx0 + bw) x2 = bx0 + bw;  // corrupt input guards
                if (y2 > by0 + bh) y2 = by0 + bh;
                if (x1 >= x2) x1 = x2 - 1;
                if (y1 >= y2) y1 = y2 - 1;
            }
            int mx = lastmx, my = lastmy;
            if (!dec_bool(e)) {
                mx = dec_fxc<512, CX_MV - CX_NTAB + 0>(e) - 256;
                my = dec_fxc<512, CX_MV - CX_NTAB + 1>(e) - 256;
            }
            lastmx = mx; lastmy = my;
            const int gx = min(max(x1 + mx, 0), g.X - 1), gy = min(max(y1 + my, 0), g.Y - 1);  // corrupt input guard
            const int cbx = gx >> 4, cby = gy >> 4;
            PixSrc sq[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {  // the source rectangle spans at most 2 x 2 blocks of the previous frame
                const int qx = min(cbx + (q & 1), g.nbx - 1), qy = min(cby + (q >> 1), g.nby - 1);
                const int b = qy * g.nbx + qx;
                sq[q] = resolve_src(w, cc.stamp[b] == f ? cc.src_prev[b] : cc.src_cur[b]);
            }
            const PixSrc sP = resolve_src(w, cc.stamp[bi] == f ? cc.src_prev[bi] : cc.src_cur[bi]);
            uint32_t px[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {  // all loads first (a source block may be this very block)
                const int p = lane + 32 * u, xx = p & 15, yy = p >> 4;
                const int x = bx0 + xx, y = by0 + yy;
                px[u] = 0;
                if (xx < bw && yy < bh) {
                    if (x >= x1 && x < x2 && y >= y1 && y < y2) {
                        const int sx = min(max(x + mx, 0), g.X - 1), sy = min(max(y + my, 0), g.Y - 1);
                        const int q = ((sx >> 4) > cbx ? 1 : 0) + ((sy >> 4) > cby ? 2 : 0);
                        px[u] = src_px(q == 0 ? sq[0] : q == 1 ? sq[1] : q == 2 ? sq[2] : sq[3], g, sx, sy);
                    } else
                        px[u] = src_px(sP, g, x, y);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int p = lane + 32 * u, xx = p & 15, yy = p >> 4;
                if (xx < bw && yy < bh) store_px(frame, g, bx0 + xx, by0 + yy, px[u]);
            }
            if (lane == 0) {
                if (cc.stamp[bi] != f) {
                    cc.src_prev[bi] = cc.src_cur[bi];
                    cc.stamp[bi] = f;
                }
                cc.src_cur[bi] = f;
                upd[bi] = 1;
            }
            __syncwarp();
            PROF_ADD(c_mv)
            continue;
        }
        // ---- pixel-coded block, decoded in a shared-memory tile ----
        // tile[1+yy][1+xx] = block pixel; row 0 / column 0 = the neighbours above / left (current frame)
        { PROF_T0
        {   // the tile touches four blocks: resolve where each one's pixels live once, then load
            const int bA = bi - g.nbx - 1, bT = bi - g.nbx, bL = bi - 1;
            const PixSrc sA = resolve_src(w, (bx > 0 && by > 0) ? cc.src_cur[bA] : -1);
            const PixSrc sT = resolve_src(w, by > 0 ? cc.src_cur[bT] : -1);
            const PixSrc sL = resolve_src(w, bx > 0 ? cc.src_cur[bL] : -1);
            const PixSrc sP = resolve_src(w, cc.stamp[bi] == f ? cc.src_prev[bi] : cc.src_cur[bi]);
            const bool partial = (bt - 1) & 1;
            if (partial) {  // block starts as a copy of the previous frame's block
                for (int p = lane; p < 17 * 17; p += 32) {
                    const int ty = p / 17, tx = p - ty * 17;
                    const int x = bx0 + tx - 1, y = by0 + ty - 1;
                    uint32_t v = 0;
                    if (x >= 0 && y >= 0 && x < g.X && y < g.Y) v = src_px(ty == 0 ? (tx == 0 ? sA : sT) : tx == 0 ? sL : sP, g, x, y);
                    tile[ty][tx] = v;
                }
            } else {  // every pixel of the block is coded: only the 33 neighbours are needed
                for (int p = lane; p < 33; p += 32) {
                    const int ty = p < 17 ? 0 : p - 16, tx = p < 17 ? p : 0;
                    const int x = bx0 + tx - 1, y = by0 + ty - 1;
                    uint32_t v = 0;
                    if (x >= 0 && y >= 0 && x < g.X && y < g.Y) v = src_px(ty == 0 ? (tx == 0 ? sA : sT) : sL, g, x, y);
                    tile[ty][tx] = v;
                }
            }
        }
        __syncwarp();
        PROF_ADD(c_tile) }
        if ((bt - 1) & 1) {
            x1 = bx0 + dec_fxc<16, CX_SXY - CX_NTAB + 0>(e);
            y1 = by0 + dec_fxc<16, CX_SXY - CX_NTAB + 1>(e);
            x2 = bx0 + dec_fxc<16, CX_SXY - CX_NTAB + 2>(e) + 1;
            y2 = by0 + dec_fxc<16, CX_SXY - CX_NTAB + 3>(e) + 1;
            if (x2 > bx0 + bw) x2 = bx0 + bw;  // corrupt input guards
            if (y2 > by0 + bh) y2 = by0 + bh;
            if (x1 >= x2) x1 = x2 - 1;
            if (y1 >= y2) y1 = y2 - 1;
        }
        const int sw = x2 - x1, sh = y2 - y1;
        const uint32_t swinv = (65536u + (uint32_t)sw - 1) / (uint32_t)sw;  // idx / sw == (idx * swinv) >> 16 for idx < 256
        {  // pixel runs over the sub-rect in its own raster order
            PROF_T0
            int pos = 0, ptype = 0;
            const int npx = sw * sh;
            const int ox = 1 + x1 - bx0, oy = 1 + y1 - by0;
            while (pos < npx) {
                uint32_t c = 0;
                ptype = dec_ptype(e, ptype);
                if (!ptype) c = dec_rgb(e);
                int n = dec_n(e, ptype);
                if (n > npx - pos) n = npx - pos;
                if (n <= 0) break;
                uint32_t v = c;
                if (ptype == 0 || ptype == 3) {  // no dependence on pixels of this run: lanes fill in parallel
                    for (int i = lane; i < n; i += 32) {
                        const int yy = (int)(((uint32_t)(pos + i) * swinv) >> 16), xx = pos + i - yy * sw;
                        uint32_t* t = &tile[oy + yy][ox + xx];
                        if (ptype == 3) t[0] = (bt - 1) & 1 ? t[0] : prev_px(cc, x1 + xx, y1 + yy);
                        else t[0] = c;
                    }
                    __syncwarp();
                    const int li = pos + n - 1, ly = (int)(((uint32_t)li * swinv) >> 16);
                    v = tile[oy + ly][ox + li - ly * sw];
                } else {
                    // predicted from pixels of this block: one row segment of the sub-rect per step (<= 16
                    // pixels, sources lie in the row above or left of the segment); gradient chains serially
                    int yy = (int)(((uint32_t)pos * swinv) >> 16), xx = pos - yy * sw;
                    for (int rem = n; rem > 0;) {
                        const int seg = min(rem, sw - xx);
                        uint32_t* row = &tile[oy + yy][ox + xx];
                        if (ptype == 4) {
                            if (lane == 0)
                                for (int i = 0; i < seg; i++) row[i] = grad_px(row[i - 1], row[i - 17], row[i - 18]);
                        } else if (lane < seg) {
                            row[lane] = ptype == 1 ? row[-1] : ptype == 2 ? row[lane - 17] : row[lane - 18];
                        }
                        __syncwarp();
                        rem -= seg;
                        xx = 0;
                        yy++;
                    }
                    const int li = pos + n - 1, ly = (int)(((uint32_t)li * swinv) >> 16);
                    v = tile[oy + ly][ox + li - ly * sw];
                }
                set_cx_from(e, v);
                pos += n;
            }
            PROF_ADD(c_runs)
        }
        __syncwarp();
        // write the whole block and hand its ownership to this frame
        { PROF_T0
        if (bw == 16) {
            for (int p = lane; p < 16 * bh; p += 32) store_px(frame, g, bx0 + (p & 15), by0 + (p >> 4), tile[1 + (p >> 4)][1 + (p & 15)]);
        } else {
            for (int p = lane; p < bw * bh; p += 32) {
                const int xx = p % bw, yy = p / bw;
                store_px(frame, g, bx0 + xx, by0 + yy, tile[1 + yy][1 + xx]);
            }
        }
        if (lane == 0) {
            if (cc.stamp[bi] != f) {
                cc.src_prev[bi] = cc.src_cur[bi];
                cc.stamp[bi] = f;
            }
            cc.src_cur[bi] = f;
            upd[bi] = 1;
        }
        __syncwarp();
        PROF_ADD(c_blkwr) }
      }
    }
}

// one warp per chain
__global__ void __launch_bounds__(32) k_dec_chain(DecWork w) {
    extern __shared__ __align__(16) uint8_t s_mem[];
    FixedSmem* fs = reinterpret_cast<FixedSmem*>(s_mem);
    uint32_t(*tile)[17] = reinterpret_cast<uint32_t(*)[17]>(s_mem + sizeof(FixedSmem));
    uint8_t* s_bts = s_mem + sizeof(FixedSmem) + 17 * 17 * 4;
    const int lane = threadIdx.x;
    const DecChain ch = w.chains[blockIdx.x];
    const Geo& g = w.g;
    ChainCtx cc;
    cc.w = &w;
    cc.src_cur = w.src_cur + (size_t)blockIdx.x * g.nb;
    cc.src_prev = w.src_prev + (size_t)blockIdx.x * g.nb;
    cc.stamp = w.stamp + (size_t)blockIdx.x * g.nb;
    for (int i = lane; i < g.nb; i += 32) {
        cc.src_cur[i] = -1;
        cc.src_prev[i] = -1;
        cc.stamp[i] = -1;
    }
    __syncwarp();
    Ent e;
    e.m = reinterpret_cast<ModelState*>(w.states + (size_t)ch.state * sizeof(ModelState));
    e.f0 = w.f0;
    e.cx = e.cx1 = 0;
    e.ndec = 0;
    e.x = 0;
    e.p = nullptr;
    e.fs = fs;
    e.lane = lane;
#ifdef SCPR_PROF
    const long long tk0 = clock64();
#endif
    // fixed tables of the chain's model state -> shared memory (a renewing first frame overwrites them)
    for (int t = 0; t < NUM_FIXED_CX; t++) {
        const FixedState& g0 = e.m->fx[t];
        const int off = fx_off(t), nsym = fixed_nsym(CX_NTAB + t);
        for (int i = lane; i < nsym; i += 32) {
            fs->cnt[off + i] = g0.cnt[i];
            fs->fc[off + i] = ((uint32_t)g0.freq[i] << 16) | g0.cum[i];
        }
    }
    __syncwarp();
    for (int t = 0; t < NUM_FIXED_CX; t++) fixed_rebuild(*fs, t, lane, false);
    for (int f = ch.first; f < ch.first + ch.count; f++) {
        const DecFrame df = w.frames[f];
        uint8_t* frame = w.out + (size_t)f * g.frame_bytes;
        cc.f = f;
        if (df.kind == DK_PSAME) continue;  // every block keeps its source (memcpy(pDst, prev), screencap.cpp:1288-1291)
        if (df.kind == DK_FLAT || df.kind == DK_I) {
            if (df.kind == DK_I || df.renew) {  // RenewI
                for (int i = lane; i < NUM_COLOR_CX; i += 32) e.m->color[i].kind = 0;
                for (int t = 0; t < NUM_FIXED_CX; t++) {  // FixedSizeRansCtx::renew, ans_contexts.h:1114-1131
                    const int off = fx_off(t), nsym = fixed_nsym(CX_NTAB + t);
                    const int fr = PROB_SCALE / nsym, c0 = fr - (fr >> 1);
                    for (int i = lane; i < nsym; i += 32) {
                        fs->cnt[off + i] = (uint16_t)c0;
                        fs->fc[off + i] = ((uint32_t)fr << 16) | (uint32_t)(fr * i);
                    }
                }
                __syncwarp();
                for (int t = 0; t < NUM_FIXED_CX; t++) fixed_rebuild(*fs, t, lane, false);
            }
            for (int i = lane; i < g.nb; i += 32) {  // the whole frame is new
                cc.src_prev[i] = cc.src_cur[i];
                cc.stamp[i] = f;
                cc.src_cur[i] = f;
            }
            __syncwarp();
            if (df.kind == DK_I) {
                e.p = w.stream + df.src_off + 1;
                e.ndec = 0;
                e.cx = e.cx1 = 0;
                rdec_init(e);
                decode_i(w, e, frame, lane);
            }
            continue;
        }
        e.p = w.stream + df.src_off + 1;
        e.ndec = 0;
        rdec_init(e);
        decode_p(w, cc, e, frame, f, s_bts, tile, lane);
        __threadfence_block();
    }
#ifdef SCPR_PROF
    if (lane == 0)
        printf("[dec prof] chain %d frames %d: total %.1f Mcyc | fixed %.1f Mcyc / %lld sym (%.0f cyc) | color %.1f / %lld (%.0f cyc) | "
               "p-hdr %.1f | tile %.1f / %lld blocks | mv %.1f | runs(incl sym) %.1f | blkwr %.1f | ifill %.1f\n",
               (int)blockIdx.x, ch.count, (clock64() - tk0) * 1e-6, e.c_fixed * 1e-6, e.n_fixed, (double)e.c_fixed / (double)max(1LL, e.n_fixed),
               e.c_color * 1e-6, e.n_color, (double)e.c_color / (double)max(1LL, e.n_color), e.c_hdr * 1e-6, e.c_tile * 1e-6, e.n_blocks,
               e.c_mv * 1e-6, e.c_runs * 1e-6, e.c_blkwr * 1e-6, e.c_ifill * 1e-6);
#endif
    // leave the fixed tables behind for the next call
    __syncwarp();
    for (int t = 0; t < NUM_FIXED_CX; t++) {
        FixedState& g0 = e.m->fx[t];
        const int off = fx_off(t), nsym = fixed_nsym(CX_NTAB + t);
        uint32_t sum = 0;
        for (int i = lane; i < nsym; i += 32) {
            g0.cnt[i] = fs->cnt[off + i];
            g0.freq[i] = (uint16_t)(fs->fc[off + i] >> 16);
            g0.cum[i] = (uint16_t)(fs->fc[off + i] & 0xFFFFu);
            sum += fs->cnt[off + i];
        }
        sum = __reduce_add_sync(0xFFFFFFFFu, sum);
        if (lane == 0) {
            g0.cntsum = (int)sum;  // cntsum is the sum of the counters at all times
            g0.nsym = nsym;
        }
    }
}

// source frame of every block of every frame: thread per block, frames in order
__global__ void k_dec_sources(DecWork w) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= w.g.nb) return;
    int last = -1;
    for (int f = 0; f < w.n; f++) {
        const uint8_t k = w.frames[f].kind;
        if (k == DK_I || k == DK_FLAT || w.upd[(size_t)f * w.g.nb + b]) last = f;
        w.fill_src[(size_t)f * w.g.nb + b] = (int16_t)last;
    }
}

// gather every block a frame did not write itself from the frame that holds it (or paint flat
// frames).  One warp per block row strip of 8 blocks, as in the frame scan; 32 bpp fast path uses
// 128-bit accesses.
__global__ void __launch_bounds__(256) k_dec_fill(DecWork w) {
    const Geo& g = w.g;
    const int lane = threadIdx.x & 31;
    const int strips_x = (g.nbx + 7) >> 3;
    const int strips_per_frame = strips_x * g.nby;
    const long strip = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (strip >= (long)w.n * strips_per_frame) return;
    const int f = (int)(strip / strips_per_frame);
    const int s = (int)(strip - (long)f * strips_per_frame);
    const int by = s / strips_x, sx = s - by * strips_x;
    const int bx = sx * 8 + (lane >> 2);
    if (bx >= g.nbx) return;
    const int src = w.fill_src[(size_t)f * g.nb + (size_t)by * g.nbx + bx];
    const bool flat = w.frames[f].kind == DK_FLAT;
    if (src == f && !flat) return;  // written by the chain kernel
    uint8_t* dst = w.out + (size_t)f * g.frame_bytes;
    const int y0 = by * 16, rows = min(16, g.Y - y0);
    const bool src_flat = flat || (src >= 0 && w.frames[src].kind == DK_FLAT);
    const uint32_t clr = flat ? w.frames[f].flat_clr : (src >= 0 ? w.frames[src].flat_clr : 0u);
    const uint8_t* sp = src >= 0 ? w.out + (size_t)src * g.frame_bytes : w.prev0;
    const int x0 = bx * 16 + (lane & 3) * 4;
    if (g.bpp == 4 && (g.pitch & 15) == 0 && (g.X & 3) == 0) {
        if (x0 >= g.X) return;
        const size_t base = (size_t)y0 * g.pitch + (size_t)x0 * 4;
        if (src_flat) {
            const uint32_t v = clr | 0xFF000000u;
            const uint4 q = make_uint4(v, v, v, v);
            for (int r = 0; r < rows; r++) *reinterpret_cast<uint4*>(dst + base + (size_t)r * g.pitch) = q;
        } else {
            uint4 q[16];
#pragma unroll
            for (int r = 0; r < 16; r++)
                if (r < rows) q[r] = *reinterpret_cast<const uint4*>(sp + base + (size_t)r * g.pitch);
#pragma unroll
            for (int r = 0; r < 16; r++)
                if (r < rows) *reinterpret_cast<uint4*>(dst + base + (size_t)r * g.pitch) = q[r];
        }
    } else {
        for (int r = 0; r < rows; r++)
            for (int k = 0; k < 4; k++) {
                const int x = x0 + k;
                if (x < g.X) store_px(dst, g, x, y0 + r, src_flat ? clr : load_px(sp, g, x, y0 + r));
            }
    }
}

}  // namespace scpr

using namespace scpr;
#define CK(call) SCPR_CUDA_CHECK(call)
#define TRY(expr)                 \
    do {                          \
        int r__ = (expr);         \
        if (r__ < 0) return r__;  \
    } while (0)

static int ensure_dec_states(scpr_codec* c, int n) {
    if (n <= c->dec_n_states) return SCPR_OK;
    DBuf nb;
    TRY(nb.ensure((size_t)n * model_state_bytes()));
    if (c->dec_state.p)
        CK(cudaMemcpyAsync((uint8_t*)nb.p + (size_t)c->dec_cur_state * model_state_bytes(),
                           (uint8_t*)c->dec_state.p + (size_t)c->dec_cur_state * model_state_bytes(), model_state_bytes(),
                           cudaMemcpyDeviceToDevice, c->st));
    CK(cudaStreamSynchronize(c->st));
    c->dec_state.release();
    c->dec_state = nb;
    c->dec_n_states = n;
    return SCPR_OK;
}

// n frames, bitstreams on the host, decoded frames to device memory `d_out` (pitch bytes per row)
static int decode_batch(scpr_codec* c, const uint8_t* stream, const uint32_t* sizes, const uint8_t* ftypes, int n, uint8_t* d_out,
                        int pitch) {
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->st;
    if (n <= 0) return 1;
    Geo g = c->g;
    g.pitch = pitch;
    g.frame_bytes = (size_t)pitch * g.Y;
    if (pitch < g.X * g.bpp) return SCPR_E_PARAM;
    // ---- host plan: frame kinds, versions, chains (DecompressFrame, screencap.cpp:1695-1702, 1522-1557)
    std::vector<DecFrame> frames(n);
    std::vector<DecChain> chains;
    uint64_t off = 0;
    for (int f = 0; f < n; f++) {
        DecFrame& d = frames[f];
        memset(&d, 0, sizeof(d));
        d.src_off = (uint32_t)off;
        d.size = sizes[f];
        const uint8_t* s = stream + off;
        off += sizes[f];
        if (sizes[f] < 1) return SCPR_E_PARAM;
        if (!c->dec_created) {
            if (ftypes[f] > 0) return 0;  // P frame before any I frame
            const int version = (s[0] >> 4) + 1;
            if (version != 3 && version != 4) return -version;  // BadVersionException; v2 = range coder, out of scope
            c->dec_version = version;
            c->dec_created = true;
        }
        if (ftypes[f]) {
            c->dec_last_was_flat = false;
            d.kind = (s[0] & 1) ? DK_P : DK_PSAME;
            if (chains.empty()) chains.push_back(DecChain{f, 0, -2});  // continues the previous call's chain
        } else if ((s[0] & 0x0F) == 1) {
            if (sizes[f] < 4) return SCPR_E_PARAM;
            d.kind = DK_FLAT;
            d.flat_clr = (uint32_t)s[1] | ((uint32_t)s[2] << 8) | ((uint32_t)s[3] << 16);
            d.renew = !(c->dec_last_was_flat && !memcmp(c->dec_last_flat_clr, s + 1, 3));
            c->dec_last_was_flat = true;
            memcpy(c->dec_last_flat_clr, s + 1, 3);
            if (d.renew || chains.empty()) chains.push_back(DecChain{f, 0, d.renew ? -1 : -2});
        } else {
            c->dec_last_was_flat = false;
            d.kind = DK_I;
            chains.push_back(DecChain{f, 0, -1});
        }
        chains.back().count = f - chains.back().first + 1;
    }
    if (off >= 0xFFFF0000ull) return SCPR_E_PARAM;
    const int n_chains = (int)chains.size();
    TRY(ensure_dec_states(c, n_chains + 1));
    {
        int next = 0;
        for (auto& ch : chains) {
            if (ch.state == -2)
                ch.state = c->dec_cur_state;
            else {
                if (next == c->dec_cur_state) next++;
                ch.state = next++;
            }
        }
        // note: states of new chains must differ from the persistent one only while it is in use by chain 0
        c->dec_cur_state = chains.back().state;
    }
    // ---- device buffers ---------------------------------------------------------------------------
    TRY(c->dec_stream.ensure((size_t)off + 64));
    TRY(c->dec_desc.ensure((size_t)n * sizeof(DecFrame) + (size_t)n_chains * sizeof(DecChain)));
    const size_t map_bytes = (size_t)n_chains * g.nb * 4;
    TRY(c->dec_ws.ensure(3 * map_bytes + (size_t)n * g.nb * 3 + 64));
    if (c->dec_prev_pitch != pitch) {  // previous frame is kept in output format
        TRY(c->dec_prev.ensure(g.frame_bytes));
        if (c->dec_prev_pitch == 0) CK(cudaMemsetAsync(c->dec_prev.p, 0, g.frame_bytes, st));
        else {
            set_error("output pitch changed between calls");
            return SCPR_E_PARAM;
        }
        c->dec_prev_pitch = pitch;
    }
    CK(cudaMemcpyAsync(c->dec_stream.p, stream, (size_t)off, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync((uint8_t*)c->dec_stream.p + off, 0, 64, st));
    CK(cudaMemcpyAsync(c->dec_desc.p, frames.data(), (size_t)n * sizeof(DecFrame), cudaMemcpyHostToDevice, st));
    DecChain* d_chains = reinterpret_cast<DecChain*>((uint8_t*)c->dec_desc.p + (size_t)n * sizeof(DecFrame));
    CK(cudaMemcpyAsync(d_chains, chains.data(), (size_t)n_chains * sizeof(DecChain), cudaMemcpyHostToDevice, st));
    DecWork w;
    memset(&w, 0, sizeof(w));
    w.g = g;
    w.stream = (const uint8_t*)c->dec_stream.p;
    w.frames = (const DecFrame*)c->dec_desc.p;
    w.chains = d_chains;
    w.states = (uint8_t*)c->dec_state.p;
    w.f0 = c->dec_version == 3 ? 64 : 32;  // setCx6f0, screencap.cpp:1613-1614
    w.out = d_out;
    w.prev0 = (const uint8_t*)c->dec_prev.p;
    uint8_t* ws = (uint8_t*)c->dec_ws.p;
    w.src_cur = (int*)ws;
    w.src_prev = (int*)(ws + map_bytes);
    w.stamp = (int*)(ws + 2 * map_bytes);
    w.upd = ws + 3 * map_bytes;
    w.fill_src = (int16_t*)(ws + 3 * map_bytes + (((size_t)n * g.nb + 15) & ~(size_t)15));
    w.n = n;
    CK(cudaMemsetAsync(w.upd, 0, (size_t)n * g.nb, st));
    const size_t smem = sizeof(FixedSmem) + 17 * 17 * 4 + (size_t)g.nb + 16;
    CK(cudaFuncSetAttribute(k_dec_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    StageTimer tm(st);
    k_dec_chain<<<n_chains, 32, smem, st>>>(w);
    tm.mark("chain");
    k_dec_sources<<<(g.nb + 127) / 128, 128, 0, st>>>(w);
    const long strips = (long)n * g.nby * ((g.nbx + 7) >> 3);
    k_dec_fill<<<(unsigned)((strips + 7) / 8), 256, 0, st>>>(w);
    c->launches += 3;
    tm.mark("sources+fill");
    tm.report("decode_batch");
    CK(cudaMemcpyAsync(c->dec_prev.p, d_out + (size_t)(n - 1) * g.frame_bytes, g.frame_bytes, cudaMemcpyDeviceToDevice, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    return 1;
}

extern "C" {

int scpr_decompress_clip_dev(scpr_codec* c, const uint8_t* stream, const uint32_t* sizes, const uint8_t* ftypes, int n,
                             uint8_t* d_frames, int pitch) {
    if (!c || !stream || !sizes || !ftypes || !d_frames || n < 0) return SCPR_E_PARAM;
    return decode_batch(c, stream, sizes, ftypes, n, d_frames, pitch);
}

int scpr_decompress_clip(scpr_codec* c, const uint8_t* stream, const uint32_t* sizes, const uint8_t* ftypes, int n, uint8_t* frames,
                         int pitch) {
    if (!c || !stream || !sizes || !ftypes || !frames || n < 0) return SCPR_E_PARAM;
    if (n == 0) return 1;
    CK(cudaSetDevice(c->device));
    const size_t bytes = (size_t)n * pitch * c->g.Y;
    TRY(c->dec_frames.ensure(bytes));
    // row padding of the caller's buffer is not produced by the kernels: start from zeros
    if (pitch != c->g.X * c->g.bpp) CK(cudaMemsetAsync(c->dec_frames.p, 0, bytes, c->st));
    const int r = decode_batch(c, stream, sizes, ftypes, n, (uint8_t*)c->dec_frames.p, pitch);
    if (r != 1) return r;
    CK(cudaMemcpyAsync(frames, c->dec_frames.p, bytes, cudaMemcpyDeviceToHost, c->st));
    CK(cudaStreamSynchronize(c->st));
    return 1;
}

int scpr_decompress_frame(scpr_codec* c, const uint8_t* src, int src_len, uint8_t* dst, int pitch, int ftype) {
    if (!c || !src || !dst || src_len <= 0) return SCPR_E_PARAM;
    const uint32_t size = (uint32_t)src_len;
    const uint8_t ft = ftype ? 1 : 0;
    return scpr_decompress_clip(c, src, &size, &ft, 1, dst, pitch);
}

}  // extern "C"
