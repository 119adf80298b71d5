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
#include <stddef.h>
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
    uint32_t stream_bytes;  // bytes of the device copy incl. its zero padding, a multiple of 4
    const DecFrame* frames;
    const DecChain* chains;
    uint8_t* states;
    int f0;
    uint8_t* out;       // n output frames
    const uint8_t* prev0;   // frame decoded last by the previous call (output format, pitch g.pitch)
    uint32_t* gmap;     // per chain: nb block-source words when they do not fit in shared memory (see BlockMap)
    uint8_t* upd;       // n * nb flags: block written by the chain kernel in this frame
    int16_t* fill_src;  // n * nb: source frame of every block (k_dec_sources)
    int n;
    int msr_x, msr_y;   // v2 streams: motion range the vectors are offset by (screencap.cpp:77)
    volatile int* progress;  // per chain (mapped host memory, may be null): every frame below this index is complete
    int16_t* fill_last; // per chain x nb: source of every block after the last range k_dec_sources processed
    int f_begin, f_end, chain;  // range arguments of k_dec_sources / k_dec_fill
    uint32_t irows;     // offset of the two I-frame row buffers in shared memory, 0 = none (decode_i_rows)
};

// ---- block-source map ------------------------------------------------------------------------------
// For every 16x16 block one word: (source as of the previous frame) << 16 | (current source).  A source
// is a 16-bit code: 0xFFFF = prev0 (the frame the previous call left behind), otherwise the index of the
// frame of this call that last wrote the block, with bit 14 set when that frame is a flat frame (its
// pixels are the frame's colour; k_dec_fill paints it later).  A frame writes a block at most once, so
// "the source as of the previous frame" is the high half exactly when the low half names the frame being
// decoded.  The map lives in shared memory when the frame is small enough (SM = true), else in global
// memory; it is only ever touched by the warps of one CTA.
constexpr uint32_t SRC_PREV0 = 0xFFFFu, SRC_FLAT = 0x4000u, SRC_IDX = 0x3FFFu;
constexpr int DEC_MAX_FRAMES = 16000;  // frames per launch (14-bit frame index in the map)
template <bool SM>
struct BlockMap {
    uint32_t sbase;   // shared address (SM)
    uint32_t* g;      // global words (!SM)
    __device__ __forceinline__ uint32_t get(int b) const {
        if (SM) {
            uint32_t v;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sbase + 4u * (uint32_t)b));
            return v;
        }
        return *reinterpret_cast<volatile uint32_t*>(g + b);
    }
    __device__ __forceinline__ void set(int b, uint32_t v) const {
        if (SM)
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(sbase + 4u * (uint32_t)b), "r"(v));
        else
            *reinterpret_cast<volatile uint32_t*>(g + b) = v;
    }
    // block b is written by frame code `fc` (frame index, | SRC_FLAT for a flat frame)
    __device__ __forceinline__ void write(int b, uint32_t fc) const { set(b, (get(b) << 16) | fc); }
};
__device__ __forceinline__ uint32_t src_now(uint32_t m) { return m & 0xFFFFu; }
__device__ __forceinline__ uint32_t src_before(uint32_t m, int f) {  // as of frame f - 1
    const uint32_t cur = m & 0xFFFFu;
    return (cur != SRC_PREV0 && (cur & SRC_IDX) == (uint32_t)f) ? (m >> 16) : cur;
}

struct PixSrc {
    const uint8_t* base;
    uint32_t clr;
    bool flat;
};
__device__ __forceinline__ PixSrc resolve_src(const DecWork& w, uint32_t code) {
    PixSrc s;
    s.flat = false;
    s.clr = 0u;
    if (code == SRC_PREV0) {
        s.base = w.prev0;
    } else {
        const uint32_t idx = code & SRC_IDX;
        s.base = w.out + (size_t)idx * w.g.frame_bytes;
        if (code & SRC_FLAT) {
            s.flat = true;
            s.clr = w.frames[idx].flat_clr;
        }
    }
    return s;
}
__device__ __forceinline__ uint32_t src_px(const PixSrc& s, const Geo& g, int x, int y) { return s.flat ? s.clr : load_px(s.base, g, x, y); }
__device__ __forceinline__ void store_px(uint8_t* frame, const Geo& g, int x, int y, uint32_t v) {
    uint8_t* p = frame + (uint32_t)(y * g.pitch + x * g.bpp);
    if (g.bpp == 4)
        *reinterpret_cast<uint32_t*>(p) = v | 0xFF000000u;  // alpha := 255, screencap.cpp:1721
    else {
        p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16);
    }
}

// ---- shared memory, addressed explicitly -------------------------------------------------------------
// One resident warp per chain is bound by the latency of its dependent instruction chain, so the
// per-symbol path must be short: all table state lives in shared memory and is reached with 32-bit
// shared addresses (ld.shared / st.shared, never generic loads).
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v)); }
__device__ __forceinline__ void sts128(uint32_t a, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(x), "r"(y), "r"(z), "r"(w));
}

// Fixed tables t = 0..20 (context id CX_NTAB + t): 0-5 ntab[ptype], 6 ntab2, 7 xx, 8 bt, 9-12 sxy, 13-14 mv,
// 15-20 ptype[last].  A single warp issues roughly one instruction every 4-6 cycles on this serial code, so
// the per-symbol path is kept to a few dozen instructions and everything rare sits behind a uniform branch:
//   * ntab[0..5] (one symbol per pixel run, the hottest tables): because a table's intervals are frozen
//     between rescales (ans_contexts.h:1070-1091), a slot -> (sym<<24 | freq<<12 | cum) map lut32[4096] is
//     rebuilt together with the table (every ~128 symbols of that table, by the whole warp), so a run length
//     costs ONE shared load on the rANS dependency chain;
//   * every other table is searched by the lanes: lane j holds interval j (freq<<16 | cum), tests the slot
//     against it and forms its own candidate successor state; a ballot names the winner and one shuffle
//     delivers its state (tables of 256 / 512 symbols take a first ballot over every 8th / 16th cum).
// "Events until the next rescale" is a countdown per table (the rescale instant depends only on the number
// of events, each adds 16 to cntsum).
constexpr int FX_TOTAL = 3192;
constexpr int NCACHE = 256;                              // SmallContext cache entries
constexpr uint32_t S_FC = 0;                             // u32[FX_TOTAL]: freq << 16 | cum
constexpr uint32_t S_CNT = S_FC + FX_TOTAL * 4;          // u16[FX_TOTAL]: adaptive counters
constexpr uint32_t S_LEFT = S_CNT + FX_TOTAL * 2;        // i32[24]
constexpr uint32_t S_HEADS = S_LEFT + 24 * 4;            // u32[128]: bit s = an interval starts at slot s (lut32 rebuild)
constexpr uint32_t S_LASTH = S_HEADS + 128 * 4;          // u32[128]: last interval start before each 32-slot word
constexpr uint32_t S_LUT32 = S_LASTH + 128 * 4;          // 6 x u32[4096] (tables 0..5)
constexpr uint32_t S_KMAP = S_LUT32 + 6 * 16384;         // u8[12288]: kind of every colour context
constexpr uint32_t S_TILE = S_KMAP + NUM_COLOR_CX;       // u32[17][17]
constexpr uint32_t S_CTAG = S_TILE + 1168;               // u32[NCACHE]: context id held by a cache entry (~0 = none)
constexpr uint32_t S_CHDR = S_CTAG + NCACHE * 4;         // uint2[NCACHE]: totFr | maxpos<<16 | d<<20 | shift<<28, bonus | sfreq[maxpos]<<16
constexpr uint32_t S_CENT = S_CHDR + NCACHE * 8;         // u32[NCACHE][16]: entry k = ssym | sfreq << 8 | start << 20
#ifndef SCPR_RECON_NAP
#define SCPR_RECON_NAP 0         // ns slept per empty look of the idle reconstruction warp (0: spin)
#endif
#ifndef SCPR_COPY_BACKOFF
#define SCPR_COPY_BACKOFF 2048  // longest sleep (ns) of an idle copy warp between two looks at its ring slot
#endif
constexpr int RING = 256;                                // MV-copy commands in flight (see "helper warps" below)
constexpr uint32_t S_RING = S_CENT + NCACHE * 64;        // uint4[RING]
constexpr uint32_t S_SYNC = S_RING + RING * 16;          // u32: head, next, done, quit
constexpr int RQ = 256;                                  // pixel-block commands in flight (see "reconstruction warp" below)
constexpr uint32_t S_RQ = S_SYNC + 16;                   // uint4[RQ]
constexpr uint32_t S_RSYNC = S_RQ + RQ * 16;             // u32: -, done, {commands up to the last run, its last pixel}
constexpr uint32_t S_BTS = S_RSYNC + 16;                 // u8[nb], padded to 16; then (SM maps) u32[nb]
static_assert(S_HEADS % 16 == 0 && S_LUT32 % 16 == 0 && S_TILE % 16 == 0 && S_CTAG % 16 == 0 && S_RING % 16 == 0, "shared layout alignment");
constexpr int DEC_WARPS = 8;                             // warp 0 = the chain, warp 1 = reconstruction, warp 4 (warp 0's scheduler) idles, the rest copy
__host__ __device__ constexpr uint32_t s_map_off(int nb) { return S_BTS + (((uint32_t)nb + 15u) & ~15u); }
__host__ __device__ constexpr int fx_off(int t) {
    return t < 8 ? t * 256 : t == 8 ? 2048 : t < 13 ? 2056 + (t - 9) * 16 : t < 15 ? 2120 + (t - 13) * 512 : 3144 + (t - 15) * 8;
}

#ifdef SCPR_PROF
#define PROF_T0 const long long t0__ = clock64();
#define PROF_ADD(field) e.field += clock64() - t0__;
#define PROF_CNT(field) e.field++;
#else
#define PROF_T0
#define PROF_ADD(field)
#define PROF_CNT(field)
#endif

// The rANS state, the byte window and the colour-context registers are held identically by all 32
// lanes (every lane executes the same arithmetic on the same values), so no broadcast is needed
// between symbols.  The stream is read through a two-word window (aligned 32-bit loads, one word
// ahead), so no load sits on the renormalisation path.
struct Ent {
#ifdef SCPR_PROF
    long long c_fixed = 0, n_fixed = 0, c_color = 0, n_color = 0, c_tile = 0, n_blocks = 0, c_blkwr = 0, c_mv = 0, c_runs = 0, c_ifill = 0,
              c_hdr = 0, c_total = 0, n_gen = 0, n_resc = 0, c_rebuild = 0, n_rebuild = 0, c_small = 0, n_small = 0, c_flat = 0, n_flat = 0,
              c_raw = 0, n_raw = 0, n_miss = 0, c_drain = 0, c_lpwait = 0, n_lpwait = 0, n_lppoll = 0, n_lplong = 0;
#endif
#ifdef SCPR_LPSTAT
    uint32_t s_waits = 0, s_polls = 0, s_long = 0;  // light counters (no clock reads on the per-symbol path)
#endif
    uint32_t x;
    uint32_t w0, w1, k8;      // window: stream bytes from bit k8 of w0 on
    const uint32_t* wp;       // address of w1
    const uint32_t* wend;     // end of the stream copy (bytes of all frames of the call + padding)
    int nleft;                // symbols until the next RansDecInit
    uint32_t lastpx;          // the pixel before the next run: the colour contexts are functions of it (screencap.cpp:371-372, 616-624)
    uint32_t sb;              // shared-memory base address
    uint32_t rposted;         // pixel-block commands posted to the reconstruction warp so far
    uint32_t lp_wait;         // 0: lastpx is current; else lastpx is what the reconstruction warp leaves after this many commands
    uint32_t head;            // motion-vector copies posted so far
    // v2 streams: range coder state (RangeCoderSub, sub.h:21-45) and the count tables
    uint32_t rc_code, rc_range;
    const uint8_t* rc_p;
    uint32_t* v2;
    int msr_x, msr_y;
    ModelState* m;
    int f0;
    int lane;
};
__device__ __forceinline__ void rd_seek(Ent& e, const uint8_t* p) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    e.wp = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3) + 1;
    e.k8 = (uint32_t)(a & 3) * 8;
    e.w0 = __ldg(e.wp - 1);
    e.w1 = __ldg(e.wp);
}
__device__ __forceinline__ uint32_t rd_peek(const Ent& e) { return __funnelshift_r(e.w0, e.w1, e.k8); }
__device__ __forceinline__ void rd_skip(Ent& e, uint32_t bits) {  // bits = 8 * bytes, <= 32
    e.k8 += bits;
    if (e.k8 >= 32) {
        e.k8 -= 32;
        e.w0 = e.w1;
        ++e.wp;
        e.w1 = e.wp < e.wend ? __ldg(e.wp) : 0u;  // a corrupt stream may ask for more bytes than the call was given: zeros
    }
}
__device__ __forceinline__ void rdec_init(Ent& e) {  // RansDecInit
    e.x = rd_peek(e);
    rd_skip(e, 32);
}
// CNT = false: the caller has checked that the block cannot end within the symbols it is about to decode
// and subtracts them from nleft itself (one subtraction per pixel run instead of a test per symbol)
template <bool CNT = true>
__device__ __forceinline__ void rdec_count(Ent& e) {  // re-init every 131072 symbols (screencap.h:327-331)
    if (CNT && --e.nleft == 0) {
        rdec_init(e);
        e.nleft = RANS_BLOCK;
    }
}
// renormalisation of RansDecAdvance: a valid stream needs at most two bytes (x >= 2^11 after the update)
__device__ __forceinline__ void rdec_renorm(Ent& e, uint32_t x) {
    if (x < RANS_L) {
        const uint32_t t = rd_peek(e);
        const bool p2 = x < (1u << 15);
        x = p2 ? ((x << 16) | __byte_perm(t, 0, 0x4401)) : ((x << 8) | (t & 0xFFu));
        rd_skip(e, p2 ? 16u : 8u);
    }
    e.x = x;
}
// RansDecAdvance with the symbol's slot offset d = (x & 4095) - start already formed
__device__ __forceinline__ void rdec_advance(Ent& e, uint32_t d, uint32_t freq) { rdec_renorm(e, freq * (e.x >> PROB_BITS) + d); }

__device__ __forceinline__ void red_or_shared(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v)); }

// (Re)build the derived data of fixed table t from its counters: `rescale` applies the
// FixedSizeRansCtx rescale (freq := cnt, cum := prefix, cnt -= freq>>1, ans_contexts.h:1075-1090);
// without it the intervals in fc[] are kept (table just loaded).  Then the countdown and, for the run
// length tables, the slot map.  Whole warp.
__device__ __noinline__ void fixed_rebuild(uint32_t sb, int t, int lane, bool rescale) {
    const int off = fx_off(t), nsym = fixed_nsym(CX_NTAB + t);
    const uint32_t fcb = sb + S_FC + off * 4, cnb = sb + S_CNT + off * 2;
    const int per = (nsym + 31) >> 5, b = lane * per;
    uint32_t ns = 0;
    if (rescale) {
        uint32_t sum = 0;
        for (int j = 0; j < per; j++)
            if (b + j < nsym) sum += lds16(cnb + (b + j) * 2);
        uint32_t inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += u;
        }
        uint32_t cf = inc - sum;
        for (int j = 0; j < per; j++)
            if (b + j < nsym) {
                const uint32_t fr = lds16(cnb + (b + j) * 2);
                sts32(fcb + (b + j) * 4, (fr << 16) | cf);
                cf += fr;
                const uint32_t nc = fr - (fr >> 1);
                sts16(cnb + (b + j) * 2, nc);
                ns += nc;
            }
    } else {
        for (int j = 0; j < per; j++)
            if (b + j < nsym) ns += lds16(cnb + (b + j) * 2);
    }
    const int cntsum = (int)__reduce_add_sync(0xFFFFFFFFu, ns);
    if (lane == 0) sts32(sb + S_LEFT + t * 4, (uint32_t)((PROB_SCALE - 16 - cntsum) / 16 + 1));  // events until cntsum + 16 > 4096
    __syncwarp();
    if (t >= 6) return;
    // ---- slot map of a 256-symbol table.  (1) every symbol drops its entry at its first slot and marks that
    // slot in a 4096-bit mask; (2) per 32-slot word, the last marked slot before it (prefix maximum); (3) each
    // lane fills groups of 4 slots, consecutive lanes consecutive groups (conflict-free 128-bit accesses): the
    // entry in force at the group's first slot comes from the mask, the other three follow the group's own marks.
    const uint32_t lut = sb + S_LUT32 + (uint32_t)t * 16384u, heads = sb + S_HEADS, lasth = sb + S_LASTH;
    sts128(heads + 16 * lane, 0, 0, 0, 0);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int j = lane + 32 * i;
        const uint32_t fc = lds32(fcb + j * 4);
        const uint32_t c = fc & 0xFFFu;
        sts32(lut + c * 4, ((uint32_t)j << 24) | ((fc >> 16) << 12) | c);
        red_or_shared(heads + ((c >> 5) << 2), 1u << (c & 31));
    }
    __syncwarp();
    {
        const uint4 m = lds128(heads + 16 * lane);
        const int w0 = 4 * lane;
        // last mark inside each of my four words (-1: none), then the running maximum before each word
        const int h0 = m.x ? 32 * w0 + 31 - __clz(m.x) : -1;
        const int h1 = m.y ? 32 * (w0 + 1) + 31 - __clz(m.y) : -1;
        const int h2 = m.z ? 32 * (w0 + 2) + 31 - __clz(m.z) : -1;
        const int h3 = m.w ? 32 * (w0 + 3) + 31 - __clz(m.w) : -1;
        const int mine = max(max(h0, h1), max(h2, h3));
        int inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int u = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc = max(inc, u);
        }
        int before = __shfl_up_sync(0xFFFFFFFFu, inc, 1);
        if (lane == 0) before = 0;
        const int b1 = max(before, h0), b2 = max(b1, h1), b3 = max(b2, h2);
        sts128(lasth + 16 * lane, (uint32_t)before, (uint32_t)b1, (uint32_t)b2, (uint32_t)b3);
    }
    __syncwarp();
#pragma unroll 4
    for (int i = 0; i < 32; i++) {
        const int g = 32 * i + lane;
        const int w = g >> 3, pos = (g & 7) * 4;
        const uint32_t m = lds32(heads + 4 * w);
        const uint32_t lb = lds32(lasth + 4 * w);
        const uint4 q = lds128(lut + 16 * g);
        const uint32_t below = m & ((2u << pos) - 1u);
        const uint32_t head = below ? (uint32_t)(32 * w + 31 - __clz(below)) : lb;
        const uint32_t cur = lds32(lut + 4 * head);
        const uint32_t e1 = (m >> (pos + 1)) & 1u ? q.y : cur;
        const uint32_t e2 = (m >> (pos + 2)) & 1u ? q.z : e1;
        const uint32_t e3 = (m >> (pos + 3)) & 1u ? q.w : e2;
        sts128(lut + 16 * g, cur, e1, e2, e3);
    }
    __syncwarp();
}

// common tail of a fixed-table symbol: count it (every lane of the converged warp performs the same
// read-modify-write with the same values: no lane predicate, no divergence, no barrier on the per-symbol
// path), run the countdown (`left` was loaded from `la` before the search), rebuild the table when it expires
template <bool CNT>
__device__ __forceinline__ void fx_update(Ent& e, int t, uint32_t la, uint32_t left, int idx) {
    const uint32_t ca = e.sb + S_CNT + (uint32_t)idx * 2u;
    sts16(ca, lds16(ca) + 16);
    sts32(la, left - 1);
    rdec_count<CNT>(e);
    if (left == 1) {
        __syncwarp();
#ifdef SCPR_PROF
        const long long tr__ = clock64();
#endif
        fixed_rebuild(e.sb, t, e.lane, true);
#ifdef SCPR_PROF
        e.c_rebuild += clock64() - tr__;
        e.n_rebuild++;
#endif
    }
}

// run length through ntab[ptype] (decodeN, screencap.h:346-359, 361)
template <bool CNT = true>
__device__ __forceinline__ int dec_n(Ent& e, int ptype) {
    PROF_T0
    const uint32_t la = e.sb + S_LEFT + (uint32_t)ptype * 4u;
    const uint32_t left = lds32(la);
    const uint32_t v = e.x & (PROB_SCALE - 1);
    const uint32_t en = lds32(e.sb + S_LUT32 + ((uint32_t)ptype << 14) + (v << 2));
    const int sym = (int)(en >> 24);
    rdec_advance(e, v - (en & 0xFFFu), (en >> 12) & 0xFFFu);
    fx_update<CNT>(e, ptype, la, left, (ptype << 8) + sym);
    PROF_ADD(c_fixed) PROF_CNT(n_fixed)
    return sym;
}
// decodeF for every other table: NSYM symbols, table t at fc[off]
template <int NSYM, bool CNT = true>
__device__ __forceinline__ int dec_tab(Ent& e, int t, int off) {
    PROF_T0
    const uint32_t la = e.sb + S_LEFT + (uint32_t)t * 4u;
    const uint32_t left = lds32(la);
    const uint32_t fcb = e.sb + S_FC + (uint32_t)off * 4u;
    const uint32_t v = e.x & (PROB_SCALE - 1);
    constexpr int W = NSYM > 32 ? NSYM / 32 : NSYM;      // intervals searched by the final ballot
    constexpr int WP = W <= 8 ? 8 : 16;                  // lanes that hold one (power of two >= W)
    uint32_t base = 0;
    if (NSYM > 32) {
        const uint32_t c0 = lds32(fcb + (uint32_t)e.lane * (W * 4)) & 0xFFFFu;
        base = (uint32_t)(31 - __clz(__ballot_sync(0xFFFFFFFFu, c0 <= v) | 1u)) * W;
    }
    const uint32_t fc = lds32(fcb + (base + (uint32_t)(e.lane & (WP - 1))) * 4u);
    const uint32_t d = v - (fc & 0xFFFFu), f = fc >> 16;
    const uint32_t bh = __ballot_sync(0xFFFFFFFFu, e.lane < W && d < f);
    const int j = 31 - __clz(bh | 1u);
    const uint32_t xk = f * (e.x >> PROB_BITS) + d;
    rdec_renorm(e, __shfl_sync(0xFFFFFFFFu, xk, j));
    const int sym = (int)base + j;
    fx_update<CNT>(e, t, la, left, off + sym);
    PROF_ADD(c_fixed) PROF_CNT(n_fixed)
    return sym;
}
template <int NSYM, int T>
__device__ __forceinline__ int dec_fxc(Ent& e) { return dec_tab<NSYM>(e, T, fx_off(T)); }
template <bool CNT = true>
__device__ __forceinline__ int dec_ptype(Ent& e, int last) { return dec_tab<6, CNT>(e, 15 + last, 3144 + 8 * last); }  // decodeP
__device__ __forceinline__ int dec_bool(Ent& e) {  // decodeBool
    const uint32_t v = e.x & (PROB_SCALE - 1);
    const int flag = v >= PROB_SCALE / 2;
    rdec_advance(e, v & (PROB_SCALE / 2 - 1), PROB_SCALE / 2);
    rdec_count(e);
    return flag;
}

// ---- colour contexts -----------------------------------------------------------------------------------
// Screen content keeps a small working set of SmallContexts (kinds 4/5: <= 16 sorted symbols) busy.  They are
// held in a direct-mapped shared-memory cache, one entry per lane: entry k = ssym | sfreq << 8 | start << 20
// with start = the frequencies before k + the symbols not met below ssym[k], plus a two-word header with
// totFr, maxpos, d and -- recomputed when the context is updated, i.e. off the rANS dependency chain -- the
// normalising shift and the bonus (the code space left over by rounding, lent to the most probable symbol for
// one look-up, ans_contexts.h:195-236).  The search is then: every lane forms its interval, tests the slot
// and advances its own candidate state; a ballot names the winner, one shuffle delivers its state and one its
// entry.  The canonical ColorState in global memory is only read when a context enters the cache and written
// when it leaves (eviction, a symbol not met before, end of the launch).
// Kinds 6/7 (flat 256-symbol tables) stay in global memory, searched and rescaled by all 32 lanes; first
// occurrences of a symbol and promotions between kinds go through the serial state machine of models.cuh.
constexpr int CS_CNT = 128, CS_FREQ = 640, CS_CUM = 1152;
static_assert(offsetof(ColorState, cnt) == CS_CNT && offsetof(ColorState, freq) == CS_FREQ && offsetof(ColorState, cum) == CS_CUM &&
                  offsetof(ColorState, ssym) == 16 && offsetof(ColorState, sfreq) == 32,
              "ColorState layout");
__device__ __forceinline__ uint32_t cache_slot(int id) { return (uint32_t)(id ^ (id >> 7)) & (NCACHE - 1); }
__device__ __forceinline__ uint2 lds64(uint32_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) { asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y)); }

// inclusive prefix sum over 16 lanes
__device__ __forceinline__ uint32_t scan16(uint32_t v, int k) {
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, v, o, 16);
        if (k >= o) v += u;
    }
    return v;
}
// header of a cached SmallContext: totFr is normalised by `shift` to (2048, 4096]; what is left of 4096 is the bonus
__device__ __forceinline__ void small_header(uint32_t a, int totFr, int maxpos, int d, uint32_t fmax) {
    const int shift = max(0, __clz(totFr - 1) - 20);  // smallest shift with totFr << shift > 2048
    const uint32_t bonus = (uint32_t)(PROB_SCALE - (totFr << shift)) >> shift;
    sts64(a, (uint32_t)totFr | ((uint32_t)maxpos << 16) | ((uint32_t)d << 20) | ((uint32_t)shift << 28), bonus | (fmax << 16));
}
// cache entry h -> its ColorState (kinds 4/5 keep kind and d; the hot path changes sfreq, maxpos, totFr only)
__device__ __noinline__ void cache_writeback(ModelState* m, uint32_t sb, int lane, uint32_t h) {
    const uint32_t tag = lds32(sb + S_CTAG + 4 * h);
    if (tag != 0xFFFFFFFFu) {
        ColorState& x = m->color[tag];
        if (lane < 16) {
            const uint32_t ent = lds32(sb + S_CENT + 64 * h + 4 * lane);
            x.ssym[lane] = (uint8_t)ent;
            x.sfreq[lane] = (uint16_t)((ent >> 8) & 0xFFFu);
        }
        if (lane == 0) {
            const uint32_t h0 = lds32(sb + S_CHDR + 8 * h);
            x.cntsum = (int)(h0 & 0xFFFFu);
            x.maxpos = (uint8_t)((h0 >> 16) & 15u);
            sts32(sb + S_CTAG + 4 * h, 0xFFFFFFFFu);
        }
    }
    __syncwarp();
}
// bring SmallContext `id` into entry h (evicting its occupant)
__device__ __noinline__ void cache_load(ModelState* m, uint32_t sb, int lane, uint32_t h, int id) {
    cache_writeback(m, sb, lane, h);
    const ColorState& x = m->color[id];
    const int k = lane & 15, d = x.d, maxpos = x.maxpos, kind = x.kind;
    const uint32_t s = x.ssym[k], f = k < d ? x.sfreq[k] : 0u;
    const uint32_t inc = scan16(f, k);
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, inc, 15, 16);
    const uint32_t fmax = __shfl_sync(0xFFFFFFFFu, f, maxpos & 15);
    // the reference recomputes totFr of a Cx4 on every call (:303) and caches it for a Cx5
    const int totFr = kind == 4 ? (int)(256 - d + total) : x.cntsum;
    if (lane < 16) sts32(sb + S_CENT + 64 * h + 4 * k, k < d ? (s | (f << 8) | ((inc - f + s - k) << 20)) : 0u);
    if (lane == 0) {
        small_header(sb + S_CHDR + 8 * h, totFr, maxpos, d, fmax);
        sts32(sb + S_CTAG + 4 * h, (uint32_t)id);
    }
    __syncwarp();
}
// after the serial state machine touched a context: publish its kind
__device__ __forceinline__ void color_refresh(ModelState* m, uint32_t sb, int lane, int id) {
    __syncwarp();
    if (lane == 0) sts8(sb + S_KMAP + id, m->color[id].kind);
    __syncwarp();
}

__device__ void color_new_small(ModelState* m, uint32_t sb, int lane, int id, int c);  // defined with the other promotions below

// SmallContext decode (ans_contexts.h:238-283) from cache entry h; `ent` = this lane's entry, `hd` = the header
__device__ __forceinline__ int dec_color_small(Ent& e, int id, uint32_t h, uint32_t ent, uint2 hd) {
    const int k = e.lane & 15;
    const int maxpos = (hd.x >> 16) & 15u, shift = hd.x >> 28;
    const uint32_t bonus = hd.y & 0xFFFFu;
    const uint32_t v0 = e.x & (PROB_SCALE - 1);
    const uint32_t fr = (((ent >> 8) & 0xFFFu) + (k == maxpos ? bonus : 0u)) << shift;
    const uint32_t st = ((ent >> 20) + (k > maxpos ? bonus : 0u)) << shift;
    const uint32_t dd = v0 - st;
    const uint32_t bh = __ballot_sync(0xFFFFFFFFu, e.lane < 16 && dd < fr);
    int c;
    if (bh) {
        const int pos = 31 - __clz(bh);
        // every lane advanced its own candidate; take the winner's
        const uint32_t xk = fr * (e.x >> PROB_BITS) + dd;
        rdec_renorm(e, __shfl_sync(0xFFFFFFFFu, xk, pos));
        const uint32_t we = __shfl_sync(0xFFFFFFFFu, ent, pos);
        c = (int)(we & 255u);
        // count the symbol (ans_contexts.h:205-214): freq += 50, totFr += 50, maxpos moves on strict >, rescale
        const int d = (hd.x >> 20) & 31u;
        int ntot = (int)(hd.x & 0xFFFFu) + 50;
        const uint32_t nf = ((we >> 8) & 0xFFFu) + 50u;
        uint32_t fmax = hd.y >> 16;
        const bool moved = pos != maxpos && nf > fmax;
        const int nmax = moved ? pos : maxpos;
        fmax = (moved || pos == maxpos) ? nf : fmax;
        uint32_t ne = ent + (k == pos ? (50u << 8) : 0u) + (k > pos ? (50u << 20) : 0u);
        if (ntot + 50 > PROB_SCALE) {  // rescale: f -= f >> 1 (:186-193); the prefix sums follow
            const uint32_t sym = ent & 255u, f2 = k < d ? ((ne >> 8) & 0xFFFu) : 0u;
            const uint32_t hf = f2 - (f2 >> 1);
            const uint32_t inc = scan16(hf, k);
            ntot = (int)(256 - d + __shfl_sync(0xFFFFFFFFu, inc, 15, 16));
            ne = sym | (hf << 8) | ((inc - hf + sym - k) << 20);
            fmax -= fmax >> 1;
            PROF_CNT(n_resc)
        }
        if (e.lane < d) sts32(e.sb + S_CENT + 64 * h + 4 * k, ne);
        small_header(e.sb + S_CHDR + 8 * h, ntot, nmax, d, fmax);
    } else {  // a symbol not met yet: width 1 in the gap after the last entry below it
        const int d = (hd.x >> 20) & 31u;
        const int v = (int)(v0 >> shift);
        const int su = (int)(st >> shift), fu = (int)(fr >> shift);
        const uint32_t bg = __ballot_sync(0xFFFFFFFFu, e.lane < d && v >= su);
        int lastSymb = 0, cumFr = 0;
        if (bg) {
            const int K = 31 - __clz(bg);
            lastSymb = (int)(__shfl_sync(0xFFFFFFFFu, ent, K) & 255u) + 1;
            cumFr = __shfl_sync(0xFFFFFFFFu, su + fu, K);
        }
        c = (lastSymb + v - cumFr) & 255;
        rdec_advance(e, v0 - (uint32_t)(v << shift), (uint32_t)(1 << shift));
        cache_writeback(e.m, e.sb, e.lane, h);  // insert / promote on the canonical state: the general path
        color_new_small(e.m, e.sb, e.lane, id, c);
        PROF_CNT(n_gen)
    }
    return c;
}

// kinds 6/7: flat tables cnt/freq/cum[256]; lane owns symbols [8*lane, 8*lane + 8)
__device__ __forceinline__ uint32_t half_at(const uint4& r, int idx) {  // halfword idx (0..7) of a 128-bit row
    const int w = idx >> 1;
    const uint32_t word = w == 0 ? r.x : w == 1 ? r.y : w == 2 ? r.z : r.w;
    return (idx & 1) ? (word >> 16) : (word & 0xFFFFu);
}
__device__ __noinline__ void flat_rescale(uint8_t* xs, int kind, int lane) {  // c6 rescale (:742-796) / Cx7 rescale (:959-981)
    __syncwarp();
    const uint4 cn = *reinterpret_cast<const uint4*>(xs + CS_CNT + lane * 16);
    const int fshift = xs[1], d = *reinterpret_cast<const uint16_t*>(xs + 4);
    const uint32_t c0 = kind == 6 ? (1u << (fshift > 0 ? fshift - 1 : 0)) : 0u;
    uint32_t fr[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint32_t w = j < 2 ? cn.x : j < 4 ? cn.y : j < 6 ? cn.z : cn.w;
        const uint32_t c = (j & 1) ? (w >> 16) : (w & 0xFFFFu);
        fr[j] = (kind == 6 && c == 0) ? c0 : c;
        sum += fr[j];
    }
    uint32_t inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += u;
    }
    uint32_t cf = inc - sum, ns = 0;
    uint32_t cu[8], nc[8];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        cu[j] = cf & 0xFFFFu;
        cf += fr[j];
        const uint32_t w = j < 2 ? cn.x : j < 4 ? cn.y : j < 6 ? cn.z : cn.w;
        const uint32_t c = (j & 1) ? (w >> 16) : (w & 0xFFFFu);
        nc[j] = c - (c >> 1);  // kind 6: symbols not met keep 0
        ns += nc[j];
    }
    *reinterpret_cast<uint4*>(xs + CS_FREQ + lane * 16) =
        make_uint4(fr[0] | (fr[1] << 16), fr[2] | (fr[3] << 16), fr[4] | (fr[5] << 16), fr[6] | (fr[7] << 16));
    *reinterpret_cast<uint4*>(xs + CS_CUM + lane * 16) =
        make_uint4(cu[0] | (cu[1] << 16), cu[2] | (cu[3] << 16), cu[4] | (cu[5] << 16), cu[6] | (cu[7] << 16));
    *reinterpret_cast<uint4*>(xs + CS_CNT + lane * 16) =
        make_uint4(nc[0] | (nc[1] << 16), nc[2] | (nc[3] << 16), nc[4] | (nc[5] << 16), nc[6] | (nc[7] << 16));
    ns = __reduce_add_sync(0xFFFFFFFFu, ns);
    if (lane == 0) {
        if (kind == 6) {
            const int nfs = fshift > 0 ? fshift - 1 : 0;
            const int shft = nfs > 0 ? nfs - 1 : 0;
            xs[1] = (uint8_t)nfs;
            *reinterpret_cast<int*>(xs + 8) = (int)((((256 - d) << shft) + ns) & 0xFFFFu);
        } else {
            *reinterpret_cast<int*>(xs + 8) = (int)ns;
        }
    }
    __syncwarp();
}

// ---- promotions and first occurrences, by the whole warp --------------------------------------------------------------
// Photo / noise content walks every colour context through the reference's chain of representations (Cx1 -> Cx2 -> Cx3
// sets, SmallContext, Cx6, Cx7; ans_contexts.cpp:3-50) and meets a new symbol in a context hundreds of thousands of
// times per frame.  The transitions that touch all 256 symbols are done here with 8 symbols per lane and warp scans; only
// the edits of the <= 16-entry lists stay on one lane (cc_encode_counted / cc_update_raw of models.cuh, which remain the
// single-lane statement of the same rules and are what the encoder's replay uses).
__device__ __forceinline__ uint32_t warp_excl_scan(uint32_t v, int lane, uint32_t& total) {
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += u;
    }
    total = __shfl_sync(0xFFFFFFFFu, inc, 31);
    return inc - v;
}
__device__ __forceinline__ void store_row8(uint8_t* row, const uint32_t* v) {  // eight u16 as one 128-bit store
    *reinterpret_cast<uint4*>(row) = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
}
// Cx6 tables from a set of met symbols with start frequencies (c6_build + c6_calcsum of models.cuh; ans_contexts.h:454-533,
// 549-555).  met = this lane's 8-bit membership mask, frs[j] = start frequency of symbol 8*lane + j (met symbols only);
// extra_c >= 0: that symbol additionally gets the count of a first occurrence and joins the set (Cx6::create(Cx5&, c)).
__device__ __noinline__ void c6_build_w(uint8_t* xs, uint32_t met, const uint32_t* frs, int d, int totFr, int extra_c, int lane) {
    int shift = 0, tot = totFr;
    while (tot <= PROB_SCALE / 2) {
        tot <<= 1;
        shift++;
    }
    uint32_t fr[8], cn[8], cu[8], sum = 0, csum = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const bool m = (met >> j) & 1u;
        fr[j] = m ? (frs[j] << shift) : (1u << shift);
        cn[j] = m ? (fr[j] - (fr[j] >> 1)) & 0xFFFFu : 0u;
        fr[j] &= 0xFFFFu;
        sum += fr[j];
    }
    if (extra_c >= 0 && (extra_c >> 3) == lane) {
        const uint32_t f1 = 1u << shift;
        cn[extra_c & 7] = (f1 - (f1 >> 1) + (25u << shift)) & 0xFFFFu;
        d++;
    }
    d = __shfl_sync(0xFFFFFFFFu, d, extra_c >= 0 ? (extra_c >> 3) : 0);
    uint32_t total;
    uint32_t cf = warp_excl_scan(sum, lane, total);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        cu[j] = cf & 0xFFFFu;
        cf += fr[j];
        csum += cn[j];
    }
    store_row8(xs + CS_FREQ + lane * 16, fr);
    store_row8(xs + CS_CUM + lane * 16, cu);
    store_row8(xs + CS_CNT + lane * 16, cn);
    csum = __reduce_add_sync(0xFFFFFFFFu, csum);
    if (lane == 0) {
        const int shft = shift > 0 ? shift - 1 : 0;
        xs[0] = 6;                                             // kind
        xs[1] = (uint8_t)shift;                                // fshift
        *reinterpret_cast<uint16_t*>(xs + 4) = (uint16_t)d;    // d
        *reinterpret_cast<int*>(xs + 8) = (int)((((256 - d) << shft) + csum) & 0xFFFFu);  // cntsum = c6_calcsum
    }
    __syncwarp();
}
__device__ __forceinline__ uint32_t seen_byte(const ColorState& x, int lane) { return (x.seen[lane >> 2] >> ((lane & 3) * 8)) & 0xFFu; }
// Cx2 -> Cx6 (create23, ans_contexts.h:491-533): every met symbol starts with f0, the repeated one with 2 * f0
__device__ __forceinline__ void c6_from_set_w(ColorState& x, int c, int f0, int lane) {
    const uint32_t met = seen_byte(x, lane);
    const int d = x.d;
    uint32_t frs[8];
#pragma unroll
    for (int j = 0; j < 8; j++) frs[j] = (lane * 8 + j == c) ? 2u * (uint32_t)f0 : (uint32_t)f0;
    __syncwarp();
    c6_build_w(reinterpret_cast<uint8_t*>(&x), met, frs, d, 256 - d + d * f0 + f0, -1, lane);
}
// Cx5 -> Cx6 (Cx6::create(Cx5&, c), ans_contexts.h:454-489): the sixteen listed symbols keep their frequencies, c joins
__device__ __forceinline__ void c6_from_small_w(ColorState& x, int c, int lane) {
    const int d = x.d;  // 16
    const uint32_t ks = lane < d ? x.ssym[lane & 15] : 0x100u, kf = lane < d ? x.sfreq[lane & 15] : 0u;
    uint32_t tot;
    warp_excl_scan(kf, lane, tot);
    uint32_t met = 0, frs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k < d; k++) {
        const uint32_t sk = __shfl_sync(0xFFFFFFFFu, ks, k), fk = __shfl_sync(0xFFFFFFFFu, kf, k);
        if ((int)(sk >> 3) == lane) {
            met |= 1u << (sk & 7);
#pragma unroll
            for (int j = 0; j < 8; j++)
                if ((int)(sk & 7) == j) frs[j] = fk;
        }
    }
    __syncwarp();
    c6_build_w(reinterpret_cast<uint8_t*>(&x), met, frs, d, (int)(256 - d + tot), c, lane);
}
// Cx6 -> Cx7 (c7_from_c6, ans_contexts.h:868-915): symbols not met yet get the count of a fresh slot; c itself is not counted
__device__ __forceinline__ void c7_from_c6_w(ColorState& x, int lane) {
    uint8_t* xs = reinterpret_cast<uint8_t*>(&x);
    const uint32_t funmet = 1u << x.fshift, cu = funmet - (funmet >> 1);
    const uint4 r = *reinterpret_cast<const uint4*>(xs + CS_CNT + lane * 16);
    uint32_t v[8] = {r.x & 0xFFFFu, r.x >> 16, r.y & 0xFFFFu, r.y >> 16, r.z & 0xFFFFu, r.z >> 16, r.w & 0xFFFFu, r.w >> 16};
#pragma unroll
    for (int j = 0; j < 8; j++)
        if (!v[j]) v[j] = cu & 0xFFFFu;
    __syncwarp();
    store_row8(xs + CS_CNT + lane * 16, v);
    if (lane == 0) xs[0] = 7;
    __syncwarp();
}
// Cx3 -> Cx7 (c7_from_set, ans_contexts.h:917-951)
__device__ __forceinline__ void c7_from_set_w(ColorState& x, int c, int lane) {
    uint8_t* xs = reinterpret_cast<uint8_t*>(&x);
    const uint32_t met = seen_byte(x, lane);
    const int d = x.d;
    const uint32_t f0 = (uint32_t)((PROB_SCALE - (256 - d)) / (d + 1)), c0 = f0 - (f0 >> 1);
    uint32_t fr[8], cn[8], cu[8], sum = 0, csum = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const bool m = (met >> j) & 1u;
        fr[j] = m ? f0 : 1u;
        cn[j] = m ? c0 : 1u;
        if (lane * 8 + j == c) {
            fr[j] += f0;
            cn[j] += 16u;
        }
        fr[j] &= 0xFFFFu;
        cn[j] &= 0xFFFFu;
        sum += fr[j];
        csum += cn[j];
    }
    uint32_t total;
    uint32_t cf = warp_excl_scan(sum, lane, total);
#pragma unroll
    for (int j = 0; j < 8; j++) {
        cu[j] = cf & 0xFFFFu;
        cf += fr[j];
    }
    __syncwarp();
    store_row8(xs + CS_FREQ + lane * 16, fr);
    store_row8(xs + CS_CUM + lane * 16, cu);
    store_row8(xs + CS_CNT + lane * 16, cn);
    csum = __reduce_add_sync(0xFFFFFFFFu, csum);
    if (lane == 0) {
        *reinterpret_cast<int*>(xs + 8) = (int)csum;
        xs[0] = 7;
    }
    __syncwarp();
}
// Cx1 -> SmallContext (small_from_set, ans_contexts.h:161-172): the <= 14 met symbols in order, 50 each, 100 for the repeated one
__device__ __forceinline__ void small_from_set_w(ColorState& x, int c, int lane) {
    uint32_t wbits = lane < 8 ? x.seen[lane & 7] : 0u;
    const int d = x.d;
    uint32_t total;
    uint32_t pos = warp_excl_scan((uint32_t)__popc(wbits), lane, total);
    __syncwarp();
    while (wbits) {
        const int sidx = lane * 32 + __ffs(wbits) - 1;
        wbits &= wbits - 1;
        if (pos < 16) {
            x.ssym[pos] = (uint8_t)sidx;
            x.sfreq[pos] = (uint16_t)(sidx == c ? 100 : 50);
            if (sidx == c) x.maxpos = (uint8_t)pos;
        }
        pos++;
    }
    if (lane >= d && lane < 16) x.sfreq[lane] = 0;
    __syncwarp();
}
// Context::update for a context without statistics (kinds 0..3: the byte was stored raw); cc_update_raw of models.cuh
__device__ __noinline__ void color_raw_w(ColorState& x, int c, int f0, int lane) {
    const int kind = x.kind, d = x.d;
    const bool have = kind != 0 && seen_has(x, c);
    __syncwarp();
    if (kind == 0 || !have) {  // start the set, or one more distinct symbol: a few scalar stores
        if (lane == 0) cc_update_raw(x, c, f0);
    } else if (kind == 1) {
        small_from_set_w(x, c, lane);
        if (lane == 0) {
            if (d <= 4)
                x.kind = 4;
            else {
                x.kind = 5;
                x.cntsum = small_calcsum(x);
            }
        }
    } else if (kind == 2) {
        c6_from_set_w(x, c, f0, lane);
    } else {
        c7_from_set_w(x, c, lane);
    }
    __syncwarp();
}
// a symbol its context (kinds 4..6) has not met before: insert / place it, or promote the context (cc_encode_counted)
__device__ __noinline__ void color_new_w(ColorState& x, int c, int lane) {
    uint8_t* xs = reinterpret_cast<uint8_t*>(&x);
    const int kind = x.kind, d = x.d;
    __syncwarp();
    if (kind == 5 && d == 16) {
        c6_from_small_w(x, c, lane);
    } else if (kind == 6) {
        if (d >= 40) {  // MaxD6 (ans_contexts.h:631)
            c7_from_c6_w(x, lane);
        } else {        // placeSymbol + incrCnt (ans_contexts.h:621-638, 686-691)
            const int fshift = x.fshift, cs = x.cntsum;
            const int step = 25 << fshift;
            __syncwarp();
            if (lane == 0) {
                const int fr = 1 << fshift;
                x.cnt[c] = (uint16_t)(fr - (fr >> 1) + step);
                x.d = (uint16_t)(d + 1);
                x.cntsum = (cs + step) & 0xFFFF;
            }
            if (((cs + step) & 0xFFFF) + step > PROB_SCALE) flat_rescale(xs, 6, lane);
        }
    } else {
        if (lane == 0) cc_encode_counted(x, c);
    }
    __syncwarp();
}
__device__ __noinline__ void color_new_small(ModelState* m, uint32_t sb, int lane, int id, int c) {
    color_new_w(m->color[id], c, lane);
    color_refresh(m, sb, lane, id);
}

__device__ __forceinline__ int dec_color_flat(Ent& e, uint8_t* xs, int id, int kind) {
    const uint4 hd = *reinterpret_cast<const uint4*>(xs);
    const uint4 cr = *reinterpret_cast<const uint4*>(xs + CS_CUM + e.lane * 16);
    const uint4 fq = *reinterpret_cast<const uint4*>(xs + CS_FREQ + e.lane * 16);
    const uint4 cq = *reinterpret_cast<const uint4*>(xs + CS_CNT + e.lane * 16);  // fetched with the rest: no second round trip for cnt[c]
    const uint32_t v = e.x & (PROB_SCALE - 1);
    // number of this lane's cumulative frequencies <= v, two halfwords at a time (values < 2^15: no borrow between halves)
    const uint32_t vv = (v * 0x10001u) | 0x80008000u;
    const uint32_t m = (((vv - cr.x) & 0x80008000u) >> 15) | (((vv - cr.y) & 0x80008000u) >> 14) | (((vv - cr.z) & 0x80008000u) >> 13) |
                       (((vv - cr.w) & 0x80008000u) >> 12);
    const int idx = __popc(m) - 1;
    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, m & 1u);
    const int L = 31 - __clz(bal | 1u);
    const uint32_t cum = half_at(cr, idx & 7), freq = half_at(fq, idx & 7);
    const uint32_t xk = freq * (e.x >> PROB_BITS) + (v - cum);
    rdec_renorm(e, __shfl_sync(0xFFFFFFFFu, xk, L));
    const int c = L * 8 + __shfl_sync(0xFFFFFFFFu, idx, L);
    // count the symbol
    uint16_t* cnt = reinterpret_cast<uint16_t*>(xs + CS_CNT);
    const int cn = (int)__shfl_sync(0xFFFFFFFFu, half_at(cq, idx & 7), L), cs = (int)hd.z;
    const int step = kind == 7 ? 16 : (25 << ((hd.x >> 8) & 255));
    if (cn != 0 || kind == 7) {  // met before: count (+ rescale), ans_contexts.h:686-691, 959-981
        if (e.lane == 0) {
            cnt[c] = (uint16_t)(cn + step);
            *reinterpret_cast<int*>(xs + 8) = (cs + step) & 0xFFFF;
        }
        if (cs + 2 * step > PROB_SCALE) {
            flat_rescale(xs, kind, e.lane);
            PROF_CNT(n_resc)
        }
    } else {  // first occurrence in a Cx6: placeSymbol or the promotion to Cx7
        color_new_w(e.m->color[id], c, e.lane);
        color_refresh(e.m, e.sb, e.lane, id);
        PROF_CNT(n_gen)
    }
    return c;
}
template <bool CNT>
__device__ __forceinline__ int dec_color(Ent& e, int id) {  // decodeC, screencap.h:318-333
    PROF_T0
    const uint32_t h = cache_slot(id);
    const uint32_t tag = lds32(e.sb + S_CTAG + 4 * h);
    uint32_t ent = lds32(e.sb + S_CENT + 64 * h + 4 * (e.lane & 15));
    uint2 hd = lds64(e.sb + S_CHDR + 8 * h);
    int c;
    if (tag != (uint32_t)id) {
        const int kind = (int)lds8(e.sb + S_KMAP + id);
        if (kind >= 6) {
            c = dec_color_flat(e, reinterpret_cast<uint8_t*>(&e.m->color[id]), id, kind);
            rdec_count<CNT>(e);
            PROF_ADD(c_flat) PROF_CNT(n_flat) PROF_ADD(c_color) PROF_CNT(n_color)
            return c;
        }
        if (kind < 4) {  // no statistics yet: the byte is stored raw (screencap.h:324-325)
            c = (int)(rd_peek(e) & 0xFFu);
            rd_skip(e, 8);
            color_raw_w(e.m->color[id], c, e.f0, e.lane);
            color_refresh(e.m, e.sb, e.lane, id);
            rdec_count<CNT>(e);
            PROF_ADD(c_raw) PROF_CNT(n_raw) PROF_ADD(c_color) PROF_CNT(n_color)
            return c;
        }
        cache_load(e.m, e.sb, e.lane, h, id);
        ent = lds32(e.sb + S_CENT + 64 * h + 4 * (e.lane & 15));
        hd = lds64(e.sb + S_CHDR + 8 * h);
        PROF_CNT(n_miss)
    }
    c = dec_color_small(e, id, h, ent, hd);
    rdec_count<CNT>(e);
    PROF_ADD(c_small) PROF_CNT(n_small) PROF_ADD(c_color) PROF_CNT(n_color)
    return c;
}
__device__ __forceinline__ uint32_t ldv_shared(uint32_t a) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void stv_shared(uint32_t a, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
// The pixel before a literal is the literal's context.  In P frames pixels are produced by the reconstruction warp: when
// the run before the literal was a predicted one, the chain warp waits here until that run has been filled (it has had the
// time of one symbol decode to do so) and takes the run's last pixel from it.
__device__ __forceinline__ void fetch_lastpx(Ent& e) {
    if (e.lp_wait) {
        // {commands executed up to and including the last run, that run's last pixel}: one 64-bit word, written at once
        PROF_T0
        uint2 lp;
#if defined(SCPR_PROF) || defined(SCPR_LPSTAT)
        int polls__ = 0;
#endif
        do {
            asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lp.x), "=r"(lp.y) : "r"(e.sb + S_RSYNC + 8u) : "memory");
#if defined(SCPR_PROF) || defined(SCPR_LPSTAT)
            polls__++;
#endif
        } while ((int)(lp.x - e.lp_wait) < 0);
#ifdef SCPR_LPSTAT
        e.s_waits++;
        e.s_polls += polls__;
        if (polls__ > 10) e.s_long++;
#endif
#ifdef SCPR_PROF
        e.n_lppoll += polls__;
        if (polls__ > 25) e.n_lplong++;
#endif
        e.lastpx = lp.y;
        e.lp_wait = 0;
        PROF_ADD(c_lpwait) PROF_CNT(n_lpwait)
    }
}
// The context of a colour byte is made of the two bytes coded before it, quantised to 6 bits: for byte 0 those
// are bytes 2 and 1 of the pixel before the run (MAKECX1, screencap.h:36; :371-372, 488-493, 1417-1419).
template <bool CNT = true>
__device__ __forceinline__ uint32_t dec_rgb(Ent& e) {  // DecodeRGB, screencap.cpp:662-679
    fetch_lastpx(e);
    uint32_t px = 0;
    uint32_t cx = (e.lastpx >> 18) & 63u, cx1 = (e.lastpx >> 4) & 0xFC0u;
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        const uint32_t v = (uint32_t)dec_color<CNT>(e, ch * 4096 + (int)(cx + cx1)) & 255u;
        cx1 = cx << 6;
        cx = v >> 2;
        px |= v << (8 * ch);
    }
    e.lastpx = px;
    return px;
}

// ---- v2 streams: the range coder of ScreenPressor 2.x (decode compatibility) -----------------
// Replaces RangeCoderSub::DecodeBegin / GetFreq / Decode / DecodeVal / DecodeValUni (sub.h:30-42, sub.cpp:44-61, 88-113,
// 146-178) and UseRC's table set (screencap.h:105-265).  Every model is a plain count table cnt[maxc] + total with a
// per-model increment; a symbol is found by cumulating the counts in index order.  DecodeValUni's 16 group sums only
// accelerate that same search (they are kept consistent with the counts at all times), so colour bytes go through the
// same routine and the group sums are not stored.  Tables live in the chain's state slot in global memory (12.6 MB);
// this is a compatibility path for old files, written for exactness, not speed: the 32 lanes cumulate a table together.
constexpr uint32_t RC_TOP = 1u << 24, RC_BOT = 1u << 16;                       // sub.h:15-18
constexpr int V2_N = 0, V2_N2 = 6 * 257, V2_XX = V2_N2 + 257, V2_BT = V2_XX + 257, V2_SXY = V2_BT + 6, V2_MV = V2_SXY + 4 * 17,
              V2_PT = V2_MV + 2 * 513, V2_C = 3200, V2_WORDS = V2_C + 3 * 4096 * 257;
static_assert(V2_PT + 6 * 7 <= V2_C && (size_t)V2_WORDS * 4 <= sizeof(ModelState), "v2 tables fit a model state slot");
__device__ __forceinline__ void rc_begin(Ent& e, const uint8_t* p) {  // DecodeBegin: five bytes, the first is the encoder's empty cache
    e.rc_code = 0;
    e.rc_range = 0xFFFFFFFFu;
    for (int i = 0; i < 5; i++) e.rc_code = (e.rc_code << 8) | p[i];
    e.rc_p = p + 5;
}
// DecodeVal(c, cnt, totfr = cnt[maxc], maxc, step).  CAP = 512 / 256 / 32 bounds maxc at compile time.
template <int CAP>
__device__ __noinline__ int rc_val(Ent& e, uint32_t* cnt, int maxc, uint32_t step) {
    constexpr int PER = CAP / 32;
    const int lane = e.lane;
    uint32_t v[PER], sum = 0;
    const uint32_t tot = cnt[maxc];
#pragma unroll
    for (int j = 0; j < PER; j++) {
        const int idx = lane * PER + j;
        v[j] = idx < maxc ? cnt[idx] : 0u;
        sum += v[j];
    }
    const uint32_t r = e.rc_range / tot;      // GetFreq: code / (range /= totFreq)
    const uint32_t value = e.rc_code / r;
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += u;
    }
    const uint32_t hit = __ballot_sync(0xFFFFFFFFu, incl > value);
    const int L = hit ? __ffs(hit) - 1 : (maxc - 1) / PER;  // a corrupt stream may point past the table: last symbol
    uint32_t cum = incl - sum;
    int j = 0;
#pragma unroll
    for (int k = 0; k < PER - 1; k++)
        if (j == k && value >= cum + v[k] && lane * PER + k + 1 < maxc) {
            cum += v[k];
            j = k + 1;
        }
    uint32_t f = v[0];
#pragma unroll
    for (int k = 1; k < PER; k++)
        if (j == k) f = v[k];
    const int c = __shfl_sync(0xFFFFFFFFu, lane * PER + j, L);
    cum = __shfl_sync(0xFFFFFFFFu, cum, L);
    f = __shfl_sync(0xFFFFFFFFu, f, L);
    // Decode (sub.cpp:49-61)
    uint32_t code = e.rc_code - cum * r, range = r * f;
    while (range < RC_TOP) {
        code = (code << 8) | (e.rc_p < reinterpret_cast<const uint8_t*>(e.wend) ? *e.rc_p : 0u);
        e.rc_p++;
        range <<= 8;
    }
    e.rc_code = code;
    e.rc_range = range;
    // count it; halve everything once the total passes 2^16 (sub.cpp:100-111)
    if (tot + step > RC_BOT) {
        uint32_t ns = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            const int idx = lane * PER + k;
            if (idx < maxc) {
                const uint32_t nv = ((v[k] + (idx == c ? step : 0u)) >> 1) + 1u;
                cnt[idx] = nv;
                ns += nv;
            }
        }
        ns = __reduce_add_sync(0xFFFFFFFFu, ns);
        if (lane == 0) cnt[maxc] = ns;
    } else if (lane == 0) {
        cnt[c] += step;
        cnt[maxc] = tot + step;
    }
    __syncwarp();
    return c;
}
// table t of the v4 numbering (0-5 ntab, 6 ntab2, 7 xx, 8 bt, 9-12 sxy, 13-14 mv, 15-20 ptype) in its v2 form
template <int T>
__device__ __forceinline__ int rc_fx(Ent& e) {
    if (T < 6) return rc_val<256>(e, e.v2 + V2_N + T * 257, 256, 400);           // SC_NSTEP
    if (T == 6) return rc_val<256>(e, e.v2 + V2_N2, 256, 20);                    // SC_BTNSTEP
    if (T == 7) return rc_val<256>(e, e.v2 + V2_XX, 256, 1);                     // SC_XXSTEP
    if (T == 8) return rc_val<32>(e, e.v2 + V2_BT, 5, 10);                       // SC_BTSTEP
    if (T < 13) return rc_val<32>(e, e.v2 + V2_SXY + (T - 9) * 17, 16, 100);     // SC_SXYSTEP
    if (T == 13) return rc_val<512>(e, e.v2 + V2_MV, 2 * e.msr_x, 100);          // SC_MSTEP
    return rc_val<512>(e, e.v2 + V2_MV + 513, 2 * e.msr_y, 100);
}
__device__ __forceinline__ int rc_n(Ent& e, int ptype) { return rc_val<256>(e, e.v2 + V2_N + ptype * 257, 256, 400); }
__device__ __forceinline__ uint32_t rc_rgb(Ent& e) {  // DecodeRGB with decodeC = DecodeValUni(cntab, step SC_STEP)
    fetch_lastpx(e);
    uint32_t px = 0;
    uint32_t cx = (e.lastpx >> 18) & 63u, cx1 = (e.lastpx >> 4) & 0xFC0u;
#pragma unroll 1
    for (int ch = 0; ch < 3; ch++) {
        const uint32_t v = (uint32_t)rc_val<256>(e, e.v2 + V2_C + (size_t)(ch * 4096 + (int)(cx + cx1)) * 257, 256, 400) & 255u;
        cx1 = cx << 6;
        cx = v >> 2;
        px |= v << (8 * ch);
    }
    e.lastpx = px;
    return px;
}
__device__ __forceinline__ int rc_run(Ent& e, int& ptype, uint32_t& c) {
    ptype = rc_val<32>(e, e.v2 + V2_PT + ptype * 7, 6, 1000);  // SC_UNSTEP
    if (!ptype) c = rc_rgb(e);
    return rc_n(e, ptype);
}
__device__ void rc_renew(Ent& e) {  // RenewI for UseRC: every count 1, totals = table sizes (screencap.h:147-262)
    for (int i = e.lane; i < V2_C; i += 32) e.v2[i] = 1u;
    __syncwarp();
    if (e.lane == 0) {
        for (int t = 0; t < 6; t++) e.v2[V2_N + t * 257 + 256] = 256u;
        e.v2[V2_N2 + 256] = 256u;
        e.v2[V2_XX + 256] = 256u;
        e.v2[V2_BT + 5] = 5u;
        for (int k = 0; k < 4; k++) e.v2[V2_SXY + k * 17 + 16] = 16u;
        e.v2[V2_MV + 2 * e.msr_x] = 2u * (uint32_t)e.msr_x;
        e.v2[V2_MV + 513 + 2 * e.msr_y] = 2u * (uint32_t)e.msr_y;
        for (int l = 0; l < 6; l++) e.v2[V2_PT + l * 7 + 6] = 6u;
    }
    for (int i = e.lane; i < 3 * 4096 * 257; i += 32) e.v2[V2_C + i] = (i % 257 == 256) ? 256u : 1u;
    __syncwarp();
}

// ---- the parse below is shared by both stream generations: these pick the coder ---------------------------------------
template <bool V2, int NSYM, int T>
__device__ __forceinline__ int sym_fx(Ent& e) {
    if (V2) return rc_fx<T>(e);
    return dec_fxc<NSYM, T>(e);
}

// the four sub-rect symbols of a partial block (SXY tables 9..12, 16 symbols each; screencap.cpp:1321-1331), packed
// x1 | y1 << 4 | x2-1 << 8 | y2-1 << 12 relative to the block.  The four tables are distinct, so their intervals and
// countdowns are fetched up front and the four decodes run back to back on registers.
template <bool V2>
__device__ __forceinline__ uint32_t dec_rect(Ent& e) {
    uint32_t r = 0;
    if (V2) {
        r = (uint32_t)rc_fx<9>(e);
        r |= (uint32_t)rc_fx<10>(e) << 4;
        r |= (uint32_t)rc_fx<11>(e) << 8;
        r |= (uint32_t)rc_fx<12>(e) << 12;
    } else {
    uint32_t fc[4], left[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        fc[k] = lds32(e.sb + S_FC + (uint32_t)(2056 + 16 * k + (e.lane & 15)) * 4u);
        left[k] = lds32(e.sb + S_LEFT + (uint32_t)(9 + k) * 4u);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t v = e.x & (PROB_SCALE - 1);
        const uint32_t d = v - (fc[k] & 0xFFFFu), f = fc[k] >> 16;
        const uint32_t bh = __ballot_sync(0xFFFFFFFFu, e.lane < 16 && d < f);
        const int j = 31 - __clz(bh | 1u);
        const uint32_t xk = f * (e.x >> PROB_BITS) + d;
        rdec_renorm(e, __shfl_sync(0xFFFFFFFFu, xk, j));
        fx_update<true>(e, 9 + k, e.sb + S_LEFT + (uint32_t)(9 + k) * 4u, left[k], 2056 + 16 * k + j);
        r |= (uint32_t)j << (4 * k);
    }
    }
    return r;
}

// the symbols of one pixel run: type, colour of a literal, length (screencap.cpp:478-486, 1400-1412).  At most five
// symbols: when the rANS block cannot end within them they are counted with one subtraction.
template <bool V2 = false>
__device__ __forceinline__ int dec_run(Ent& e, int& ptype, uint32_t& c) {
    if (V2) return rc_run(e, ptype, c);
    int n;
    if (e.nleft > 8) {
        // type and length, software-pipelined by hand (a lone warp has nobody else to fill its load latencies): the run
        // length's table look-up is issued as soon as the type is known, the type table is counted while it is in flight
        const int t = 15 + ptype, off = 3144 + 8 * ptype;
        const uint32_t la1 = e.sb + S_LEFT + (uint32_t)t * 4u;
        const uint32_t left1 = lds32(la1);
        const uint32_t v = e.x & (PROB_SCALE - 1);
        const uint32_t fc = lds32(e.sb + S_FC + (uint32_t)(off + (e.lane & 7)) * 4u);
        const uint32_t d = v - (fc & 0xFFFFu), f = fc >> 16;
        const uint32_t bh = __ballot_sync(0xFFFFFFFFu, e.lane < 6 && d < f);
        const int pt = 31 - __clz(bh | 1u);
        const uint32_t xk = f * (e.x >> PROB_BITS) + d;
        rdec_renorm(e, __shfl_sync(0xFFFFFFFFu, xk, pt));
        ptype = pt;
        if (!pt) {
            fx_update<false>(e, t, la1, left1, off);
            c = dec_rgb<false>(e);
            n = dec_n<false>(e, 0);
        } else {
            const uint32_t la2 = e.sb + S_LEFT + (uint32_t)pt * 4u;
            const uint32_t left2 = lds32(la2);
            const uint32_t v2 = e.x & (PROB_SCALE - 1);
            const uint32_t en = lds32(e.sb + S_LUT32 + ((uint32_t)pt << 14) + (v2 << 2));
            fx_update<false>(e, t, la1, left1, off + pt);
            n = (int)(en >> 24);
            rdec_advance(e, v2 - (en & 0xFFFu), (en >> 12) & 0xFFFu);
            fx_update<false>(e, pt, la2, left2, (pt << 8) + n);
        }
        e.nleft -= pt ? 2 : 5;
    } else {
        ptype = dec_ptype<true>(e, ptype);
        if (!ptype) c = dec_rgb<true>(e);
        n = dec_n<true>(e, ptype);
    }
    return n;
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

template <bool V2>
__device__ void decode_i(const DecWork& w, Ent& e, uint8_t* frame, int lane) {
    const Geo& g = w.g;
    const int X = g.X, Y = g.Y;
    const bool padded = ((X * 3 + 3) & ~3) != X * 3;
    IPos p{0, 0};
    int ptype = 0;
    int hdr = X + 1;
    uint32_t lastv = 0;  // the pixel before p in raster order
    // first row and one pixel: (rgb, n) pairs, lengths in ntab[0] (screencap.cpp:423-438)
    while (hdr > 0) {
        const uint32_t c = V2 ? rc_rgb(e) : dec_rgb(e);
        const int n = V2 ? rc_n(e, 0) : dec_n(e, 0);
        if (n <= 0) return;
        for (int i = lane; i < n; i += 32) {
            const IPos q = ipos_add(p, i, X);
            if (q.y < Y) store_px(frame, g, q.x, q.y, c);
        }
        p = ipos_add(p, n, X);
        hdr -= n;
        lastv = c;
    }
    __syncwarp();
    while (p.y < Y) {
        uint32_t c = lastv;  // type 1: the previous pixel in raster order, whatever the row
        const int n = dec_run<V2>(e, ptype, c);
        if (n <= 0) return;
        PROF_T0
        if (ptype == 0 || ptype == 1) {
            for (int i = lane; i < n; i += 32) {
                const IPos q = ipos_add(p, i, X);
                if (q.y < Y) store_px(frame, g, q.x, q.y, c);
            }
            p = ipos_add(p, n, X);
            lastv = c;
        } else {
            // the run's last pixel is the next literal's context: it is taken from the lane that produced it, not read back
            uint32_t myv = 0;
            if (n < X && (ptype == 2 || ptype == 5)) {  // sources lie strictly before the run
                for (int i = lane; i < n; i += 32) {
                    const IPos q = ipos_add(p, i, X);
                    if (q.y < Y) {
                        myv = ptype == 2 ? load_px(frame, g, q.x, q.y - 1) : tl_at(frame, g, q, padded);
                        store_px(frame, g, q.x, q.y, myv);
                    }
                }
                lastv = __shfl_sync(0xFFFFFFFFu, myv, (n - 1) & 31);
            } else {  // gradient chains through the left pixel; tiny frames may read their own run
                if (lane == 0) {
                    IPos q = p;
                    uint32_t left = lastv, ptop = 0;  // the pixel before q in raster order; the pixel above it
                    bool have_ptop = false;
                    for (int i = 0; i < n && q.y < Y; i++) {
                        if (ptype == 2) myv = load_px(frame, g, q.x, q.y - 1);
                        else if (ptype == 5) myv = tl_at(frame, g, q, padded);
                        else {
                            const uint32_t top = load_px(frame, g, q.x, q.y - 1);
                            const uint32_t tlv = (have_ptop && q.x > 0) ? ptop : tl_at(frame, g, q, padded);
                            myv = grad_px(left, top, tlv);
                            ptop = top;
                            have_ptop = true;
                        }
                        store_px(frame, g, q.x, q.y, myv);
                        left = myv;
                        q = ipos_add(q, 1, X);
                    }
                }
                lastv = __shfl_sync(0xFFFFFFFFu, myv, 0);
            }
            __syncwarp();
            p = ipos_add(p, n, X);
        }
        e.lastpx = lastv;
        PROF_ADD(c_ifill)
    }
}


// ---- I frame through two row buffers in shared memory -------------------------------------------------------------------
// Every predictor of an I frame reaches at most one row up (the corner pixel at x == 0 reaches the tail of row y - 2, which
// still sits in the buffer row y is about to reuse), so the current and the previous row are kept in shared memory: run
// fills read and write shared memory only, and a row goes to the frame in one coalesced pass when the raster leaves it.
// Needs X >= 256 (a run is at most 255 pixels: it touches at most two rows) and 8 * X bytes of shared memory; other
// frames take decode_i above.
__device__ __forceinline__ void irow_flush(const Geo& g, uint8_t* frame, uint32_t row, int y, int lane) {
    if (g.bpp == 4 && (g.X & 3) == 0 && (g.pitch & 15) == 0) {
        uint4* dst = reinterpret_cast<uint4*>(frame + (size_t)y * g.pitch);
        for (int i = lane; i < g.X / 4; i += 32) {
            uint4 v = lds128(row + 16u * (uint32_t)i);
            v.x |= 0xFF000000u; v.y |= 0xFF000000u; v.z |= 0xFF000000u; v.w |= 0xFF000000u;
            dst[i] = v;
        }
    } else {
        for (int x = lane; x < g.X; x += 32) store_px(frame, g, x, y, lds32(row + 4u * (uint32_t)x));
    }
}
// commands to the reconstruction warp (uint4): x = op | ..., see recon_loop
constexpr uint32_t RQ_BEGIN = 1u, RQ_RUN = 2u, RQ_END = 3u, RQ_LOAD = 4u, RQ_IBEGIN = 5u, RQ_IRUN = 6u;
// Chain-warp side: symbols only.  Every run goes to the reconstruction warp (recon_loop: RQ_IBEGIN / RQ_IRUN), which
// owns the two row buffers, fills and flushes; the pixel before a literal comes back through fetch_lastpx.
__device__ __forceinline__ void rq_post(Ent& e, uint32_t x, uint32_t y, uint32_t z);
template <bool V2>
__device__ void decode_i_rows(const DecWork& w, Ent& e, int f, int lane) {
    const int X = w.g.X, Y = w.g.Y;
    int x = 0, y = 0;    // the next pixel
    int ptype = 0;
    int hdr = X + 1;
    bool header = true;
    rq_post(e, RQ_IBEGIN, (uint32_t)f, 0u);
    while (y < Y) {
        uint32_t c = 0;
        int n;
        if (header) {  // first row and one pixel: (rgb, n) pairs, lengths in ntab[0] (screencap.cpp:423-438)
            c = V2 ? rc_rgb(e) : dec_rgb(e);
            n = V2 ? rc_n(e, 0) : dec_n(e, 0);
            ptype = 0;
            hdr -= n;
            if (hdr <= 0) header = false;
        } else
            n = dec_run<V2>(e, ptype, c);
        if (n <= 0) break;
        rq_post(e, RQ_IRUN | ((uint32_t)ptype << 8) | ((uint32_t)n << 16), c, 0u);
        if (ptype) e.lp_wait = e.rposted;  // the run's last pixel comes from the reconstruction warp
        x += n;
        if (x >= X) {
            x -= X;
            y++;
        }
    }
}

// ---- helper warps: motion-vector copies -------------------------------------------------------------------
// A motion-vector block is a pure gather from the previous frame (screencap.cpp:1333-1368): nothing the entropy
// walk needs comes out of it, but done by the chain warp it costs three dependent memory round trips per block,
// and scrolling content produces thousands of such blocks per frame.  The chain warp therefore only decodes the
// block's symbols and posts a command; the other warps of the CTA execute the copies concurrently.
//   * All commands of a frame read frames before it and write that frame: they are independent of each other.
//   * The chain warp waits for the queue to drain (a) before the first map update of every frame (a copy resolves
//     its sources "as of the previous frame", which the two-deep map can only answer while that frame is the
//     newest) and (b) before it reads pixels itself (tile loads of pixel-coded blocks).
// Command (uint4): x = bi | frame << 16;  y = sub-rect inside the block, x1 | y1 << 4 | (x2-1) << 8 | (y2-1) << 12;
// z = (mx & 0xFFFF) | my << 16.  S_SYNC: head (commands published), next (commands taken), done, quit.
__device__ __forceinline__ uint32_t atom_add_shared(uint32_t a, uint32_t v) {
    uint32_t r;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(a), "r"(v) : "memory");
    return r;
}

template <bool SM>
__device__ void mv_copy(const DecWork& w, const BlockMap<SM>& map, int f, int bi, uint32_t rect, int mx, int my, int lane) {
    const Geo& g = w.g;
    const int by = bi / g.nbx, bx = bi - by * g.nbx;
    const int bx0 = bx * 16, by0 = by * 16;
    const int bw = min(16, g.X - bx0), bh = min(16, g.Y - by0);
    int x1 = bx0 + (int)(rect & 15u), y1 = by0 + (int)((rect >> 4) & 15u);
    int x2 = bx0 + (int)((rect >> 8) & 15u) + 1, y2 = by0 + (int)((rect >> 12) & 15u) + 1;
    if (x2 > bx0 + bw) x2 = bx0 + bw;  // blocks cut by the frame edge, and corrupt input
    if (y2 > by0 + bh) y2 = by0 + bh;
    if (x1 >= x2) x1 = x2 - 1;
    if (y1 >= y2) y1 = y2 - 1;
    uint8_t* frame = w.out + (size_t)f * g.frame_bytes;
    const int gx = min(max(x1 + mx, 0), g.X - 1), gy = min(max(y1 + my, 0), g.Y - 1);  // corrupt input guard
    const int cbx = gx >> 4, cby = gy >> 4;
    PixSrc sq[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {  // the source rectangle spans at most 2 x 2 blocks of the previous frame
        const int qx = min(cbx + (q & 1), g.nbx - 1), qy = min(cby + (q >> 1), g.nby - 1);
        sq[q] = resolve_src(w, src_before(map.get(qy * g.nbx + qx), f));
    }
    const PixSrc sP = resolve_src(w, src_before(map.get(bi), f));
    // eight unconditional loads from clamped coordinates (all in flight together), flat colours patched in afterwards
    uint32_t px[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
        const int p = lane + 32 * u, xx = p & 15, yy = p >> 4;
        const int x = min(bx0 + xx, g.X - 1), y = min(by0 + yy, g.Y - 1);
        const bool in = x >= x1 && x < x2 && y >= y1 && y < y2;
        const int sx = in ? min(max(x + mx, 0), g.X - 1) : x, sy = in ? min(max(y + my, 0), g.Y - 1) : y;
        const int q = ((sx >> 4) > cbx ? 1 : 0) + ((sy >> 4) > cby ? 2 : 0);
        const PixSrc& sS = !in ? sP : q == 0 ? sq[0] : q == 1 ? sq[1] : q == 2 ? sq[2] : sq[3];  // outside the sub-rect: the previous frame, same place
        const uint32_t v = load_px(sS.base, g, sx, sy);
        px[u] = sS.flat ? sS.clr : v;
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
        const int p = lane + 32 * u, xx = p & 15, yy = p >> 4;
        if (xx < bw && yy < bh) store_px(frame, g, bx0 + xx, by0 + yy, px[u]);
    }
    __syncwarp();
    if (lane == 0) {
        // the block now belongs to frame f.  Readers that still see the old word get the same answer: they ask for the
        // source as of frame f - 1 (other copies of this frame) or come after the chain warp's drain (tile loads, frame f + 1)
        map.write(bi, (uint32_t)f);
        w.upd[(size_t)f * g.nb + bi] = 1;
    }
}

template <bool SM>
__device__ void helper_loop(const DecWork& w, const BlockMap<SM>& map, uint32_t sb, int lane) {
    const uint32_t sy = sb + S_SYNC;
    for (;;) {
        uint32_t idx = 0;
        if (lane == 0) idx = atom_add_shared(sy + 4, 1u);
        idx = __shfl_sync(0xFFFFFFFFu, idx, 0);
        // the slot carries the sequence number of the command it holds (written last, as one 64-bit word with the first
        // argument): poll the slot itself.  Idle most of the time on sparse content: back off.
        uint4 cmd;
        const uint32_t slot = sb + S_RING + 16u * (idx % RING);
        for (uint32_t ns = 64;;) {
            asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(cmd.x), "=r"(cmd.w) : "r"(slot) : "memory");
            if (cmd.w == idx + 1u) {
                asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(cmd.y), "=r"(cmd.z) : "r"(slot + 8u) : "memory");
                break;
            }
            if (ldv_shared(sy + 12)) return;  // quit is only raised after a drain: nothing posted is left behind
            __nanosleep(ns);
            ns = min(ns * 2u, (uint32_t)SCPR_COPY_BACKOFF);
        }
        mv_copy<SM>(w, map, (int)(cmd.x >> 16), (int)(cmd.x & 0xFFFFu), cmd.y, (int)(int16_t)(cmd.z & 0xFFFFu), (int)(int16_t)(cmd.z >> 16), lane);
        __threadfence_block();
        __syncwarp();
        if (lane == 0) atom_add_shared(sy + 8, 1u);
    }
}
// chain warp side.  Completion may be out of order by at most the number of helper warps, so a ring slot is
// reused only while fewer than RING - DEC_WARPS commands are outstanding.
__device__ __forceinline__ void cmd_push(Ent& e, uint32_t x, uint32_t y, uint32_t z) {
    const uint32_t sy = e.sb + S_SYNC;
    if ((e.head & 15u) == 0u)  // ring space, checked once per 16 commands
        while (e.head - ldv_shared(sy + 8) > (uint32_t)(RING - DEC_WARPS - 16)) __nanosleep(100);
    const uint32_t slot = e.sb + S_RING + 16u * (e.head % RING);
    e.head++;
    asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(slot + 8u), "r"(y), "r"(z) : "memory");
    asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(slot), "r"(x), "r"(e.head) : "memory");
    stv_shared(sy, e.head);  // the count, for "have all copies landed" (cmd_drain, recon_loop)
}
// every frame below f is final (its copies have drained): tell the host, which overlaps the gather of untouched blocks
// and the download of finished frames with the rest of the chain
__device__ __forceinline__ void publish_progress(const DecWork& w, int f, int lane) {
    if (w.progress && lane == 0) {
        __threadfence_system();
        w.progress[blockIdx.x] = f;
    }
}
__device__ __forceinline__ void cmd_drain(Ent& e) {
    const uint32_t sy = e.sb + S_SYNC;
    if (ldv_shared(sy + 8) != e.head) {
        PROF_T0
        while (ldv_shared(sy + 8) != e.head) __nanosleep(100);
        PROF_ADD(c_drain)
    }
    __threadfence_block();
}

__device__ __forceinline__ void rq_post(Ent& e, uint32_t x, uint32_t y, uint32_t z) {
    const uint32_t rs = e.sb + S_RSYNC;
    if ((e.rposted & 63u) == 0u)  // ring space, checked once per 64 commands
        while (e.rposted - ldv_shared(rs + 4) > (uint32_t)(RQ - 64)) {
        }
    // the slot carries its own sequence number: the reconstruction warp polls the slot itself.  Arguments first, then the
    // 64-bit {command, sequence number} word it polls (volatile stores of one thread, in program order)
    const uint32_t slot = e.sb + S_RQ + 16u * (e.rposted % RQ);
    e.rposted++;
    asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(slot + 8u), "r"(y), "r"(z) : "memory");
    asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(slot), "r"(x), "r"(e.rposted) : "memory");
}
// everything posted to the reconstruction warp has been executed (its frame writes and map updates are in place)
__device__ __forceinline__ void rq_drain(Ent& e) {
    const uint32_t rs = e.sb + S_RSYNC;
    while (ldv_shared(rs + 4) != e.rposted) __nanosleep(50);
    __threadfence_block();
}

// ---- P frame (DecompressP, screencap.cpp:1275-1432) ------------------------------------------------
__device__ __forceinline__ uint32_t tile_at(uint32_t tb, int ty, int tx) { return lds32(tb + (uint32_t)(ty * 17 + tx) * 4u); }

template <bool SM, bool V2>
__device__ void decode_p(const DecWork& w, const BlockMap<SM>& map, Ent& e, uint8_t* frame, int f, int lane) {
    const Geo& g = w.g;
    const uint32_t btsb = e.sb + S_BTS;
#ifdef SCPR_PROF
    const long long thdr__ = clock64();
#endif
    int t0 = sym_fx<V2, 256, CX_XX - CX_NTAB>(e);
    const int xx1 = (sym_fx<V2, 256, CX_XX - CX_NTAB>(e) << 8) + t0;
    t0 = sym_fx<V2, 256, CX_XX - CX_NTAB>(e);
    int xx2 = (sym_fx<V2, 256, CX_XX - CX_NTAB>(e) << 8) + t0;
    if (xx2 >= g.nb) xx2 = g.nb - 1;  // corrupt input guard
    // block types of [xx1, xx2] as (type, run) pairs (screencap.cpp:1306-1313)
    for (int x = xx1; x <= xx2;) {
        const int c = sym_fx<V2, 5, CX_BT - CX_NTAB>(e);
        const int n = sym_fx<V2, 256, CX_NTAB2 - CX_NTAB>(e);
        if (n <= 0) break;
        for (int i = lane; i < n && x + i < g.nb; i += 32) sts8(btsb + x + i, c);
        x += n;
    }
    __syncwarp();
#ifdef SCPR_PROF
    e.c_hdr += clock64() - thdr__;
#endif
    rq_drain(e);   // the blocks and ...
    cmd_drain(e);  // ... the copies of the previous frame are complete before this frame touches the map
    publish_progress(w, f, lane);
    e.lastpx = 0;  // cx = cx1 = 0, screencap.cpp:1319
    e.lp_wait = 0;
    int lastmx = 0, lastmy = 0;
    // visit changed blocks only: 32 block types per step, ballot, iterate the set bits
    for (int b0 = xx1 & ~31; b0 <= xx2; b0 += 32) {
      const int myb = b0 + lane;
      uint32_t chm = __ballot_sync(0xFFFFFFFFu, myb >= xx1 && myb <= xx2 && lds8(btsb + myb) != 0);
      while (chm) {
        const int bi = b0 + __ffs(chm) - 1;
        chm &= chm - 1;
        const int bt = (int)lds8(btsb + bi);
        PROF_CNT(n_blocks)
        if ((bt - 1) & 2) {
            // ---- motion-vector block: decode its symbols, post the copy (screencap.cpp:1333-1368).  The helper that
            // executes it clips the rectangle to the block and records the block's new owner in the map.
            PROF_T0
            uint32_t rect = 0xFF00u;  // the whole block
            if ((bt - 1) & 1) rect = dec_rect<V2>(e);
            int mx = lastmx, my = lastmy;
            if (V2) {  // no repeat flag before v3, vectors offset by the stream's own motion range (screencap.cpp:1358-1361)
                mx = rc_fx<13>(e) - e.msr_x;
                my = rc_fx<14>(e) - e.msr_y;
            } else if (!dec_bool(e)) {
                mx = dec_fxc<512, CX_MV - CX_NTAB + 0>(e) - 256;
                my = dec_fxc<512, CX_MV - CX_NTAB + 1>(e) - 256;
            }
            lastmx = mx; lastmy = my;
            cmd_push(e, (uint32_t)bi | ((uint32_t)f << 16), rect, ((uint32_t)mx & 0xFFFFu) | ((uint32_t)my << 16));
            PROF_ADD(c_mv)
            continue;
        }
        const int by = bi / g.nbx, bx = bi - by * g.nbx;
        const int bx0 = bx * 16, by0 = by * 16;
        const int bw = min(16, g.X - bx0), bh = min(16, g.Y - by0);
        int x1 = bx0, y1 = by0, x2 = bx0 + bw, y2 = by0 + bh;
        // ---- pixel-coded block: the chain warp decodes its symbols, the reconstruction warp builds its pixels ----
        rq_post(e, RQ_LOAD | ((uint32_t)bi << 8), (uint32_t)f, 0u);  // the tile load starts while the sub-rect is still being decoded
        if ((bt - 1) & 1) {
            const uint32_t rc4 = dec_rect<V2>(e);
            x1 = bx0 + (int)(rc4 & 15u);
            y1 = by0 + (int)((rc4 >> 4) & 15u);
            x2 = bx0 + (int)((rc4 >> 8) & 15u) + 1;
            y2 = by0 + (int)((rc4 >> 12) & 15u) + 1;
            if (x2 > bx0 + bw) x2 = bx0 + bw;  // corrupt input guards
            if (y2 > by0 + bh) y2 = by0 + bh;
            if (x1 >= x2) x1 = x2 - 1;
            if (y1 >= y2) y1 = y2 - 1;
        }
        rq_post(e, RQ_BEGIN, 0u, (uint32_t)(x1 - bx0) | ((uint32_t)(y1 - by0) << 4) | ((uint32_t)(x2 - 1 - bx0) << 8) | ((uint32_t)(y2 - 1 - by0) << 12));
        {
            PROF_T0
            int pos = 0, ptype = 0;
            const int npx = (x2 - x1) * (y2 - y1);
            while (pos < npx) {
                uint32_t c = 0;
                int n = dec_run<V2>(e, ptype, c);
                if (n > npx - pos) n = npx - pos;
                if (n <= 0) break;
#ifdef SCPR_PROF
                rq_post(e, RQ_RUN | ((uint32_t)ptype << 8) | ((uint32_t)n << 16), c, (uint32_t)clock64());
#else
                rq_post(e, RQ_RUN | ((uint32_t)ptype << 8) | ((uint32_t)n << 16), c, 0u);
#endif
                if (ptype) e.lp_wait = e.rposted;  // a predicted run: its last pixel is known once this command is done
                pos += n;
            }
            PROF_ADD(c_runs)
        }
        rq_post(e, RQ_END, 0u, 0u);
      }
    }
}

// ---- reconstruction warp ---------------------------------------------------------------------------------------------------
// Builds the pixels of pixel-coded P blocks from the commands of the chain warp: LOAD (block, frame) loads the
// 17 x 17 tile -- tile[1+yy][1+xx] = block pixel, initially the previous frame's block (so pixels of type 3, and the part
// of a partial block outside the sub-rect, are already in place); row 0 / column 0 = the neighbours above / left in the
// current frame --, BEGIN (sub-rect) sets the cursor, RUN (type, length, colour) fills a run, END writes the block to the frame and records its owner.
// After every command it publishes the command count and, after a run, the run's last pixel (the context of a following
// literal, fetch_lastpx).  Commands are executed in order, so the map and the frame see the blocks in bitstream order;
// before a tile is loaded the motion-vector copies posted so far must have landed (neighbours may be such blocks).
template <bool SM>
__device__ void recon_loop(const DecWork& w, const BlockMap<SM>& map, uint32_t sb, int lane) {
    const Geo& g = w.g;
    const uint32_t tb = sb + S_TILE, rs = sb + S_RSYNC, sy = sb + S_SYNC;
    // block state
    int bi = 0, f = 0, bx0 = 0, by0 = 0, bw = 16, bh = 16, sw = 1, pos = 0, xx0 = 0, yy0 = 0, ox = 1, oy = 1;
    uint32_t swinv = 65536u, sw1inv = 32768u, ca = tb;
    uint8_t* frame = w.out;
    // I frame state (row buffers, see irow_flush)
    const bool padded = ((g.X * 3 + 3) & ~3) != g.X * 3;
    const uint32_t rb = sb + w.irows, rowbytes = (uint32_t)g.X * 4u;
    int ix = 0, iy = 0;
    uint32_t ilast = 0, rdone = 0;
#ifdef SCPR_PROF
    long long r_pub = 0, r_run = 0, r_nrun = 0, r_load = 0, r_nload = 0, r_end = 0, r_idle = 0, r_qlat = 0, r_qlong = 0, r_qidle = 0, r_qprev[4] = {0, 0, 0, 0}, r_spins = 0, r_t[6] = {0, 0, 0, 0, 0, 0}, r_n[6] = {0, 0, 0, 0, 0, 0};
    uint32_t prev_op__ = 0, last_spins__ = 0;
#endif
    for (uint32_t ridx = 0;; ridx++) {
#ifdef SCPR_PROF
        const long long ti__ = clock64();
#endif
        uint4 cmd;
        const uint32_t slot = sb + S_RQ + 16u * (ridx % RQ);
        for (uint32_t spins = 0;; spins++) {
            // both halves every time, {command, sequence number} first: the writer stores the arguments before that word, so when
            // the sequence number read here is the awaited one, the arguments read after it are the command's (one round trip
            // instead of two on the way to every run)
            asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(cmd.x), "=r"(cmd.w) : "r"(slot) : "memory");
            asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(cmd.y), "=r"(cmd.z) : "r"(slot + 8u) : "memory");
            if (cmd.w == ridx + 1u) {
#ifdef SCPR_PROF
                last_spins__ = spins;
#endif
                break;
            }
#if SCPR_RECON_NAP > 0
            if (spins >= 2u) __nanosleep(SCPR_RECON_NAP);  // an idle poll loop takes shared-memory bandwidth from the chain warp
#endif
            if (spins == 0u && rdone != ridx) {  // nothing waiting: let the chain warp see how far this warp has come (drains)
                __threadfence_block();
                stv_shared(rs + 4, ridx);
                rdone = ridx;
            }
            if ((spins & 31u) == 31u) {
#ifdef SCPR_PROF
                if (ldv_shared(sy + 12) && lane == 0)
                    printf("[rec prof] runs %lld: detect->publish %.0f cyc, whole run %.0f cyc | loads %lld: %.0f cyc | ends %.0f cyc | idle %.1f Mcyc | post->detect %.0f cyc, %lld over 1000 (of those %lld had polled before, %lld polls; previous op run %lld load %lld begin %lld other %lld) | by type n/cyc: 0 %lld/%.0f 1 %lld/%.0f 2 %lld/%.0f 3 %lld/%.0f 4 %lld/%.0f 5 %lld/%.0f\n", r_nrun,
                           (double)r_pub / (double)max(1LL, r_nrun), (double)r_run / (double)max(1LL, r_nrun), r_nload, (double)r_load / (double)max(1LL, r_nload),
                           (double)r_end / (double)max(1LL, r_nload), r_idle * 1e-6, (double)r_qlat / (double)max(1LL, r_nrun), r_qlong, r_qidle, r_spins, r_qprev[0], r_qprev[1], r_qprev[2], r_qprev[3],
                           r_n[0], (double)r_t[0] / (double)max(1LL, r_n[0]), r_n[1], (double)r_t[1] / (double)max(1LL, r_n[1]), r_n[2], (double)r_t[2] / (double)max(1LL, r_n[2]),
                           r_n[3], (double)r_t[3] / (double)max(1LL, r_n[3]), r_n[4], (double)r_t[4] / (double)max(1LL, r_n[4]), r_n[5], (double)r_t[5] / (double)max(1LL, r_n[5]));
#endif
                if (ldv_shared(sy + 12)) return;
                if (spins > 2048) __nanosleep(spins > 65536 ? 400 : 50);
            }
        }
        const uint32_t op = cmd.x & 0xFFu;
#ifdef SCPR_PROF
        const long long td__ = clock64();
        r_idle += td__ - ti__;
#endif
        if (op == RQ_RUN) {
            const int ptype = (int)((cmd.x >> 8) & 0xFFu), n = (int)(cmd.x >> 16);
            const uint32_t c = cmd.y;
            const int xe = xx0 + n;
#ifdef SCPR_PROF
            {
                const uint32_t lat = (uint32_t)td__ - cmd.z;  // post -> detect
                r_qlat += lat;
                if (lat > 1000u) {
                    r_qlong++;
                    if (last_spins__ > 0) r_qidle++;
                    r_qprev[prev_op__ == RQ_RUN ? 0 : prev_op__ == RQ_LOAD ? 1 : prev_op__ == RQ_BEGIN ? 2 : 3]++;
                    r_spins += last_spins__;
                }
            }
#endif
            // The chain warp may be waiting for the run's last pixel (the context of a literal that follows), and it produces a
            // predicted run every ~500 cycles: this warp has to be quicker than that per run or the chain queues up behind it
            // (measured: 35 % of the chain's time went into that wait when a run cost 550 cycles here).
            if (xe <= sw && ptype != 4) {
                // The run stays in its row (the common case; at most 16 pixels): every source is one fixed step away -- left: the
                // pixel before the run; top: one tile row up; top-left: one row up, one left; type 3 keeps the previous frame's
                // pixel that is already in the tile.  Lane i computes pixel i, so the run's last pixel is simply lane n - 1's
                // value: that lane publishes it, then the lanes store -- no separate look-up, no division, no shuffle.
                const uint32_t a = ca + 4u * (uint32_t)min(lane, n - 1);
                uint32_t v = c;
                if (ptype == 1) v = lds32(ca - 4u);
                else if (ptype == 2) v = lds32(a - 68u);
                else if (ptype == 5) v = lds32(a - 72u);
                else if (ptype == 3) v = lds32(a);
                if (lane == n - 1) asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(rs + 8u), "r"(ridx + 1u), "r"(v) : "memory");
#ifdef SCPR_PROF
                r_pub += clock64() - td__;
#endif
                if (ptype != 3 && lane < n) sts32(a, v);
                __syncwarp();
            } else {
                if (ptype != 4) {  // every source of a predicted pixel lies outside its run: the last pixel first, the fill follows
                    const int li = pos + n - 1, ly = (int)(((uint32_t)li * swinv) >> 16), lx = li - ly * sw;
                    uint32_t vlast = c;
                    if (ptype == 1) vlast = ly == yy0 ? tile_at(tb, oy + yy0, ox + xx0 - 1) : tile_at(tb, oy + ly, ox - 1);
                    else if (ptype == 2) vlast = tile_at(tb, oy + (lx >= xx0 ? yy0 : yy0 + 1) - 1, ox + lx);
                    else if (ptype == 3) vlast = tile_at(tb, oy + ly, ox + lx);
                    else if (ptype == 5) {
                        const int k = min((int)(((uint32_t)(n - 1) * sw1inv) >> 16) + 1, lx + 1);
                        vlast = tile_at(tb, oy + ly - k, ox + lx - k);
                    }
                    asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(rs + 8u), "r"(ridx + 1u), "r"(vlast) : "memory");
#ifdef SCPR_PROF
                    r_pub += clock64() - td__;
#endif
                }
                if (ptype == 4) {  // gradient chains through the left pixel: serial
                    if (lane == 0) {
                        int xx = xx0, yy = yy0;
                        for (int i = 0; i < n; i++) {
                            const uint32_t a = tb + (uint32_t)((oy + yy) * 17 + ox + xx) * 4u;
                            sts32(a, grad_px(lds32(a - 4), lds32(a - 68), lds32(a - 72)));
                            if (++xx == sw) {
                                xx = 0;
                                yy++;
                            }
                        }
                    }
                } else if (ptype != 3) {
                    for (int i = lane; i < n; i += 32) {
                        const int idx = pos + i;
                        const int yy = (int)(((uint32_t)idx * swinv) >> 16), xx = idx - yy * sw;
                        uint32_t v = c;
                        if (ptype == 1) {  // left: the pixel before the row segment the pixel lies in
                            v = yy == yy0 ? tile_at(tb, oy + yy0, ox + xx0 - 1) : tile_at(tb, oy + yy, ox - 1);
                        } else if (ptype == 2) {  // top: the pixel above the run's first row in this column
                            v = tile_at(tb, oy + (xx >= xx0 ? yy0 : yy0 + 1) - 1, ox + xx);
                        } else if (ptype == 5) {  // top-left: walk the diagonal until it leaves the run or the sub-rect
                            const int k = min((int)(((uint32_t)i * sw1inv) >> 16) + 1, xx + 1);
                            v = tile_at(tb, oy + yy - k, ox + xx - k);
                        }
                        sts32(tb + (uint32_t)((oy + yy) * 17 + ox + xx) * 4u, v);
                    }
                }
                __syncwarp();
                if (ptype == 4) {
                    const int li = pos + n - 1, ly = (int)(((uint32_t)li * swinv) >> 16);
                    const uint32_t vlast = tile_at(tb, oy + ly, ox + li - ly * sw);
                    asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(rs + 8u), "r"(ridx + 1u), "r"(vlast) : "memory");
                }
            }
            pos += n;
            xx0 = xe;
            ca += 4u * (uint32_t)n;
            if (xx0 >= sw) {  // next row(s)
                const int q = (int)(((uint32_t)xx0 * swinv) >> 16);
                xx0 -= q * sw;
                yy0 += q;
                ca = tb + (uint32_t)((oy + yy0) * 17 + ox + xx0) * 4u;
            }
        } else if (op == RQ_IRUN) {
            // ---- I frame run through the two row buffers: every predictor reaches at most one row up (the corner pixel
            // at x == 0 reaches the tail of row y - 2, which still sits in the buffer row y is about to reuse); a run is at
            // most 255 pixels and the frame at least 256 wide, so it touches at most two rows
            const int X = g.X, Y = g.Y;
            const int ptype = (int)((cmd.x >> 8) & 0xFFu), n = (int)(cmd.x >> 16);
            const uint32_t c = cmd.y;
            const uint32_t r0 = rb + (uint32_t)(iy & 1) * rowbytes, r1 = rb + (uint32_t)((iy + 1) & 1) * rowbytes;  // rows y (and y - 2) / y + 1 (and y - 1)
            uint32_t vlast = ptype == 0 ? c : ilast;  // type 1: the previous pixel in raster order, whatever the row
            if (ptype == 2 || ptype == 5) {  // the last pixel first (its sources lie before the run), then the fill
                int xl = ix + n - 1, yl = iy;
                uint32_t cur = r0, up = r1;
                if (xl >= X) {
                    xl -= X; yl++;
                    cur = r1; up = r0;
                }
                if (yl < Y) {
                    if (ptype == 2) vlast = lds32(up + 4u * (uint32_t)xl);
                    else if (xl > 0) vlast = lds32(up + 4u * (uint32_t)(xl - 1));
                    else vlast = padded ? tl_padded(frame, g, yl) : lds32(cur + 4u * (uint32_t)(X - 1));
                } else
                    vlast = 0;
            }
            if (ptype != 4) asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(rs + 8u), "r"(ridx + 1u), "r"(vlast) : "memory");
            if (ptype == 0 || ptype == 1) {
                for (int i = lane; i < n; i += 32) {
                    const int xi = ix + i;
                    if (xi < X) sts32(r0 + 4u * (uint32_t)xi, vlast);
                    else if (iy + 1 < Y) sts32(r1 + 4u * (uint32_t)(xi - X), vlast);
                }
            } else if (ptype == 2 || ptype == 5) {
                for (int i = lane; i < n; i += 32) {
                    int xi = ix + i, yi = iy;
                    uint32_t cur = r0, up = r1;
                    if (xi >= X) {
                        xi -= X; yi++;
                        cur = r1; up = r0;
                    }
                    if (yi < Y) {
                        uint32_t v;
                        if (ptype == 2) v = lds32(up + 4u * (uint32_t)xi);
                        else if (xi > 0) v = lds32(up + 4u * (uint32_t)(xi - 1));
                        else v = padded ? tl_padded(frame, g, yi) : lds32(cur + 4u * (uint32_t)(X - 1));  // tail of row yi - 2
                        sts32(cur + 4u * (uint32_t)xi, v);
                    }
                }
            } else {  // gradient chains through the left pixel: one lane
                uint32_t myv = 0;
                if (lane == 0) {
                    int xi = ix, yi = iy;
                    uint32_t cur = r0, up = r1, left = ilast;
                    for (int i = 0; i < n && yi < Y; i++) {
                        const uint32_t top = lds32(up + 4u * (uint32_t)xi);
                        const uint32_t tlv = xi > 0 ? lds32(up + 4u * (uint32_t)(xi - 1))
                                                    : (padded ? tl_padded(frame, g, yi) : lds32(cur + 4u * (uint32_t)(X - 1)));
                        myv = grad_px(left, top, tlv);
                        sts32(cur + 4u * (uint32_t)xi, myv);
                        left = myv;
                        if (++xi == X) {
                            xi = 0; yi++;
                            const uint32_t t = cur; cur = up; up = t;
                        }
                    }
                }
                vlast = __shfl_sync(0xFFFFFFFFu, myv, 0);
                asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(rs + 8u), "r"(ridx + 1u), "r"(vlast) : "memory");
            }
            __syncwarp();
            ilast = vlast;
            ix += n;
            if (ix >= X) {  // the raster left row y: it is complete
                if (iy < Y) irow_flush(g, frame, r0, iy, lane);
                __syncwarp();
                ix -= X;
                iy++;
            }
        } else if (op == RQ_IBEGIN) {
            f = (int)cmd.y;
            frame = w.out + (size_t)f * g.frame_bytes;
            ix = iy = 0;
            ilast = 0;
        } else if (op == RQ_BEGIN) {  // the sub-rect the runs cover
            const int x1 = (int)(cmd.z & 15u), y1 = (int)((cmd.z >> 4) & 15u), x2 = (int)((cmd.z >> 8) & 15u) + 1;
            sw = x2 - x1;
            swinv = (65536u + (uint32_t)sw - 1) / (uint32_t)sw;    // idx / sw == (idx * swinv) >> 16 for idx < 272
            sw1inv = (65536u + (uint32_t)sw) / (uint32_t)(sw + 1);  // same for sw + 1
            ox = 1 + x1; oy = 1 + y1;
            pos = 0; xx0 = 0; yy0 = 0;
            ca = tb + (uint32_t)(oy * 17 + ox) * 4u;
        } else if (op == RQ_LOAD) {  // block and frame: load the tile
            bi = (int)(cmd.x >> 8);
            f = (int)cmd.y;
            frame = w.out + (size_t)f * g.frame_bytes;
            const int by = bi / g.nbx, bx = bi - by * g.nbx;
            bx0 = bx * 16; by0 = by * 16;
            bw = min(16, g.X - bx0); bh = min(16, g.Y - by0);
            // neighbours may be motion-vector blocks of this frame: every copy posted so far must have landed
            while (ldv_shared(sy + 8) != ldv_shared(sy)) __nanosleep(50);
            __threadfence_block();
            const int bA = bi - g.nbx - 1, bT = bi - g.nbx, bL = bi - 1;
            const PixSrc sA = resolve_src(w, (bx > 0 && by > 0) ? src_now(map.get(bA)) : SRC_PREV0);
            const PixSrc sT = resolve_src(w, by > 0 ? src_now(map.get(bT)) : SRC_PREV0);
            const PixSrc sL = resolve_src(w, bx > 0 ? src_now(map.get(bL)) : SRC_PREV0);
            const PixSrc sP = resolve_src(w, src_before(map.get(bi), f));
            // block rows 2u + (lane >> 4), column lane & 15; then the row above (17 pixels with the corner) and the column to
            // the left: ten loads per lane with additive addressing
            // All ten loads are issued unconditionally from clamped coordinates and masked afterwards: with a branch around
            // each of them they went to memory one after the other (ten round trips instead of one).
            uint32_t tv[10];
            const int cxl = lane & 15, ryl = lane >> 4;
            const int xc = min(bx0 + cxl, g.X - 1);
#pragma unroll
            for (int u = 0; u < 8; u++) tv[u] = load_px(sP.base, g, xc, min(by0 + 2 * u + ryl, g.Y - 1));
            const PixSrc& sH = lane == 0 ? sA : sT;
            tv[8] = load_px(sH.base, g, min(max(bx0 + lane - 1, 0), g.X - 1), max(by0 - 1, 0));
            tv[9] = load_px(sL.base, g, max(bx0 - 1, 0), min(by0 + (lane & 15), g.Y - 1));
#pragma unroll
            for (int u = 0; u < 8; u++) tv[u] = (cxl < bw && 2 * u + ryl < bh) ? (sP.flat ? sP.clr : tv[u]) : 0u;
            tv[8] = (lane < 17 && by > 0 && (lane > 0 ? lane - 1 < bw : bx > 0)) ? (sH.flat ? sH.clr : tv[8]) : 0u;
            tv[9] = (lane < 16 && bx > 0 && lane < bh) ? (sL.flat ? sL.clr : tv[9]) : 0u;
            const uint32_t a0 = tb + (uint32_t)((1 + (lane >> 4)) * 17 + 1 + (lane & 15)) * 4u;
#pragma unroll
            for (int u = 0; u < 8; u++) sts32(a0 + (uint32_t)u * 136u, tv[u]);
            if (lane < 17) sts32(tb + (uint32_t)lane * 4u, tv[8]);
            if (lane < 16) sts32(tb + (uint32_t)(1 + lane) * 68u, tv[9]);
            __syncwarp();
        } else {  // RQ_END: write the whole block and hand its ownership to this frame
            if (bw == 16) {
                for (int p = lane; p < 16 * bh; p += 32) store_px(frame, g, bx0 + (p & 15), by0 + (p >> 4), tile_at(tb, 1 + (p >> 4), 1 + (p & 15)));
            } else {
                for (int p = lane; p < bw * bh; p += 32) {
                    const int xx = p % bw, yy = p / bw;
                    store_px(frame, g, bx0 + xx, by0 + yy, tile_at(tb, 1 + yy, 1 + xx));
                }
            }
            __syncwarp();
            if (lane == 0) {
                map.write(bi, (uint32_t)f);
                w.upd[(size_t)f * g.nb + bi] = 1;
            }
            __syncwarp();
        }
#ifdef SCPR_PROF
        prev_op__ = op;
        if (op == RQ_RUN) { r_run += clock64() - td__; r_nrun++; const int pt__ = min((int)((cmd.x >> 8) & 0xFFu), 5); r_t[pt__] += clock64() - td__; r_n[pt__]++; }
        else if (op == RQ_LOAD) { r_load += clock64() - td__; r_nload++; }
        else if (op == RQ_END) r_end += clock64() - td__;
#endif
        // executed-command count: the chain warp needs it when it drains (always right after an END) and, coarsely, for ring space
        if ((op != RQ_RUN && op != RQ_IRUN) || (ridx & 15u) == 15u) {
            __threadfence_block();
            stv_shared(rs + 4, ridx + 1);
            rdone = ridx + 1;
        }
    }
}

// One CTA per chain: warp 0 walks the chain, the other warps copy motion-vector blocks (warp 4 shares warp 0's
// scheduler and leaves at once).
template <bool SM, bool V2>
__global__ void __launch_bounds__(32 * DEC_WARPS, 1) k_dec_chain(DecWork w) {
    extern __shared__ __align__(16) uint8_t s_mem[];
    uint32_t sb = (uint32_t)__cvta_generic_to_shared(s_mem);
    asm volatile("mov.u32 %0, %0;" : "+r"(sb));  // keep the base in a register: otherwise it is rematerialised (S2R) at every use
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const DecChain ch = w.chains[blockIdx.x];
    const Geo& g = w.g;
    BlockMap<SM> map;
    map.sbase = sb + s_map_off(g.nb);
    map.g = SM ? nullptr : w.gmap + (size_t)blockIdx.x * g.nb;
    if (threadIdx.x < 4) {
        sts32(sb + S_SYNC + 4u * threadIdx.x, 0u);
        sts32(sb + S_RSYNC + 4u * threadIdx.x, 0u);
    }
    for (int i = threadIdx.x; i < RQ; i += 32 * DEC_WARPS) sts128(sb + S_RQ + 16u * i, 0u, 0u, 0u, 0u);  // no slot carries a sequence number yet
    for (int i = threadIdx.x; i < RING; i += 32 * DEC_WARPS) sts128(sb + S_RING + 16u * i, 0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < g.nb; i += 32 * DEC_WARPS) map.set(i, 0xFFFFFFFFu);  // everything lives in prev0
    __threadfence_block();
    __syncthreads();
    if (warp != 0) {
        if (warp == 1) recon_loop<SM>(w, map, sb, lane);
        else if (warp != 4) helper_loop<SM>(w, map, sb, lane);
        return;
    }
    Ent e;
    e.m = reinterpret_cast<ModelState*>(w.states + (size_t)ch.state * sizeof(ModelState));
    e.f0 = w.f0;
    e.lastpx = 0;
    e.head = 0;
    e.rposted = 0;
    e.lp_wait = 0;
    e.nleft = RANS_BLOCK;
    e.x = 0;
    e.w0 = e.w1 = e.k8 = 0;
    e.wp = nullptr;
    e.wend = reinterpret_cast<const uint32_t*>(w.stream + w.stream_bytes);
    e.sb = sb;
    e.lane = lane;
    e.v2 = reinterpret_cast<uint32_t*>(e.m);
    e.msr_x = w.msr_x;
    e.msr_y = w.msr_y;
    e.rc_code = e.rc_range = 0;
    e.rc_p = nullptr;
#if defined(SCPR_PROF) || defined(SCPR_LPSTAT)
    const long long tk0 = clock64();
#endif
    // fixed tables and context kinds of the chain's model state -> shared memory (a renewing first frame overwrites them)
    if (!V2)
    for (int t = 0; t < NUM_FIXED_CX; t++) {
        const FixedState& g0 = e.m->fx[t];
        const int off = fx_off(t), nsym = fixed_nsym(CX_NTAB + t);
        for (int i = lane; i < nsym; i += 32) {
            sts16(sb + S_CNT + (off + i) * 2, g0.cnt[i]);
            sts32(sb + S_FC + (off + i) * 4, ((uint32_t)g0.freq[i] << 16) | g0.cum[i]);
        }
    }
    if (!V2) {
        for (int i = lane; i < NUM_COLOR_CX / 4; i += 32) sts32(sb + S_KMAP + 4 * i, reinterpret_cast<const uint32_t*>(e.m->kmap)[i]);
        for (int i = lane; i < NCACHE; i += 32) sts32(sb + S_CTAG + 4 * i, 0xFFFFFFFFu);
        __syncwarp();
        for (int t = 0; t < NUM_FIXED_CX; t++) fixed_rebuild(sb, t, lane, false);
    }
    for (int f = ch.first; f < ch.first + ch.count; f++) {
        const DecFrame df = w.frames[f];
        uint8_t* frame = w.out + (size_t)f * g.frame_bytes;
        if (df.kind == DK_PSAME) continue;  // every block keeps its source (memcpy(pDst, prev), screencap.cpp:1288-1291)
        if (df.kind == DK_FLAT || df.kind == DK_I) {
            if (V2 && (df.kind == DK_I || df.renew)) rc_renew(e);
            if (!V2 && (df.kind == DK_I || df.renew)) {  // RenewI
                for (int i = lane; i < NUM_COLOR_CX; i += 32) e.m->color[i].kind = 0;
                for (int i = lane; i < NUM_COLOR_CX / 4; i += 32) sts32(sb + S_KMAP + 4 * i, 0);
                for (int i = lane; i < NCACHE; i += 32) sts32(sb + S_CTAG + 4 * i, 0xFFFFFFFFu);  // cached contexts are void
                for (int t = 0; t < NUM_FIXED_CX; t++) {  // FixedSizeRansCtx::renew, ans_contexts.h:1114-1131
                    const int off = fx_off(t), nsym = fixed_nsym(CX_NTAB + t);
                    const int fr = PROB_SCALE / nsym, c0 = fr - (fr >> 1);
                    for (int i = lane; i < nsym; i += 32) {
                        sts16(sb + S_CNT + (off + i) * 2, c0);
                        sts32(sb + S_FC + (off + i) * 4, ((uint32_t)fr << 16) | (uint32_t)(fr * i));
                    }
                }
                __syncwarp();
                for (int t = 0; t < NUM_FIXED_CX; t++) fixed_rebuild(sb, t, lane, false);
            }
            rq_drain(e);
            cmd_drain(e);
            publish_progress(w, f, lane);
            e.lp_wait = 0;
            const uint32_t code = (uint32_t)f | (df.kind == DK_FLAT ? SRC_FLAT : 0u);
            for (int i = lane; i < g.nb; i += 32) map.write(i, code);  // the whole frame is new
            __syncwarp();
            if (df.kind == DK_I) {
                e.lastpx = 0;
                if (V2)
                    rc_begin(e, w.stream + df.src_off + 1);
                else {
                    rd_seek(e, w.stream + df.src_off + 1);
                    e.nleft = RANS_BLOCK;
                    rdec_init(e);
                }
                if (w.irows) decode_i_rows<V2>(w, e, f, lane);
                else decode_i<V2>(w, e, frame, lane);
            }
            continue;
        }
        if (V2)
            rc_begin(e, w.stream + df.src_off + 1);
        else {
            rd_seek(e, w.stream + df.src_off + 1);
            e.nleft = RANS_BLOCK;
            rdec_init(e);
        }
        decode_p<SM, V2>(w, map, e, frame, f, lane);
        __threadfence_block();
    }
    rq_drain(e);
    cmd_drain(e);
    __threadfence();
    publish_progress(w, ch.first + ch.count, lane);
    stv_shared(sb + S_SYNC + 12, 1u);  // helpers leave
#ifdef SCPR_LPSTAT
    if (lane == 0) printf("[lp stat] chain %d frames %d: %.1f Mcyc, %u waits for a run's last pixel, %u polls, %u waits of > 10 polls\n", (int)blockIdx.x, ch.count, (clock64() - tk0) * 1e-6, e.s_waits, e.s_polls, e.s_long);
#endif
#ifdef SCPR_PROF
    if (lane == 0)
        printf("[dec prof] chain %d frames %d: total %.1f Mcyc | fixed %.1f Mcyc / %lld sym (%.0f cyc) | color %.1f / %lld (%.0f cyc; %lld serial, %lld rescales) | "
               "p-hdr %.1f | tile %.1f / %lld blocks | mv %.1f | runs(incl sym) %.1f | blkwr %.1f | ifill %.1f | rebuild %.1f / %lld | small %.1f / %lld flat %.1f / %lld raw %.1f / %lld | drain %.1f | cache misses %lld | lastpx waits %.1f / %lld (%lld polls, %lld waits of > 25 polls)\n",
               (int)blockIdx.x, ch.count, (clock64() - tk0) * 1e-6, e.c_fixed * 1e-6, e.n_fixed, (double)e.c_fixed / (double)max(1LL, e.n_fixed),
               e.c_color * 1e-6, e.n_color, (double)e.c_color / (double)max(1LL, e.n_color), e.n_gen, e.n_resc, e.c_hdr * 1e-6, e.c_tile * 1e-6,
               e.n_blocks, e.c_mv * 1e-6, e.c_runs * 1e-6, e.c_blkwr * 1e-6, e.c_ifill * 1e-6, e.c_rebuild * 1e-6, e.n_rebuild, e.c_small * 1e-6, e.n_small, e.c_flat * 1e-6, e.n_flat, e.c_raw * 1e-6, e.n_raw, e.c_drain * 1e-6, e.n_miss, e.c_lpwait * 1e-6, e.n_lpwait, e.n_lppoll, e.n_lplong);
#endif
    // leave the cached contexts, the fixed tables and the kinds behind for the next call (v2 tables are already in place)
    __syncwarp();
    if (!V2) {
    for (uint32_t h = 0; h < NCACHE; h++) cache_writeback(e.m, sb, lane, h);
    for (int t = 0; t < NUM_FIXED_CX; t++) {
        FixedState& g0 = e.m->fx[t];
        const int off = fx_off(t), nsym = fixed_nsym(CX_NTAB + t);
        uint32_t sum = 0;
        for (int i = lane; i < nsym; i += 32) {
            const uint32_t cn = lds16(sb + S_CNT + (off + i) * 2), fc = lds32(sb + S_FC + (off + i) * 4);
            g0.cnt[i] = (uint16_t)cn;
            g0.freq[i] = (uint16_t)(fc >> 16);
            g0.cum[i] = (uint16_t)(fc & 0xFFFFu);
            sum += cn;
        }
        sum = __reduce_add_sync(0xFFFFFFFFu, sum);
        if (lane == 0) {
            g0.cntsum = (int)sum;  // cntsum is the sum of the counters at all times
            g0.nsym = nsym;
        }
    }
    for (int i = lane; i < NUM_COLOR_CX / 4; i += 32) reinterpret_cast<uint32_t*>(e.m->kmap)[i] = lds32(sb + S_KMAP + 4 * i);
    }
}

// source frame of every block of every frame: thread per block, frames in order
__global__ void k_dec_sources(DecWork w) {  // frames [f_begin, f_end) of chain w.chain, resuming from the previous range
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= w.g.nb) return;
    int16_t* keep = w.fill_last + (size_t)w.chain * w.g.nb;
    int last = w.f_begin == w.chains[w.chain].first ? -1 : keep[b];
    for (int f = w.f_begin; f < w.f_end; f++) {
        const uint8_t k = w.frames[f].kind;
        if (k == DK_I || k == DK_FLAT || w.upd[(size_t)f * w.g.nb + b]) last = f;
        w.fill_src[(size_t)f * w.g.nb + b] = (int16_t)last;
    }
    keep[b] = (int16_t)last;
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
    if (strip >= (long)(w.f_end - w.f_begin) * strips_per_frame) return;
    const int fr = (int)(strip / strips_per_frame);
    const int f = w.f_begin + fr;
    const int s = (int)(strip - (long)fr * strips_per_frame);
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

// Bytes of one output row that carry pixels (the rest of `pitch` is the caller's padding, never written: the reference
// writes X * bytes-per-pixel per row, screencap.cpp:1713-1738)
static size_t row_bytes(const scpr_codec* c) { return (size_t)c->g.X * (c->rgb16 ? 2 : c->g.bpp); }

// device -> host copy of frames [a, b) of a batch that keeps the padding of the caller's rows
static cudaError_t download_frames(uint8_t* h, const uint8_t* d, size_t frame_bytes, int pitch, size_t rowb, int Y, int a, int b, cudaStream_t s) {
    if ((size_t)pitch == rowb) return cudaMemcpyAsync(h, d, (size_t)(b - a) * frame_bytes, cudaMemcpyDeviceToHost, s);
    return cudaMemcpy2DAsync(h, (size_t)pitch, d, (size_t)pitch, rowb, (size_t)(b - a) * Y, cudaMemcpyDeviceToHost, s);
}

// What one decode call works on.  Single-stream calls (clip_start == nullptr) continue the codec's decoder state -- open
// model chain, previous frame, flat-frame memory -- exactly as consecutive DecompressFrame calls do.  Multi-clip calls mark
// the first frame of every independent clip: each starts like a fresh codec (its first frame must be an I frame) and all
// chains of all clips run in ONE launch of the chain kernel, one CTA each.
struct DecJob {
    const uint8_t* stream;      // host: the frames' bitstreams back to back
    const uint32_t* sizes;
    const uint8_t* ftypes;
    int n;
    uint8_t* d_out;             // device: n frames back to back
    int pitch;
    const uint8_t* clip_start;  // n flags or nullptr
    uint8_t* const* h_clip;     // multi-clip: host destination of every clip (frames back to back), or nullptr
    uint8_t* h_out;             // single stream: host destination, or nullptr
    int* clip_result;           // multi-clip: per clip 1 / 0 / < 0
};

// n frames, bitstreams on the host, decoded frames to device memory (pitch bytes per row) and, when asked, on to host memory
static int decode_range(scpr_codec* c, const DecJob& j) {
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->st;
    const int n = j.n, pitch = j.pitch;
    if (n <= 0) return 1;
    const bool multi = j.clip_start != nullptr;
    Geo g = c->g;
    g.pitch = pitch;
    g.frame_bytes = (size_t)pitch * g.Y;
    if (pitch < g.X * g.bpp) return SCPR_E_PARAM;
    // ---- host plan: frame kinds, versions, chains (DecompressFrame, screencap.cpp:1695-1702, 1522-1557).  The plan works
    // on copies of the codec's decoder state; they are committed after the kernels have run (a rejected call must
    // not change the renew decision or the version of the next valid frame).
    bool created = multi ? false : c->dec_created, last_was_flat = multi ? false : c->dec_last_was_flat;
    int version = multi ? 0 : c->dec_version;
    uint8_t last_flat_clr[3] = {c->dec_last_flat_clr[0], c->dec_last_flat_clr[1], c->dec_last_flat_clr[2]};
    std::vector<DecFrame> frames(n);
    std::vector<DecChain> chains;
    std::vector<int> clip_of(multi ? n : 0), clip_first;
    uint64_t off = 0;
    int clip = -1;
    bool clip_dead = false;  // multi: this clip was refused, its frames are left alone
    for (int f = 0; f < n; f++) {
        DecFrame& d = frames[f];
        memset(&d, 0, sizeof(d));
        d.src_off = (uint32_t)off;
        d.size = j.sizes[f];
        const uint8_t* s = j.stream + off;
        off += j.sizes[f];
        if (j.sizes[f] < 1) return SCPR_E_PARAM;
        if (multi) {
            if (j.clip_start[f] || f == 0) {
                clip++;
                clip_first.push_back(f);
                created = last_was_flat = clip_dead = false;
                j.clip_result[clip] = 1;
            }
            clip_of[f] = clip;
            if (clip_dead) {
                d.kind = DK_PSAME;
                continue;
            }
        }
        if (!created) {
            int bad = 1;
            const int v = (s[0] >> 4) + 1;
            if (j.ftypes[f] > 0) bad = 0;  // P frame before any I frame
            else if (v < 2 || v > 4) bad = -v;  // BadVersionException (screencap.cpp:1589-1590)
            else if (v == 2 && (c->p.high_range_x < 1 || c->p.high_range_x > 256 || c->p.high_range_y < 1 || c->p.high_range_y > 256)) {
                set_error("v2 stream: motion range %u x %u outside 1..256", c->p.high_range_x, c->p.high_range_y);
                bad = SCPR_E_UNSUPPORTED;
            } else if (multi && version && v != version) {
                set_error("clips of one call must share a stream version (%d and %d)", version, v);
                bad = SCPR_E_UNSUPPORTED;
            }
            if (bad != 1) {
                if (!multi) return bad;
                j.clip_result[clip] = bad;
                clip_dead = true;
                d.kind = DK_PSAME;
                continue;
            }
            version = v;
            created = true;
        }
        const bool fresh_clip = multi && f == clip_first.back();
        if (j.ftypes[f]) {
            last_was_flat = false;
            d.kind = (s[0] & 1) ? DK_P : DK_PSAME;
            if (chains.empty()) chains.push_back(DecChain{f, 0, -2});  // continues the previous call's chain (single stream only: a clip starts with an I frame)
        } else if ((s[0] & 0x0F) == 1) {
            if (j.sizes[f] < 4) return SCPR_E_PARAM;
            d.kind = DK_FLAT;
            d.flat_clr = (uint32_t)s[1] | ((uint32_t)s[2] << 8) | ((uint32_t)s[3] << 16);
            d.renew = fresh_clip || !(last_was_flat && !memcmp(last_flat_clr, s + 1, 3));
            last_was_flat = true;
            memcpy(last_flat_clr, s + 1, 3);
            if (d.renew || chains.empty()) chains.push_back(DecChain{f, 0, d.renew ? -1 : -2});
        } else {
            last_was_flat = false;
            d.kind = DK_I;
            chains.push_back(DecChain{f, 0, -1});
        }
        chains.back().count = f - chains.back().first + 1;
    }
    if (off >= 0xFFFF0000ull) return SCPR_E_PARAM;
    const int n_chains = (int)chains.size();
    if (n_chains == 0) return 1;  // multi: every clip was refused
    TRY(ensure_dec_states(c, n_chains + 1));
    int new_cur_state = c->dec_cur_state;
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
        new_cur_state = chains.back().state;
    }
    // ---- device buffers ---------------------------------------------------------------------------
    TRY(c->dec_stream.ensure((size_t)off + 64));
    TRY(c->dec_desc.ensure((size_t)n * sizeof(DecFrame) + (size_t)n_chains * sizeof(DecChain)));
    // the block-source map sits in shared memory when it fits next to the tables, else in global memory
    // shared memory beyond the tables: block types (nb bytes), the block-source map when it fits (4 nb), two I-frame rows
    // when they fit (8 X; rows first -- I frames are where the time goes on big frames)
    const size_t smem_lim = 227 * 1024;
    const size_t smem_sm = ((size_t)s_map_off(g.nb) + (size_t)g.nb * 4 + 31) & ~(size_t)15, smem_gm = (size_t)s_map_off(g.nb) + 16;
    const size_t rows_bytes = g.X >= 256 ? (size_t)g.X * 8 : 0;
    bool map_shared = smem_sm + rows_bytes <= smem_lim;
    const bool rows = rows_bytes && (map_shared || smem_gm + rows_bytes <= smem_lim);
    if (!rows) map_shared = smem_sm <= smem_lim;
    const uint32_t irows_off = rows ? (uint32_t)(map_shared ? smem_sm : smem_gm) : 0u;
    const size_t smem = (map_shared ? smem_sm : smem_gm) + (rows ? rows_bytes : 0);
    if (smem > 227 * 1024) {
        set_error("frame has too many blocks for the decoder's shared-memory block map");
        return SCPR_E_PARAM;
    }
    const size_t map_bytes = map_shared ? 0 : (size_t)n_chains * g.nb * 4;
    const size_t upd_bytes = ((size_t)n * g.nb + 15) & ~(size_t)15, fsrc_bytes = ((size_t)n * g.nb * 2 + 15) & ~(size_t)15;
    TRY(c->dec_ws.ensure(map_bytes + upd_bytes + fsrc_bytes + (size_t)n_chains * g.nb * 2 + 64));
    if (c->dec_prev_pitch != pitch) {  // the previous frame is kept in output format: the reference takes the pitch per call (screencap.cpp:1705-1708)
        DBuf nb;
        TRY(nb.ensure(g.frame_bytes));
        if (c->dec_prev_pitch == 0 || !c->dec_prev.p)
            CK(cudaMemsetAsync(nb.p, 0, g.frame_bytes, st));
        else  // same pixels, new row pitch
            CK(cudaMemcpy2DAsync(nb.p, (size_t)pitch, c->dec_prev.p, (size_t)c->dec_prev_pitch, (size_t)g.X * g.bpp, (size_t)g.Y, cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
        c->dec_prev.release();
        c->dec_prev = nb;
        c->dec_prev_pitch = pitch;
    }
    CK(cudaMemcpyAsync(c->dec_stream.p, j.stream, (size_t)off, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync((uint8_t*)c->dec_stream.p + off, 0, 64, st));
    CK(cudaMemcpyAsync(c->dec_desc.p, frames.data(), (size_t)n * sizeof(DecFrame), cudaMemcpyHostToDevice, st));
    DecChain* d_chains = reinterpret_cast<DecChain*>((uint8_t*)c->dec_desc.p + (size_t)n * sizeof(DecFrame));
    CK(cudaMemcpyAsync(d_chains, chains.data(), (size_t)n_chains * sizeof(DecChain), cudaMemcpyHostToDevice, st));
    DecWork w;
    memset(&w, 0, sizeof(w));
    w.g = g;
    w.stream = (const uint8_t*)c->dec_stream.p;
    w.stream_bytes = (uint32_t)(((size_t)off + 64) & ~(size_t)3);
    w.frames = (const DecFrame*)c->dec_desc.p;
    w.chains = d_chains;
    w.states = (uint8_t*)c->dec_state.p;
    w.f0 = version == 3 ? 64 : 32;  // setCx6f0, screencap.cpp:1613-1614
    w.out = j.d_out;
    w.prev0 = (const uint8_t*)c->dec_prev.p;
    uint8_t* ws = (uint8_t*)c->dec_ws.p;
    w.gmap = (uint32_t*)ws;
    w.upd = ws + map_bytes;
    w.fill_src = (int16_t*)(ws + map_bytes + upd_bytes);
    w.fill_last = (int16_t*)(ws + map_bytes + upd_bytes + fsrc_bytes);
    w.n = n;
    w.msr_x = (int)c->p.high_range_x;
    w.msr_y = (int)c->p.high_range_y;
    w.irows = irows_off;
    w.progress = nullptr;
    const bool to_host = j.h_out || j.h_clip;
    const size_t rowb = (size_t)g.X * g.bpp;
    auto host_of = [&](int f) -> uint8_t* {  // where frame f goes on the host
        if (j.h_out) return j.h_out + (size_t)f * g.frame_bytes;
        const int k = clip_of[f];
        return j.h_clip[k] ? j.h_clip[k] + (size_t)(f - clip_first[k]) * g.frame_bytes : nullptr;
    };
    if (to_host && n >= 64) {  // worth overlapping: progress words in mapped host memory, a second stream for the tail work
        if (!c->copy_st) CK(cudaStreamCreateWithFlags(&c->copy_st, cudaStreamNonBlocking));
        if (c->dec_progress_cap < n_chains) {
            if (c->dec_progress) cudaFreeHost((void*)c->dec_progress);
            c->dec_progress = nullptr;
            c->dec_progress_cap = 0;
            void* hp = nullptr;
            CK(cudaHostAlloc(&hp, (size_t)(n_chains + 16) * sizeof(int), cudaHostAllocMapped));
            c->dec_progress = (volatile int*)hp;
            c->dec_progress_cap = n_chains + 16;
        }
        for (int k = 0; k < n_chains; k++) c->dec_progress[k] = chains[k].first;
        void* dp = nullptr;
        CK(cudaHostGetDevicePointer(&dp, (void*)c->dec_progress, 0));
        w.progress = (volatile int*)dp;
    }
    CK(cudaMemsetAsync(w.upd, 0, (size_t)n * g.nb, st));
    StageTimer tm(st);
    const bool v2 = version == 2;
#define LAUNCH_CHAIN(SMV, V2V)                                                                                      \
    do {                                                                                                            \
        CK(cudaFuncSetAttribute(k_dec_chain<SMV, V2V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
        k_dec_chain<SMV, V2V><<<n_chains, 32 * DEC_WARPS, smem, st>>>(w);                                           \
    } while (0)
    if (map_shared && !v2) LAUNCH_CHAIN(true, false);
    else if (!map_shared && !v2) LAUNCH_CHAIN(false, false);
    else if (map_shared) LAUNCH_CHAIN(true, true);
    else LAUNCH_CHAIN(false, true);
#undef LAUNCH_CHAIN
    tm.mark("chain");
    // ---- untouched blocks + (host output) download, range by range -----------------------------------------------
    auto finish_range = [&](int k, int a, int b, cudaStream_t s2) {
        DecWork r = w;
        r.chain = k; r.f_begin = a; r.f_end = b;
        k_dec_sources<<<(g.nb + 127) / 128, 128, 0, s2>>>(r);
        const long strips = (long)(b - a) * g.nby * ((g.nbx + 7) >> 3);
        k_dec_fill<<<(unsigned)((strips + 7) / 8), 256, 0, s2>>>(r);
        c->launches += 2;
        if (to_host) {
            uint8_t* h = host_of(a);
            if (h) download_frames(h, j.d_out + (size_t)a * g.frame_bytes, g.frame_bytes, pitch, rowb, g.Y, a, b, s2);
        }
    };
    if (w.progress) {
        // the chain kernel reports finished frames; finished ranges are completed and sent home while it keeps going
        const int step = 24;
        std::vector<int> next(n_chains);
        for (int k = 0; k < n_chains; k++) next[k] = chains[k].first;
        bool idle = false;  // the chain kernel has left the stream: whatever it reported last is final
        for (bool busy = true; busy;) {
            busy = false;
            for (int k = 0; k < n_chains; k++) {
                const int end = chains[k].first + chains[k].count;
                const int done = idle ? end : c->dec_progress[k];
                while (next[k] < end && (next[k] + step <= done || done >= end)) {
                    const int b = done >= end ? (next[k] + 4 * step < end ? next[k] + 4 * step : end) : next[k] + step;
                    finish_range(k, next[k], b, c->copy_st);
                    next[k] = b;
                }
                if (next[k] < end) busy = true;
            }
            if (busy) {
                if (cudaStreamQuery(st) != cudaErrorNotReady) idle = true;  // finished (or failed: reported by the sync below)
                else {
                    struct timespec ts = {0, 100000};
                    nanosleep(&ts, nullptr);
                }
            }
        }
        CK(cudaStreamSynchronize(c->copy_st));
    } else {
        CK(cudaStreamSynchronize(st));  // chains may have finished in any order; ranges of one chain go in frame order
        for (int k = 0; k < n_chains; k++) finish_range(k, chains[k].first, chains[k].first + chains[k].count, st);
    }
    c->launches += 1;
    tm.mark("sources+fill");
    tm.report("decode_batch");
    CK(cudaStreamSynchronize(st));
    if (!multi) CK(cudaMemcpyAsync(c->dec_prev.p, j.d_out + (size_t)(n - 1) * g.frame_bytes, g.frame_bytes, cudaMemcpyDeviceToDevice, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    // commit the decoder state: the frames have been decoded
    if (multi) {  // a multi-clip call leaves no stream open
        c->dec_created = false;
        c->dec_last_was_flat = false;
    } else {
        c->dec_created = created;
        c->dec_version = version;
        c->dec_last_was_flat = last_was_flat;
        memcpy(c->dec_last_flat_clr, last_flat_clr, 3);
        c->dec_cur_state = new_cur_state;
    }
    return 1;
}

// any number of frames: launches of at most DEC_MAX_FRAMES frames (the open chain carries over, as between calls)
static int decode_batch(scpr_codec* c, const uint8_t* stream, const uint32_t* sizes, const uint8_t* ftypes, int n, uint8_t* d_out,
                        int pitch, uint8_t* h_out = nullptr) {
    if (c->rgb16) {  // decode to RGB24 on the device, then put the 16-bit words together (screencap.cpp:1726-1734)
        if (pitch < c->g.X * 2) return SCPR_E_PARAM;
        CK(cudaSetDevice(c->device));
        uint8_t* const d16 = d_out;
        for (int f0 = 0; f0 < n; f0 += DEC_MAX_FRAMES) {
            const int m = n - f0 < DEC_MAX_FRAMES ? n - f0 : DEC_MAX_FRAMES;
            TRY(c->dec24.ensure((size_t)m * c->g.frame_bytes));
            const DecJob j = {stream, sizes + f0, ftypes + f0, m, (uint8_t*)c->dec24.p, c->g.pitch, nullptr, nullptr, nullptr, nullptr};
            const int r = decode_range(c, j);
            if (r != 1) return r;
            launch_pack16((const uint8_t*)c->dec24.p, d16 + (size_t)f0 * pitch * c->g.Y, m, c->g, pitch, c->m16, c->st, &c->launches);
            CK(cudaStreamSynchronize(c->st));
            for (int f = f0; f < f0 + m; f++) stream += sizes[f];
        }
        return 1;
    }
    const size_t frame_bytes = (size_t)pitch * c->g.Y;
    for (int f0 = 0; f0 < n; f0 += DEC_MAX_FRAMES) {
        const int m = n - f0 < DEC_MAX_FRAMES ? n - f0 : DEC_MAX_FRAMES;
        const DecJob j = {stream, sizes + f0, ftypes + f0, m, d_out + (size_t)f0 * frame_bytes, pitch, nullptr, nullptr,
                          h_out ? h_out + (size_t)f0 * frame_bytes : nullptr, nullptr};
        const int r = decode_range(c, j);
        if (r != 1) return r;
        for (int f = f0; f < f0 + m; f++) stream += sizes[f];
    }
    return 1;
}

// Independent clips, all their GOP chains in one launch per DEC_MAX_FRAMES frames.  d_base: device frames of all clips back to
// back; h_out[k]: host destination of clip k or nullptr.
static int decode_clips(scpr_codec* c, scpr_clip* clips, int n_clips, uint8_t* d_base, int pitch, bool to_host) {
    const size_t frame_bytes = (size_t)pitch * c->g.Y;
    int all_ok = 1;
    for (int k0 = 0; k0 < n_clips;) {
        // a group of clips of at most DEC_MAX_FRAMES frames
        int k1 = k0;
        long nf = 0;
        while (k1 < n_clips && (k1 == k0 || nf + clips[k1].n <= DEC_MAX_FRAMES)) nf += clips[k1++].n;
        size_t first_frame = 0;
        for (int k = 0; k < k0; k++) first_frame += (size_t)clips[k].n;
        if (k1 == k0 + 1 && clips[k0].n > DEC_MAX_FRAMES) {  // one very long clip: as a single stream, in several launches
            scpr_clip& q = clips[k0];
            c->dec_created = false;
            c->dec_last_was_flat = false;
            q.result = decode_batch(c, q.stream, q.sizes, q.ftypes, q.n, d_base + first_frame * frame_bytes, pitch, to_host ? q.frames : nullptr);
            c->dec_created = false;
        } else {
            std::vector<uint8_t> stream, ftypes, starts;
            std::vector<uint32_t> sizes;
            std::vector<uint8_t*> hdst;
            std::vector<int> results(k1 - k0, 1);
            for (int k = k0; k < k1; k++) {
                const scpr_clip& q = clips[k];
                size_t bytes = 0;
                for (int f = 0; f < q.n; f++) bytes += q.sizes[f];
                stream.insert(stream.end(), q.stream, q.stream + bytes);
                sizes.insert(sizes.end(), q.sizes, q.sizes + q.n);
                ftypes.insert(ftypes.end(), q.ftypes, q.ftypes + q.n);
                for (int f = 0; f < q.n; f++) starts.push_back(f == 0);
                hdst.push_back(to_host ? q.frames : nullptr);
            }
            const DecJob j = {stream.data(), sizes.data(), ftypes.data(), (int)nf, d_base + first_frame * frame_bytes, pitch, starts.data(),
                              to_host ? hdst.data() : nullptr, nullptr, results.data()};
            const int r = nf ? decode_range(c, j) : 1;
            for (int k = k0; k < k1; k++) clips[k].result = r == 1 ? results[k - k0] : r;
        }
        for (int k = k0; k < k1; k++)
            if (clips[k].result != 1 && all_ok == 1) all_ok = clips[k].result;
        k0 = k1;
    }
    return all_ok;
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
    // the download is part of decode_batch: finished frame ranges travel while the chains are still decoding; only the
    // pixel bytes of every row travel, the caller's row padding stays as it is
    const bool direct = !c->rgb16;
    const int r = decode_batch(c, stream, sizes, ftypes, n, (uint8_t*)c->dec_frames.p, pitch, direct ? frames : nullptr);
    if (r != 1) return r;
    if (!direct) {
        CK(download_frames(frames, (const uint8_t*)c->dec_frames.p, (size_t)pitch * c->g.Y, pitch, row_bytes(c), c->g.Y, 0, n, c->st));
        CK(cudaStreamSynchronize(c->st));
    }
    return 1;
}

int scpr_decompress_frame(scpr_codec* c, const uint8_t* src, int src_len, uint8_t* dst, int pitch, int ftype) {
    if (!c || !src || !dst || src_len <= 0) return SCPR_E_PARAM;
    const uint32_t size = (uint32_t)src_len;
    const uint8_t ft = ftype ? 1 : 0;
    return scpr_decompress_clip(c, src, &size, &ft, 1, dst, pitch);
}

static int check_clips(const scpr_codec* c, const scpr_clip* clips, int n_clips, bool need_frames) {
    if (!c || !clips || n_clips < 0) return SCPR_E_PARAM;
    if (c->rgb16) {
        set_error("multi-clip decoding of 16 bpp clients is not built");
        return SCPR_E_UNSUPPORTED;
    }
    for (int k = 0; k < n_clips; k++)
        if (clips[k].n < 0 || (clips[k].n && (!clips[k].stream || !clips[k].sizes || !clips[k].ftypes || (need_frames && !clips[k].frames)))) return SCPR_E_PARAM;
    return SCPR_OK;
}

int scpr_decompress_clips_dev(scpr_codec* c, scpr_clip* clips, int n_clips, uint8_t* d_frames, int pitch) {
    TRY(check_clips(c, clips, n_clips, false));
    if (!d_frames) return SCPR_E_PARAM;
    return decode_clips(c, clips, n_clips, d_frames, pitch, false);
}

int scpr_decompress_clips(scpr_codec* c, scpr_clip* clips, int n_clips, int pitch) {
    TRY(check_clips(c, clips, n_clips, true));
    CK(cudaSetDevice(c->device));
    size_t nf = 0;
    for (int k = 0; k < n_clips; k++) nf += (size_t)clips[k].n;
    if (nf == 0) return 1;
    TRY(c->dec_frames.ensure(nf * pitch * c->g.Y));
    return decode_clips(c, clips, n_clips, (uint8_t*)c->dec_frames.p, pitch, true);
}

}  // extern "C"
