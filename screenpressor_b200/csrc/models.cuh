// models.cuh -- device-side adaptive models of the v4 entropy layer, flat per-symbol layout.
//
// Replaces the reference's Context / Cx1..Cx7 / SmallContext / FixedSizeRansCtx
// (ans_contexts.h:73-1132, ans_contexts.cpp:3-84).  The reference keeps Cx6 in a Robin-Hood hash
// (encoder) or a frequency-sorted list (decoder); every interval it produces is a function of the
// symbol values alone, so here kinds 6 and 7 share one flat table indexed by
// symbol: cnt[256] (0 = symbol not met yet, kind 6 only), and the frozen interval table
// freq[256] / cum[256] that only changes at a rescale.
#pragma once
#include "scpr_dev.cuh"

namespace scpr {

struct FixedState {  // FixedSizeRansCtx<N>, ans_contexts.h:1053-1132
    int cntsum;
    int nsym;
    uint16_t cnt[512], freq[512], cum[512];
};

// 128-byte head = everything kinds 0..5 touch (one cache line), then the flat tables of kinds 6/7.
struct alignas(128) ColorState {
    uint8_t kind;    // 0 empty, 1..3 "every symbol met once" sets, 4/5 SmallContext, 6 Cx6, 7 Cx7
    uint8_t fshift;  // kind 6
    uint8_t maxpos;  // kinds 4/5
    uint8_t pad;
    uint16_t d;      // distinct symbols met
    uint16_t pad2;
    int cntsum;      // kind 5: cached totFr (the decoder also caches it for kind 4); kinds 6/7: counter sum
    uint32_t pad3;
    uint8_t ssym[16];     // kinds 4/5: sorted symbols ...
    uint16_t sfreq[16];   // ... and their frequencies
    union {
        uint32_t seen[8];   // kinds 1..3: bitmap of symbols met
        uint32_t sent[16];  // kinds 4/5, decoder only: entry k packed as sym | freq << 8 | start << 20 with
                            // start = sum of the frequencies before k + (ssym[k] - k), see decode.cu
    };
    uint16_t cnt[256], freq[256], cum[256];  // kinds 6/7 (16-byte aligned rows for 128-bit table scans)
};
static_assert(sizeof(ColorState) == 1664, "ColorState layout");

struct ModelState {
    ColorState color[NUM_COLOR_CX];
    FixedState fx[NUM_FIXED_CX];
    uint8_t kmap[NUM_COLOR_CX];  // decoder: kind of every colour context (kept in shared memory while a chain runs)
};

// ---- kinds 1..3 ------------------------------------------------------------------------------
__device__ __forceinline__ bool seen_has(const ColorState& x, int c) { return (x.seen[c >> 5] >> (c & 31)) & 1; }
__device__ __forceinline__ void seen_add(ColorState& x, int c) { x.seen[c >> 5] |= 1u << (c & 31); }

// ---- SmallContext (kinds 4/5), ans_contexts.h:154-290 -------------------------------------------
__device__ inline void small_from_set(ColorState& x, int c) {  // create from Cx1, :161-172
    int d = 0;
    for (int s = 0; s < 256; s++)
        if (seen_has(x, s)) {
            x.ssym[d] = (uint8_t)s;
            if (s == c) {
                x.sfreq[d] = 100;
                x.maxpos = (uint8_t)d;
            } else
                x.sfreq[d] = 50;
            d++;
        }
    for (int i = d; i < 16; i++) x.sfreq[i] = 0;
    x.d = (uint16_t)d;
}
__device__ inline int small_calcsum(const ColorState& x) {  // :303, :334-338
    int t = 256 - x.d;
    for (int i = 0; i < x.d; i++) t += x.sfreq[i];
    return t;
}
__device__ inline void small_rescale(ColorState& x, int& totFr) {  // :186-193
    int s = 256 - x.d;
    for (int i = 0; i < x.d; i++) {
        x.sfreq[i] = (uint16_t)(x.sfreq[i] - (x.sfreq[i] >> 1));
        s += x.sfreq[i];
    }
    totFr = s & 0xFFFF;
}
__device__ inline bool small_add(ColorState& x, int S, int pos, int c, int& totFr) {  // :174-184
    if (x.d == S) return false;
    for (int i = x.d - 1; i >= pos; i--) {
        x.ssym[i + 1] = x.ssym[i];
        x.sfreq[i + 1] = x.sfreq[i];
    }
    x.ssym[pos] = (uint8_t)c;
    x.sfreq[pos] = 50;
    x.d++;
    if (x.maxpos >= pos) x.maxpos++;
    totFr = (totFr + 50) & 0xFFFF;
    if (totFr + 50 > PROB_SCALE) small_rescale(x, totFr);
    return true;
}
// :195-236.  false = new symbol and the table is full (the caller promotes); iv is valid either way.
__device__ inline bool small_encode(ColorState& x, int S, int c, uint32_t& iv, int& totFr) {
    int shift = 0, tot = totFr;
    while (tot <= PROB_SCALE / 2) {
        tot <<= 1;
        shift++;
    }
    const int bonus = (PROB_SCALE - tot) >> shift;  // unused code space goes to the most probable symbol
    const int d = x.d, maxpos = x.maxpos;
    int cumFr = 0, lastSymb = 0, pos = 0;
    while (pos < d) {
        const int s = x.ssym[pos];
        const int fr = (x.sfreq[pos] + (pos == maxpos ? bonus : 0)) & 0xFFFF;
        if (s == c) {
            cumFr += c - lastSymb;
            iv = make_iv((uint32_t)(fr << shift) & 0xFFFF, (uint32_t)(cumFr << shift) & 0xFFFF);
            x.sfreq[pos] = (uint16_t)(x.sfreq[pos] + 50);
            totFr = (totFr + 50) & 0xFFFF;
            if (pos != maxpos && x.sfreq[pos] > x.sfreq[maxpos]) x.maxpos = (uint8_t)pos;
            if (totFr + 50 > PROB_SCALE) small_rescale(x, totFr);
            return true;
        }
        if (c < s) break;
        cumFr += s - lastSymb + fr;
        lastSymb = s + 1;
        pos++;
    }
    cumFr += c - lastSymb;
    iv = make_iv((uint32_t)(1 << shift), (uint32_t)(cumFr << shift) & 0xFFFF);
    return small_add(x, S, pos, c, totFr);
}
// decode-side search, :238-283
__device__ inline int small_find(const ColorState& x, int totFr, int v) {
    int shift = 0, tot = totFr;
    while (tot <= PROB_SCALE / 2) {
        tot <<= 1;
        shift++;
    }
    v >>= shift;
    const int bonus = (PROB_SCALE - tot) >> shift;
    int cumFr = 0, lastSymb = 0;
    for (int pos = 0; pos < x.d; pos++) {
        const int s = x.ssym[pos];
        const int startFr = cumFr + s - lastSymb;
        if (v < startFr) return v - cumFr + lastSymb;
        const int fr = (x.sfreq[pos] + (pos == x.maxpos ? bonus : 0)) & 0xFFFF;
        if (startFr + fr > v) return s;
        cumFr += s - lastSymb + fr;
        lastSymb = s + 1;
    }
    return lastSymb + v - cumFr;
}

// ---- Cx6 (kind 6), ans_contexts.h:377-829 ---------------------------------------------------------
__device__ inline int c6_calcsum(const ColorState& x) {  // :549-555
    const int shft = x.fshift > 0 ? x.fshift - 1 : 0;
    int sum = (256 - x.d) << shft;
    for (int s = 0; s < 256; s++) sum += x.cnt[s];
    return sum;
}
// met symbols (sorted) get frs<<shift, every other symbol an implicit 1<<shift slot (:454-531, :387-415)
__device__ inline void c6_build(ColorState& x, const uint8_t* syms, const uint16_t* frs, int d, int totFr) {
    int shift = 0, tot = totFr;
    while (tot <= PROB_SCALE / 2) {
        tot <<= 1;
        shift++;
    }
    int cumFr = 0, k = 0;
    for (int s = 0; s < 256; s++) {
        int fr;
        if (k < d && syms[k] == s) {
            fr = (int)frs[k] << shift;
            x.cnt[s] = (uint16_t)(fr - (fr >> 1));
            k++;
        } else {
            fr = 1 << shift;
            x.cnt[s] = 0;
        }
        x.freq[s] = (uint16_t)fr;
        x.cum[s] = (uint16_t)cumFr;
        cumFr += fr;
    }
    x.kind = 6;
    x.fshift = (uint8_t)shift;
    x.d = (uint16_t)d;
}
__device__ inline void c6_rescale(ColorState& x) {  // :742-796
    const int sh = x.fshift > 0 ? x.fshift - 1 : 0;
    const int c0 = 1 << sh;
    int cumFr = 0;
    for (int s = 0; s < 256; s++) {
        const int c = x.cnt[s] ? x.cnt[s] : c0;
        x.freq[s] = (uint16_t)c;
        x.cum[s] = (uint16_t)cumFr;
        cumFr += c;
    }
    if (x.fshift > 0) x.fshift--;
    const int shft = x.fshift > 0 ? x.fshift - 1 : 0;
    int sum = (256 - x.d) << shft;
    for (int s = 0; s < 256; s++)
        if (x.cnt[s]) {
            x.cnt[s] = (uint16_t)(x.cnt[s] - (x.cnt[s] >> 1));
            sum += x.cnt[s];
        }
    x.cntsum = sum & 0xFFFF;
}
__device__ inline void c6_incr(ColorState& x, int c) {  // :686-691
    const int step = 25 << x.fshift;
    x.cnt[c] = (uint16_t)(x.cnt[c] + step);
    x.cntsum = (x.cntsum + step) & 0xFFFF;
    if (x.cntsum + step > PROB_SCALE) c6_rescale(x);
}

// ---- Cx7 (kind 7), ans_contexts.h:847-998; also the fixed tables ----------------------------------
__device__ inline void table_incr(uint16_t* cnt, uint16_t* freq, uint16_t* cum, int nsym, int& cntsum, int c) {  // :959-981, :1070-1091
    cnt[c] = (uint16_t)(cnt[c] + 16);
    cntsum += 16;
    if (cntsum + 16 > PROB_SCALE) {
        int cf = 0, sum = 0;
        for (int j = 0; j < nsym; j++) {
            const int fr = cnt[j];
            cum[j] = (uint16_t)cf;
            freq[j] = (uint16_t)fr;
            cf += fr;
            cnt[j] = (uint16_t)(cnt[j] - (fr >> 1));
            sum += cnt[j];
        }
        cntsum = sum;
    }
}
__device__ inline void c7_from_c6(ColorState& x) {  // :868-915
    const int funmet = 1 << x.fshift, cu = funmet - (funmet >> 1);
    for (int s = 0; s < 256; s++)
        if (!x.cnt[s]) x.cnt[s] = (uint16_t)cu;
    x.kind = 7;
}
__device__ inline void c7_from_set(ColorState& x, int c) {  // :917-951
    const int d = x.d;
    const int f0 = (PROB_SCALE - (256 - d)) / (d + 1), c0 = f0 - (f0 >> 1);
    for (int s = 0; s < 256; s++) {
        const bool m = seen_has(x, s);
        x.freq[s] = (uint16_t)(m ? f0 : 1);
        x.cnt[s] = (uint16_t)(m ? c0 : 1);
    }
    x.freq[c] = (uint16_t)(x.freq[c] + f0);
    x.cnt[c] = (uint16_t)(x.cnt[c] + 16);
    int sum = 0, cf = 0;
    for (int s = 0; s < 256; s++) {
        sum += x.cnt[s];
        x.cum[s] = (uint16_t)cf;
        cf += x.freq[s];
    }
    x.cntsum = sum;
    x.kind = 7;
}

// ---- Context::update for kinds 0..3 (the byte is stored raw), ans_contexts.cpp:3-31, 52-59 --------
static __device__ __noinline__ void cc_update_raw(ColorState& x, int c, int f0) {
    switch (x.kind) {
    case 0:
        for (int i = 0; i < 8; i++) x.seen[i] = 0;
        seen_add(x, c);
        x.d = 1;
        x.kind = 1;
        break;
    case 1:
        if (seen_has(x, c)) {
            const bool small4 = x.d <= 4;
            small_from_set(x, c);
            if (small4)
                x.kind = 4;
            else {
                x.kind = 5;
                x.cntsum = small_calcsum(x);
            }
        } else {
            seen_add(x, c);
            x.d++;
            if (x.d > 14) x.kind = 2;
        }
        break;
    case 2:
        if (seen_has(x, c)) {  // Cx2 -> Cx6, create23 (:491-533)
            uint8_t syms[64];
            uint16_t frs[64];
            int d = 0;
            for (int s = 0; s < 256; s++)
                if (seen_has(x, s)) {
                    syms[d] = (uint8_t)s;
                    frs[d] = (uint16_t)((s == c) ? 2 * f0 : f0);
                    d++;
                }
            c6_build(x, syms, frs, d, 256 - d + d * f0 + f0);
            x.cntsum = c6_calcsum(x) & 0xFFFF;
        } else {
            seen_add(x, c);
            x.d++;
            if (x.d > 64) x.kind = 3;
        }
        break;
    case 3:
        if (seen_has(x, c))
            c7_from_set(x, c);
        else {
            seen_add(x, c);
            x.d++;
        }
        break;
    }
}

// ---- Context::encode for kinds >= 4, ans_contexts.cpp:34-50 -----------------------------------------
static __device__ __noinline__ uint32_t cc_encode_counted(ColorState& x, int c) {
    uint32_t iv = 0;
    switch (x.kind) {
    case 4: {
        int totFr = small_calcsum(x) & 0xFFFF;
        if (!small_encode(x, 4, c, iv, totFr)) {  // Cx5::create(Cx4&, c), :350-369; maxpos restarts at 0
            int i = x.d, sum = 0;
            while (i > 0 && x.ssym[i - 1] > c) {
                x.ssym[i] = x.ssym[i - 1];
                x.sfreq[i] = x.sfreq[i - 1];
                i--;
            }
            x.ssym[i] = (uint8_t)c;
            x.sfreq[i] = 50;
            x.d++;
            x.maxpos = 0;
            for (int k = 0; k < x.d; k++) sum += x.sfreq[k];
            if (sum > PROB_SCALE) {
                int t = 0;
                small_rescale(x, t);
            }
            x.cntsum = small_calcsum(x);
            x.kind = 5;
        }
        break;
    }
    case 5:
        if (!small_encode(x, 16, c, iv, x.cntsum)) {  // Cx6::create(Cx5&, c), :454-489
            uint8_t syms[16];
            uint16_t frs[16];
            const int d = x.d;
            for (int k = 0; k < d; k++) {
                syms[k] = x.ssym[k];
                frs[k] = x.sfreq[k];
            }
            c6_build(x, syms, frs, d, small_calcsum(x));
            const int fr = 1 << x.fshift;
            x.cnt[c] = (uint16_t)(fr - (fr >> 1) + (25 << x.fshift));
            x.d++;
            x.cntsum = c6_calcsum(x) & 0xFFFF;
        }
        break;
    case 6:
        iv = make_iv(x.freq[c], x.cum[c]);
        if (x.cnt[c])
            c6_incr(x, c);
        else if (x.d >= 40)  // MaxD6: Cx6 -> Cx7, the promoting symbol is not counted (:631, :868-915)
            c7_from_c6(x);
        else {  // placeSymbol, :621-638
            const int fr = 1 << x.fshift;
            x.cnt[c] = (uint16_t)(fr - (fr >> 1));
            x.d++;
            c6_incr(x, c);
        }
        break;
    case 7:
        iv = make_iv(x.freq[c], x.cum[c]);
        table_incr(x.cnt, x.freq, x.cum, 256, x.cntsum, c);
        break;
    }
    return iv;
}

__device__ inline int table_find(const uint16_t* cum, int nsym, int v) {  // :983-997, :1093-1112
    for (int j = 0; j < nsym - 1; j++)
        if (cum[j + 1] > v) return j;
    return nsym - 1;
}
__device__ inline int cc_find(const ColorState& x, int v) {
    if (x.kind == 4) return small_find(x, small_calcsum(x) & 0xFFFF, v);
    if (x.kind == 5) return small_find(x, x.cntsum, v);
    return table_find(x.cum, 256, v);
}

__device__ inline void fixed_renew(FixedState& f, int nsym) {  // :1114-1131
    const int fr = PROB_SCALE / nsym, c0 = fr - (fr >> 1);
    int cf = 0;
    f.nsym = nsym;
    f.cntsum = c0 * nsym;
    for (int i = 0; i < nsym; i++) {
        f.freq[i] = (uint16_t)fr;
        f.cum[i] = (uint16_t)cf;
        f.cnt[i] = (uint16_t)c0;
        cf += fr;
    }
}

}  // namespace scpr
