/* oracle/scpr_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * Plain-C CPU restatement of ScreenPressor's v4 per-frame encode/decode path, written from the
 * behaviour of the reference (file:line citations are relative to /root/reference).  It is the
 * checker the CUDA path is compared against; it is never linked into, imported by or called from
 * the product.
 *
 * PARITY PIN: the reference ships no tests or golden vectors, so this
 * restatement is pinned against outputs of the reference itself: tests/test_oracle_vs_ref.py
 * byte-compares it with oracle/_ref/libscpr_ref.so (the unmodified reference core, 1 thread =
 * canonical bitstream) and tests/golden/ holds per-frame digests generated from that library.
 *
 * The restatement is stage-separated, which is how the CUDA path is organised:
 *   A. frame -> ordered list of (context id, symbol) events        (screencap.cpp:319-403, 876-1271)
 *   B. events -> 12-bit intervals through the adaptive models      (ans_contexts.h, ans_contexts.cpp)
 *   C. intervals -> independent rANS blocks of 131072              (ransmt.h:116-134, rans_byte.h:47-102)
 * Models are held in a flat per-symbol layout instead of the reference's hash / sorted tables;
 * interval arithmetic in the reference is a function of symbol values only.
 */
#include "scpr_oracle.h"

#include <stdlib.h>
#include <string.h>

#define PROB_BITS 12
#define PROB_SCALE 4096
#define RANS_L (1u << 23)
#define RANS_BLOCK 131072 /* ransmt.h:38 */

/* ------------------------------------------------------------------------------------------
 * Stage B: adaptive models
 * ---------------------------------------------------------------------------------------- */

/* FixedSizeRansCtx<N> (ans_contexts.h:1053-1132); also the storage of colour kinds 6 and 7. */
typedef struct {
    int nsym;
    int cntsum;
    uint16_t cnt[512], freq[512], cum[512];
} FixedCtx;

static void fx_renew(FixedCtx* f, int nsym) { /* ans_contexts.h:1114-1131 */
    int fr = PROB_SCALE / nsym, c0 = fr - (fr >> 1), cf = 0;
    f->nsym = nsym;
    f->cntsum = c0 * nsym;
    for (int i = 0; i < nsym; i++) {
        f->freq[i] = (uint16_t)fr;
        f->cum[i] = (uint16_t)cf;
        f->cnt[i] = (uint16_t)c0;
        cf += fr;
    }
}

/* shared by Fixed (step 16), Cx7 (step 16): ans_contexts.h:1070-1091, 959-981 */
static void table_incr(uint16_t* cnt, uint16_t* freq, uint16_t* cum, int nsym, int* cntsum, int c, int step) {
    cnt[c] = (uint16_t)(cnt[c] + step);
    *cntsum += step;
    if (*cntsum + step > PROB_SCALE) {
        int cf = 0, sum = 0;
        for (int j = 0; j < nsym; j++) {
            int fr = cnt[j];
            cum[j] = (uint16_t)cf;
            freq[j] = (uint16_t)fr;
            cf += fr;
            cnt[j] = (uint16_t)(cnt[j] - (fr >> 1));
            sum += cnt[j];
        }
        *cntsum = sum;
    }
}

static orc_freq fx_encode(FixedCtx* f, int c) { /* ans_contexts.h:1063-1068 */
    orc_freq r;
    r.freq = f->freq[c];
    r.cum = f->cum[c];
    table_incr(f->cnt, f->freq, f->cum, f->nsym, &f->cntsum, c, 16);
    return r;
}

static int table_find(const uint16_t* cum, int nsym, int v) {
    /* symbol whose interval contains v (ans_contexts.h:1093-1112: any start <= answer works) */
    for (int j = 0; j < nsym - 1; j++)
        if (cum[j + 1] > v) return j;
    return nsym - 1;
}

/* Colour context: Context / Cx1..Cx7 (ans_contexts.h:73-1051, ans_contexts.cpp:3-84). */
typedef struct {
    uint8_t kind;   /* 0 empty, 1..3 "seen once" sets, 4/5 SmallContext, 6 Cx6, 7 Cx7 */
    uint8_t fshift; /* kind 6 */
    uint8_t maxpos; /* kinds 4/5 */
    uint16_t d;     /* distinct symbols */
    int cntsum;     /* kind 5: cached totFr; kinds 6/7: counter sum */
    uint8_t seen[32];                 /* kinds 1..3: bitmap of symbols met (each exactly once) */
    uint8_t ssym[16];                 /* kinds 4/5: sorted symbols */
    uint16_t sfreq[16];               /*            and their frequencies */
    uint16_t cnt[256], freq[256], cum[256]; /* kinds 6/7: flat per-symbol tables; kind 6: cnt==0 <=> unmet */
} ColorCtx;

static int seen_has(const ColorCtx* x, int c) { return (x->seen[c >> 3] >> (c & 7)) & 1; }
static void seen_add(ColorCtx* x, int c) { x->seen[c >> 3] |= (uint8_t)(1 << (c & 7)); }

/* SmallContext::create from Cx1 (ans_contexts.h:161-172): sorted symbols, f0 each, 2*f0 for c */
static void small_from_set(ColorCtx* x, int c) {
    int d = 0;
    for (int s = 0; s < 256; s++)
        if (seen_has(x, s)) {
            x->ssym[d] = (uint8_t)s;
            if (s == c) {
                x->sfreq[d] = 100;
                x->maxpos = (uint8_t)d;
            } else
                x->sfreq[d] = 50;
            d++;
        }
    for (int i = d; i < 16; i++) x->sfreq[i] = 0;
    x->d = (uint16_t)d;
}

static int small_calcsum(const ColorCtx* x) { /* Cx5::calcSum ans_contexts.h:334-338, Cx4 :303 */
    int t = 256 - x->d;
    for (int i = 0; i < x->d; i++) t += x->sfreq[i];
    return t;
}

static void small_rescale(ColorCtx* x, int* totFr) { /* ans_contexts.h:186-193 */
    int s = 256 - x->d;
    for (int i = 0; i < x->d; i++) {
        x->sfreq[i] = (uint16_t)(x->sfreq[i] - (x->sfreq[i] >> 1));
        s += x->sfreq[i];
    }
    *totFr = s & 0xFFFF;
}

static int small_add(ColorCtx* x, int S, int pos, int c, int* totFr) { /* ans_contexts.h:174-184 */
    if (x->d == S) return 0;
    for (int i = x->d - 1; i >= pos; i--) {
        x->ssym[i + 1] = x->ssym[i];
        x->sfreq[i + 1] = x->sfreq[i];
    }
    x->ssym[pos] = (uint8_t)c;
    x->sfreq[pos] = 50;
    x->d++;
    if (x->maxpos >= pos) x->maxpos++;
    *totFr = (*totFr + 50) & 0xFFFF;
    if (*totFr + 50 > PROB_SCALE) small_rescale(x, totFr);
    return 1;
}

/* SmallContext::encode (ans_contexts.h:195-236).  Returns 0 when the symbol is new and the table
 * is full (caller promotes); the interval is valid either way. */
static int small_encode(ColorCtx* x, int S, int c, orc_freq* iv, int* totFr) {
    int shift = 0, tot = *totFr;
    while (tot <= PROB_SCALE / 2) {
        tot <<= 1;
        shift++;
    }
    const int bonus = (PROB_SCALE - tot) >> shift;
    const int d = x->d, maxpos = x->maxpos;
    int cumFr = 0, lastSymb = 0, pos = 0;
    while (pos < d) {
        int s = x->ssym[pos];
        int fr = x->sfreq[pos] + (pos == maxpos ? bonus : 0);
        if (s == c) {
            cumFr += c - lastSymb;
            iv->cum = (uint16_t)(cumFr << shift);
            iv->freq = (uint16_t)((fr & 0xFFFF) << shift);
            x->sfreq[pos] = (uint16_t)(x->sfreq[pos] + 50);
            *totFr = (*totFr + 50) & 0xFFFF;
            if (pos != maxpos && x->sfreq[pos] > x->sfreq[maxpos]) x->maxpos = (uint8_t)pos;
            if (*totFr + 50 > PROB_SCALE) small_rescale(x, totFr);
            return 1;
        }
        if (c < s) break;
        cumFr += s - lastSymb + (fr & 0xFFFF);
        lastSymb = s + 1;
        pos++;
    }
    cumFr += c - lastSymb;
    iv->cum = (uint16_t)(cumFr << shift);
    iv->freq = (uint16_t)(1 << shift);
    return small_add(x, S, pos, c, totFr);
}

/* SmallContext::decode's search (ans_contexts.h:238-283): the symbol whose interval holds v */
static int small_find(const ColorCtx* x, int totFr, int v) {
    int shift = 0, tot = totFr;
    while (tot <= PROB_SCALE / 2) {
        tot <<= 1;
        shift++;
    }
    v >>= shift;
    const int bonus = (PROB_SCALE - tot) >> shift;
    int cumFr = 0, lastSymb = 0;
    for (int pos = 0; pos < x->d; pos++) {
        int s = x->ssym[pos];
        int startFr = cumFr + s - lastSymb;
        if (v < startFr) return v - cumFr + lastSymb;
        int fr = (x->sfreq[pos] + (pos == x->maxpos ? bonus : 0)) & 0xFFFF;
        if (startFr + fr > v) return s;
        cumFr += s - lastSymb + fr;
        lastSymb = s + 1;
    }
    return lastSymb + v - cumFr;
}

/* Cx6::calcSum (ans_contexts.h:549-555) on the flat layout */
static int c6_calcsum(const ColorCtx* x) {
    int shft = x->fshift > 0 ? x->fshift - 1 : 0;
    int sum = (256 - x->d) << shft;
    for (int s = 0; s < 256; s++) sum += x->cnt[s];
    return sum;
}

/* Build kind 6 from a sorted (symbol, freq) list: met symbols get freq<<shift, every other symbol
 * an implicit 1<<shift slot (Cx6::create ans_contexts.h:454-489, create23 :491-531, add :387-415). */
static void c6_build(ColorCtx* x, const uint8_t* syms, const int* frs, int d, int totFr) {
    int shift = 0, tot = totFr;
    while (tot <= PROB_SCALE / 2) {
        tot <<= 1;
        shift++;
    }
    memset(x->cnt, 0, sizeof(x->cnt));
    int cumFr = 0, k = 0;
    for (int s = 0; s < 256; s++) {
        int fr;
        if (k < d && syms[k] == s) {
            fr = frs[k] << shift;
            x->cnt[s] = (uint16_t)(fr - (fr >> 1));
            k++;
        } else
            fr = 1 << shift;
        x->freq[s] = (uint16_t)fr;
        x->cum[s] = (uint16_t)cumFr;
        cumFr += fr;
    }
    x->kind = 6;
    x->fshift = (uint8_t)shift;
    x->d = (uint16_t)d;
}

static void c6_rescale(ColorCtx* x) { /* Cx6::rescale ans_contexts.h:742-796 */
    int sh = x->fshift > 0 ? x->fshift - 1 : 0;
    int c0 = 1 << sh, cumFr = 0;
    for (int s = 0; s < 256; s++) {
        int c = x->cnt[s] ? x->cnt[s] : c0;
        x->freq[s] = (uint16_t)c;
        x->cum[s] = (uint16_t)cumFr;
        cumFr += c;
    }
    if (x->fshift > 0) x->fshift--;
    int shft = x->fshift > 0 ? x->fshift - 1 : 0;
    int sum = (256 - x->d) << shft;
    for (int s = 0; s < 256; s++)
        if (x->cnt[s]) {
            x->cnt[s] = (uint16_t)(x->cnt[s] - (x->cnt[s] >> 1));
            sum += x->cnt[s];
        }
    x->cntsum = sum & 0xFFFF;
}

static void c6_incr(ColorCtx* x, int c) { /* Cx6::incrCnt ans_contexts.h:686-691 */
    int step = 25 << x->fshift;
    x->cnt[c] = (uint16_t)(x->cnt[c] + step);
    x->cntsum = (x->cntsum + step) & 0xFFFF;
    if (x->cntsum + step > PROB_SCALE) c6_rescale(x);
}

/* Cx7::create from Cx6 (ans_contexts.h:868-915): tables carry over, unmet symbols get real counters;
 * the promoting symbol is not counted. */
static void c7_from_c6(ColorCtx* x) {
    int funmet = 1 << x->fshift, cu = funmet - (funmet >> 1);
    for (int s = 0; s < 256; s++)
        if (!x->cnt[s]) x->cnt[s] = (uint16_t)cu;
    x->kind = 7;
}

/* Cx7::create from Cx3 (ans_contexts.h:917-951) */
static void c7_from_set(ColorCtx* x, int c) {
    int d = x->d;
    int f0 = (PROB_SCALE - (256 - d)) / (d + 1), c0 = f0 - (f0 >> 1);
    for (int s = 0; s < 256; s++) {
        if (seen_has(x, s)) {
            x->freq[s] = (uint16_t)f0;
            x->cnt[s] = (uint16_t)c0;
        } else {
            x->freq[s] = 1;
            x->cnt[s] = 1;
        }
    }
    x->freq[c] = (uint16_t)(x->freq[c] + f0);
    x->cnt[c] = (uint16_t)(x->cnt[c] + 16);
    int sum = 0, cf = 0;
    for (int s = 0; s < 256; s++) {
        sum += x->cnt[s];
        x->cum[s] = (uint16_t)cf;
        cf += x->freq[s];
    }
    x->cntsum = sum;
    x->kind = 7;
}

/* kinds 0..3 (Context::update, updateC1/2/3: ans_contexts.cpp:3-31, 52-59); the byte goes out raw */
static void cc_update_raw(ColorCtx* x, int c, int f0) {
    switch (x->kind) {
    case 0:
        memset(x->seen, 0, sizeof(x->seen));
        seen_add(x, c);
        x->d = 1;
        x->kind = 1;
        break;
    case 1:
        if (seen_has(x, c)) { /* second sighting: start counting */
            if (x->d <= 4) {
                small_from_set(x, c);
                x->kind = 4;
            } else {
                small_from_set(x, c);
                x->kind = 5;
                x->cntsum = small_calcsum(x);
            }
        } else {
            seen_add(x, c);
            x->d++;
            if (x->d > 14) x->kind = 2; /* 15th distinct symbol: Cx1 -> Cx2 */
        }
        break;
    case 2:
        if (seen_has(x, c)) { /* Cx2 -> Cx6 via create23 (ans_contexts.h:491-533) */
            uint8_t syms[64];
            int frs[64], d = 0;
            for (int s = 0; s < 256; s++)
                if (seen_has(x, s)) {
                    syms[d] = (uint8_t)s;
                    frs[d] = (s == c) ? 2 * f0 : f0;
                    d++;
                }
            c6_build(x, syms, frs, d, 256 - d + d * f0 + f0);
            x->cntsum = c6_calcsum(x) & 0xFFFF;
        } else {
            seen_add(x, c);
            x->d++;
            if (x->d > 64) x->kind = 3; /* 65th distinct symbol: Cx2 -> Cx3 */
        }
        break;
    case 3:
        if (seen_has(x, c))
            c7_from_set(x, c);
        else {
            seen_add(x, c);
            x->d++;
        }
        break;
    }
}

/* Context::encode for kinds >= 4 (ans_contexts.cpp:34-50).  Interval + state update. */
static void cc_encode_counted(ColorCtx* x, int c, orc_freq* iv) {
    switch (x->kind) {
    case 4: {
        int totFr = small_calcsum(x) & 0xFFFF; /* Cx4 recomputes per call, ans_contexts.h:303 */
        if (!small_encode(x, 4, c, iv, &totFr)) {
            /* Cx5::create(Cx4&, c) ans_contexts.h:350-369: merge c in with f0; maxpos restarts at 0 */
            int i = x->d, sum = 0;
            while (i > 0 && x->ssym[i - 1] > c) {
                x->ssym[i] = x->ssym[i - 1];
                x->sfreq[i] = x->sfreq[i - 1];
                i--;
            }
            x->ssym[i] = (uint8_t)c;
            x->sfreq[i] = 50;
            x->d++;
            x->maxpos = 0;
            for (int k = 0; k < x->d; k++) sum += x->sfreq[k];
            if (sum > PROB_SCALE) {
                int t = 0;
                small_rescale(x, &t);
            }
            x->cntsum = small_calcsum(x);
            x->kind = 5;
        }
        break;
    }
    case 5:
        if (!small_encode(x, 16, c, iv, &x->cntsum)) {
            /* Cx6::create(Cx5&, c) ans_contexts.h:454-489 */
            uint8_t syms[16];
            int frs[16], d = x->d;
            for (int k = 0; k < d; k++) {
                syms[k] = x->ssym[k];
                frs[k] = x->sfreq[k];
            }
            c6_build(x, syms, frs, d, small_calcsum(x));
            /* add(c, unmet interval) then incrCnt; the sum is recomputed afterwards (:485-488) */
            int fr = 1 << x->fshift;
            x->cnt[c] = (uint16_t)(fr - (fr >> 1) + (25 << x->fshift));
            x->d++;
            x->cntsum = c6_calcsum(x) & 0xFFFF;
        }
        break;
    case 6:
        iv->freq = x->freq[c];
        iv->cum = x->cum[c];
        if (x->cnt[c])
            c6_incr(x, c);
        else if (x->d >= 40) /* MaxD6 reached: Cx6 -> Cx7 (ans_contexts.h:631, 670) */
            c7_from_c6(x);
        else { /* placeSymbol ans_contexts.h:621-638 */
            int fr = 1 << x->fshift;
            x->cnt[c] = (uint16_t)(fr - (fr >> 1));
            x->d++;
            c6_incr(x, c);
        }
        break;
    case 7:
        iv->freq = x->freq[c];
        iv->cum = x->cum[c];
        table_incr(x->cnt, x->freq, x->cum, 256, &x->cntsum, c, 16);
        break;
    }
}

static int cc_find(const ColorCtx* x, int v) { /* decode-side symbol search, kinds >= 4 */
    if (x->kind == 4) return small_find(x, small_calcsum(x) & 0xFFFF, v);
    if (x->kind == 5) return small_find(x, x->cntsum, v);
    return table_find(x->cum, 256, v);
}

/* ------------------------------------------------------------------------------------------
 * model set = everything RenewI resets (screencap.cpp:178-198, screencap.h:436-443)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    ColorCtx* color; /* [3*4096] */
    FixedCtx fx[ORC_CX_BOOL - ORC_CX_NTAB];
    int f0;
} Models;

static int fx_nsym(int id) {
    if (id < ORC_CX_BT) return 256; /* ntab[6], ntab2, xxtab */
    if (id == ORC_CX_BT) return 5;
    if (id < ORC_CX_MV) return 16;
    if (id < ORC_CX_PTYPE) return 512;
    return 6;
}

static void models_renew(Models* m) {
    for (int i = 0; i < 3 * 4096; i++) m->color[i].kind = 0;
    for (int id = ORC_CX_NTAB; id < ORC_CX_BOOL; id++) fx_renew(&m->fx[id - ORC_CX_NTAB], fx_nsym(id));
}

static void models_init(Models* m, int f0) {
    m->color = (ColorCtx*)calloc(3 * 4096, sizeof(ColorCtx));
    m->f0 = f0;
    models_renew(m);
}

/* coverage: how many times each colour-model promotion fired (tests assert the fuzz reaches all) */
static unsigned long g_trans[8][8];
unsigned long orc_transition_count(int from, int to) { return g_trans[from & 7][to & 7]; }

/* one event through its model (UseANS::encodeC / encodeF / encodeBool, screencap.h:311-317, 339-344, 407-410) */
static orc_freq models_encode(Models* m, uint32_t ev) {
    int id = (int)(ev >> 16), c = (int)(ev & 0xFFFF);
    orc_freq iv;
    if (id < ORC_CX_NTAB) {
        ColorCtx* x = &m->color[id];
        int k0 = x->kind;
        if (x->kind < 4) {
            cc_update_raw(x, c, m->f0);
            iv.freq = 0;
            iv.cum = (uint16_t)c;
        } else
            cc_encode_counted(x, c, &iv);
        if (x->kind != k0) g_trans[k0][x->kind]++;
    } else if (id < ORC_CX_BOOL)
        iv = fx_encode(&m->fx[id - ORC_CX_NTAB], c);
    else {
        iv.freq = PROB_SCALE / 2;
        iv.cum = c ? PROB_SCALE / 2 : 0;
    }
    return iv;
}

/* ------------------------------------------------------------------------------------------
 * Stage C: rANS (rans_byte.h:47-102, ransmt.h:116-134)
 * ---------------------------------------------------------------------------------------- */
static size_t rans_block(const orc_freq* fq, int len, unsigned char* dst, unsigned char* tmp) {
    uint32_t x = RANS_L;
    unsigned char* end = tmp + 2 * RANS_BLOCK + 8;
    unsigned char* p = end;
    for (int i = len - 1; i >= 0; i--) {
        uint32_t freq = fq[i].freq, start = fq[i].cum;
        if (freq) {
            uint32_t x_max = ((RANS_L >> PROB_BITS) << 8) * freq;
            while (x >= x_max) {
                *--p = (unsigned char)(x & 0xFF);
                x >>= 8;
            }
            x = ((x / freq) << PROB_BITS) + (x % freq) + start;
        } else
            *--p = (unsigned char)start;
    }
    p -= 4;
    p[0] = (unsigned char)x;
    p[1] = (unsigned char)(x >> 8);
    p[2] = (unsigned char)(x >> 16);
    p[3] = (unsigned char)(x >> 24);
    size_t sz = (size_t)(end - p);
    memcpy(dst, p, sz);
    return sz;
}

size_t orc_rans_encode(const orc_freq* fq, size_t n, unsigned char* dst) {
    unsigned char* tmp = (unsigned char*)malloc(2 * RANS_BLOCK + 8);
    size_t out = 0;
    for (size_t b = 0; b < n; b += RANS_BLOCK) {
        size_t len = n - b < RANS_BLOCK ? n - b : RANS_BLOCK;
        out += rans_block(fq + b, (int)len, dst + out, tmp);
    }
    free(tmp);
    return out;
}

void orc_replay_events(const uint32_t* ev, size_t n, orc_freq* out, int f0) {
    Models m;
    models_init(&m, f0);
    for (size_t i = 0; i < n; i++) out[i] = models_encode(&m, ev[i]);
    free(m.color);
}

/* rANS decoder state (UseANS decode side, screencap.h:295-301, 318-359, 411-421) */
typedef struct {
    const unsigned char* p;
    uint32_t x;
    int ndec;
} RDec;

static void rdec_init(RDec* r) { /* RansDecInit rans_byte.h:105-119 */
    r->x = (uint32_t)r->p[0] | ((uint32_t)r->p[1] << 8) | ((uint32_t)r->p[2] << 16) | ((uint32_t)r->p[3] << 24);
    r->p += 4;
}
static void rdec_count(RDec* r) { /* re-init every 131072 symbols, raw bytes and bools included */
    if (++r->ndec == RANS_BLOCK) {
        rdec_init(r);
        r->ndec = 0;
    }
}
static void rdec_advance(RDec* r, uint32_t start, uint32_t freq) { /* rans_byte.h:130-146 */
    uint32_t x = r->x;
    x = freq * (x >> PROB_BITS) + (x & (PROB_SCALE - 1)) - start;
    while (x < RANS_L) x = (x << 8) | *r->p++;
    r->x = x;
}

/* ------------------------------------------------------------------------------------------
 * codec object
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int X, Y, bpp; /* bytes per pixel of the caller's format: 3 or 4 */
    int bands;     /* I-frame row bands = worker threads of the reference that is being restated (1: canonical) */
    int stride;    /* RGB24 padded row pitch (screencap.cpp:75) */
    int nbx, nby;
    unsigned fn;
    int created, version;
    int loss, loss_mask, corr_mask;
    int last_was_flat;
    unsigned char last_flat_clr[4];
    unsigned char* prev;  /* RGB24 */
    unsigned char* work;  /* RGB24 conversion buffer */
    unsigned char* bts;
    int* sxy[4];
    int* mvs[2];
    Models m;
    unsigned cx, cx1;
    /* stage outputs of the last compressed frame */
    uint32_t* ev;
    size_t nev, capev;
    orc_freq* fq;
    size_t capfq;
    unsigned char* ranstmp;
    RDec rd;
} Orc;

static void set_loss(Orc* o, int loss) { /* SetupLossMask screencap.cpp:127-139 */
    int mask = 0;
    for (int i = 0; i < loss; i++) mask = (mask << 1) | 1;
    mask = (mask << 8) + mask;
    mask = (int)(((unsigned)mask << 16) + (unsigned)mask);
    o->loss_mask = ~mask;
    int cmask = (1 << loss) >> 1;
    cmask = (cmask << 8) + cmask;
    o->corr_mask = (int)(((unsigned)cmask << 16) + (unsigned)cmask);
    o->loss = loss;
}

/* threads: I frames only.  The reference classifies an I frame in one row band per worker thread (CSquadWorker::GetSegment,
 * squad.cpp:16-31: band b covers rows [Y*b/n, Y*(b+1)/n)) and every band starts a new run (ClassifyPixelsI, screencap.cpp:876-919;
 * serialised band by band, :365-388).  1 = the canonical stream.  P frames of the multi-threaded reference depend on thread timing
 * and are not restated: they are always coded in the one-thread order. */
static int band_start(const Orc* o, int y) { /* is row y the first row of a band b >= 1?  (squad.cpp:18-22, totalsize >= nw) */
    if (o->bands <= 1 || o->Y < o->bands || y <= 0) return 0;
    long b = ((long)y * o->bands + o->Y - 1) / o->Y; /* smallest b with Y*b/n >= y */
    return b > 0 && b < o->bands && (long)o->Y * b / o->bands == y;
}

void* orc_create(int width, int height, int bits_per_pixel, int loss, int threads) {
    if (bits_per_pixel != 24 && bits_per_pixel != 32) return NULL;
    Orc* o = (Orc*)calloc(1, sizeof(Orc));
    o->bands = threads < 1 ? 1 : threads;
    o->X = width;
    o->Y = height;
    o->bpp = bits_per_pixel / 8;
    o->stride = (width * 3 + 3) & ~3;
    o->nbx = (width + 15) / 16;
    o->nby = (height + 15) / 16;
    set_loss(o, loss);
    return o;
}

static void create_codec(Orc* o, int version) { /* ScreenCodec::CreateCodec screencap.cpp:1587-1617 */
    size_t nb = (size_t)o->nbx * o->nby;
    o->version = version;
    o->prev = (unsigned char*)calloc((size_t)o->Y, (size_t)o->stride);
    o->work = (unsigned char*)calloc((size_t)o->Y, (size_t)o->stride);
    o->bts = (unsigned char*)calloc(nb, 1);
    for (int i = 0; i < 4; i++) o->sxy[i] = (int*)calloc(nb, sizeof(int));
    for (int i = 0; i < 2; i++) o->mvs[i] = (int*)calloc(nb, sizeof(int));
    models_init(&o->m, version == 3 ? 64 : 32); /* setCx6f0 screencap.cpp:1613-1614 */
    o->ranstmp = (unsigned char*)malloc(2 * RANS_BLOCK + 8);
    o->fn = 0;
    o->last_was_flat = 0;
    o->created = 1;
}

void orc_destroy(void* h) {
    Orc* o = (Orc*)h;
    if (!o) return;
    if (o->created) {
        free(o->prev);
        free(o->work);
        free(o->bts);
        for (int i = 0; i < 4; i++) free(o->sxy[i]);
        for (int i = 0; i < 2; i++) free(o->mvs[i]);
        free(o->m.color);
        free(o->ranstmp);
    }
    free(o->ev);
    free(o->fq);
    free(o);
}

size_t orc_last_events(void* h, const uint32_t** ev) {
    *ev = ((Orc*)h)->ev;
    return ((Orc*)h)->nev;
}
size_t orc_last_freqs(void* h, const orc_freq** fq) {
    *fq = ((Orc*)h)->fq;
    return ((Orc*)h)->nev;
}
const uint8_t* orc_last_bts(void* h) { return ((Orc*)h)->bts; }
const int* orc_last_sxy(void* h, int k) { return ((Orc*)h)->sxy[k]; }
const int* orc_last_mvs(void* h, int k) { return ((Orc*)h)->mvs[k]; }

/* ------------------------------------------------------------------------------------------
 * Stage A: event generation
 * ---------------------------------------------------------------------------------------- */
static void emit(Orc* o, int ctx, int sym) {
    if (o->nev == o->capev) {
        o->capev = o->capev ? o->capev * 2 : (1u << 16);
        o->ev = (uint32_t*)realloc(o->ev, o->capev * sizeof(uint32_t));
    }
    o->ev[o->nev++] = ((uint32_t)ctx << 16) | (uint32_t)sym;
}

#define MAKECX1(o) ((o)->cx1 = ((o)->cx << 6) & 0xFC0) /* screencap.h:36 */

static void emit_rgb(Orc* o, const unsigned char* p) { /* EncodeRGB screencap.cpp:631-643 */
    emit(o, ORC_CX_COLOR + 0 * 4096 + (int)(o->cx + o->cx1), p[0]);
    MAKECX1(o);
    o->cx = p[0] >> 2;
    emit(o, ORC_CX_COLOR + 1 * 4096 + (int)(o->cx + o->cx1), p[1]);
    MAKECX1(o);
    o->cx = p[1] >> 2;
    emit(o, ORC_CX_COLOR + 2 * 4096 + (int)(o->cx + o->cx1), p[2]);
    MAKECX1(o);
    o->cx = p[2] >> 2;
}

static void emit_pixel(Orc* o, int ptype, int lastptype, const unsigned char* p) { /* WritePixel screencap.cpp:609-627 */
    emit(o, ORC_CX_PTYPE + lastptype, ptype);
    if (ptype) return;
    emit_rgb(o, p);
}

static int eq3(const unsigned char* a, const unsigned char* b) { return a[0] == b[0] && a[1] == b[1] && a[2] == b[2]; }
static int grad3(const unsigned char* p, const unsigned char* l, const unsigned char* t, const unsigned char* tl) {
    return (p[0] == (int)l[0] + (int)t[0] - (int)tl[0]) && (p[1] == (int)l[1] + (int)t[1] - (int)tl[1]) &&
           (p[2] == (int)l[2] + (int)t[2] - (int)tl[2]);
}

/* GetPixelType screencap.cpp:502-521: priority last(1) -> topleft(5) -> top(2) -> gradient(4) -> literal(0) */
static int ptype_i(const unsigned char* p, const unsigned char* last, int off) {
    if (eq3(p, last)) return 1;
    if (eq3(p, p + off)) return 5;
    if (eq3(p, p + off + 3)) return 2;
    if (grad3(p, last, p + off + 3, p + off)) return 4;
    return 0;
}
/* PixelTypeFits screencap.cpp:560-574 */
static int fits_i(int t, const unsigned char* p, const unsigned char* last, int off) {
    switch (t) {
    case 0:
    case 1: return eq3(p, last);
    case 2: return eq3(p, p + off + 3);
    case 4: return grad3(p, last, p + off + 3, p + off);
    case 5: return eq3(p, p + off);
    }
    return 0;
}
/* GetPixelTypeP / P0 screencap.cpp:525-556 */
static int ptype_p(const unsigned char* p, const unsigned char* pr, int off, int notedge) {
    if (!notedge) return eq3(p, pr) ? 3 : 0;
    if (eq3(p, p - 3)) return 1;
    if (eq3(p, pr)) return 3;
    if (eq3(p, p + off)) return 5;
    if (eq3(p, p + off + 3)) return 2;
    if (grad3(p, p - 3, p + off + 3, p + off)) return 4;
    return 0;
}
/* PixelTypeFitsP / P0 screencap.cpp:578-604 */
static int fits_p(int t, const unsigned char* p, const unsigned char* pr, const unsigned char* last, int off, int notedge) {
    if (!notedge) {
        if (t == 0) return eq3(p, last);
        if (t == 3) return eq3(p, pr);
        return 0;
    }
    switch (t) {
    case 0: return eq3(p, last);
    case 1: return eq3(p, p - 3);
    case 2: return eq3(p, p + off + 3);
    case 3: return eq3(p, pr);
    case 4: return grad3(p, p - 3, p + off + 3, p + off);
    case 5: return eq3(p, p + off);
    }
    return 0;
}

/* CompressI + ClassifyPixelsI with one band (screencap.cpp:319-403, 876-919) */
static void events_i(Orc* o, const unsigned char* s) {
    const int X = o->X, Y = o->Y, stride = o->stride, off = -stride - 3;
    o->cx = o->cx1 = 0;
    emit_rgb(o, s);
    int n = 1, lasti = 0;
    for (int k = 1; k < X + 1; k++) { /* first row and one pixel: (n, rgb) pairs in ntab[0] */
        int i = (k / X) * stride + (k % X) * 3;
        if (eq3(s + i, s + lasti) && n < 255)
            n++;
        else {
            emit(o, ORC_CX_NTAB + 0, n);
            emit_rgb(o, s + i);
            n = 1;
        }
        lasti = i;
    }
    emit(o, ORC_CX_NTAB + 0, n);

    /* runs from pixel (1,1) on */
    int x = 1, y = 1, lastptype = 0;
    lasti = stride; /* pixel (0,1) */
    while (y < Y) {
        int i0 = y * stride + x * 3;
        int ptype = ptype_i(s + i0, s + lasti, off);
        /* context = bytes 1,2 of the last pixel of the previous run (screencap.cpp:371-372) */
        o->cx1 = ((unsigned)(s[lasti + 1] >> 2) << 6) & 0xFC0;
        o->cx = s[lasti + 2] >> 2;
        emit_pixel(o, ptype, lastptype, s + i0);
        lastptype = ptype;
        n = 1;
        lasti = i0;
        if (++x >= X) {
            x = 0;
            y++;
        }
        while (y < Y) {
            int i = y * stride + x * 3;
            if (x == 0 && band_start(o, y)) break; /* a new band starts a new run (screencap.cpp:881-891) */
            if (n < 255 && fits_i(ptype, s + i, s + lasti, off)) {
                n++;
                lasti = i;
                if (++x >= X) {
                    x = 0;
                    y++;
                }
            } else
                break;
        }
        emit(o, ORC_CX_NTAB + ptype, n);
    }
}

static int same_blocks(const Orc* o, const unsigned char* s, int is, int ip, int wb, int h) { /* screencap.cpp:817-825 */
    for (int y = 0; y < h; y++) {
        if (memcmp(s + is, o->prev + ip, (size_t)wb)) return 0;
        is += o->stride;
        ip += o->stride;
    }
    return 1;
}

/* FindMV screencap.cpp:684-814.  Candidate order: last_mv, MV of the block above, vertical
 * alternating up/down then rest up / rest down, horizontal left then right, +-8 box. */
static int find_mv(Orc* o, const unsigned char* s, int bi, int* lmx, int* lmy, int upperBI) {
    const int X = o->X, Y = o->Y, stride = o->stride;
    const int x1 = o->sxy[0][bi], y1 = o->sxy[1][bi], x2 = o->sxy[2][bi], y2 = o->sxy[3][bi];
    int rx1 = x1 - 8, rx2 = x1 + 8, ry1 = y1 - 8, ry2 = y1 + 8;
    if (rx1 < 0) rx1 = 0;
    if (ry1 < 0) ry1 = 0;
    if (rx2 + x2 - x1 > X) rx2 = X - x2 + x1 + 1;
    if (ry2 + y2 - y1 > Y) ry2 = Y - y2 + y1 + 1;
    int fx1 = x1 - 256, fx2 = x1 + 256, fy1 = y1 - 256, fy2 = y1 + 256;
    if (fx1 < 0) fx1 = 0;
    if (fy1 < 0) fy1 = 0;
    if (fx2 + x2 - x1 > X) fx2 = X - x2 + x1 + 1;
    if (fy2 + y2 - y1 > Y) fy2 = Y - y2 + y1 + 1;
    const int is = y1 * stride + x1 * 3, wb = (x2 - x1) * 3, h = y2 - y1;
#define TRY(xx, yy) same_blocks(o, s, is, (yy)*stride + (xx)*3, wb, h)
#define HIT(mx, my, setlast)            \
    do {                                \
        o->mvs[0][bi] = (mx);           \
        o->mvs[1][bi] = (my);           \
        if (setlast) {                  \
            *lmx = (mx);                \
            *lmy = (my);                \
        }                               \
        return 1;                       \
    } while (0)
    {
        int sx = x1 + *lmx, sy = y1 + *lmy;
        if (sx >= fx1 && sx < fx2 && sy >= fy1 && sy < fy2 && TRY(sx, sy)) HIT(*lmx, *lmy, 0);
    }
    if (upperBI >= 0 && (o->mvs[0][upperBI] != *lmx || o->mvs[1][upperBI] != *lmy)) {
        int ux = o->mvs[0][upperBI], uy = o->mvs[1][upperBI];
        int x = x1 + ux, y = y1 + uy;
        if (x >= fx1 && x < fx2 && y >= fy1 && y < fy2 && TRY(x, y)) HIT(ux, uy, 0);
    }
    int a = y1 - fy1, b = fy2 - y1 - 1;
    int common = a < b ? a : b;
    int yup = y1 - 1, ydown = y1 + 1;
    for (int k = 0; k < common; k++, yup--, ydown++) {
        if (TRY(x1, yup)) HIT(0, yup - y1, 1);
        if (TRY(x1, ydown)) HIT(0, ydown - y1, 1);
    }
    for (; yup >= fy1; yup--)
        if (TRY(x1, yup)) HIT(0, yup - y1, 1);
    for (; ydown < fy2; ydown++)
        if (TRY(x1, ydown)) HIT(0, ydown - y1, 1);
    for (int x = x1; x >= fx1; x--)
        if (TRY(x, y1)) HIT(x - x1, 0, 1);
    for (int x = x1; x < fx2; x++)
        if (TRY(x, y1)) HIT(x - x1, 0, 1);
    for (int x = x1; x >= rx1; x--) {
        for (int y = y1; y >= ry1; y--)
            if (TRY(x, y)) HIT(x - x1, y - y1, 1);
        for (int y = y1 + 1; y < ry2; y++)
            if (TRY(x, y)) HIT(x - x1, y - y1, 1);
    }
    for (int x = x1 + 1; x < rx2; x++) {
        for (int y = y1; y >= ry1; y--)
            if (TRY(x, y)) HIT(x - x1, y - y1, 1);
        for (int y = y1 + 1; y < ry2; y++)
            if (TRY(x, y)) HIT(x - x1, y - y1, 1);
    }
#undef TRY
#undef HIT
    return 0;
}

/* DecideBlockTypes, 1-thread order (screencap.cpp:928-1087): block type, changed sub-rect = exact
 * bounding box of differing pixels, motion search.  Returns 0 if nothing changed. */
static int decide_blocks(Orc* o, const unsigned char* s, int* xx1, int* xx2) {
    const int X = o->X, Y = o->Y, stride = o->stride, nbx = o->nbx, nby = o->nby;
    int bx1 = nbx, bx2 = -1, by1 = nby, by2 = -1;
    int lmx = 0, lmy = 0;
    for (int by = 0; by < nby; by++)
        for (int bx = 0; bx < nbx; bx++) {
            const int x1 = bx * 16, x2 = x1 + 16 < X ? x1 + 16 : X;
            const int y1 = by * 16, y2 = y1 + 16 < Y ? y1 + 16 : Y;
            const int bi = by * nbx + bx;
            int sx1 = x2, sx2 = -1, sy1 = y2, sy2 = -1;
            for (int y = y1; y < y2; y++)
                for (int x = x1; x < x2; x++) {
                    int i = y * stride + x * 3;
                    if (!eq3(s + i, o->prev + i)) {
                        if (x < sx1) sx1 = x;
                        if (x > sx2) sx2 = x;
                        if (y < sy1) sy1 = y;
                        if (y > sy2) sy2 = y;
                    }
                }
            if (sx2 < 0) {
                o->bts[bi] = 0;
                continue;
            }
            sx2++;
            sy2++;
            int cp;
            if (sx1 > x1 || sy1 > y1 || sx2 < x2 || sy2 < y2) {
                cp = 2;
                o->sxy[0][bi] = sx1; o->sxy[1][bi] = sy1; o->sxy[2][bi] = sx2; o->sxy[3][bi] = sy2;
            } else {
                cp = 1;
                o->sxy[0][bi] = x1; o->sxy[1][bi] = y1; o->sxy[2][bi] = x2; o->sxy[3][bi] = y2;
            }
            if (find_mv(o, s, bi, &lmx, &lmy, by > 0 ? bi - nbx : -1)) cp += 2;
            o->bts[bi] = (unsigned char)cp;
            if (bx < bx1) bx1 = bx;
            if (bx > bx2) bx2 = bx;
            if (by < by1) by1 = by;
            if (by > by2) by2 = by;
        }
    if (bx2 < 0) return 0;
    *xx1 = by1 * nbx + bx1;
    *xx2 = by2 * nbx + bx2;
    return 1;
}

/* CompressP's serialisation (screencap.cpp:1144-1248) */
static void events_p(Orc* o, const unsigned char* s, int xx1, int xx2) {
    const int stride = o->stride, nbx = o->nbx, nby = o->nby, off = -stride - 3;
    emit(o, ORC_CX_XX, xx1 & 255);
    emit(o, ORC_CX_XX, (xx1 >> 8) & 255);
    emit(o, ORC_CX_XX, xx2 & 255);
    emit(o, ORC_CX_XX, (xx2 >> 8) & 255);
    int oldt = -1, n = -1;
    for (int x = xx1; x <= xx2; x++) { /* block-type RLE */
        if (o->bts[x] == oldt && n < 255)
            n++;
        else {
            if (n > 0) emit(o, ORC_CX_NTAB2, n);
            emit(o, ORC_CX_BT, o->bts[x]);
            oldt = o->bts[x];
            n = 1;
        }
    }
    emit(o, ORC_CX_NTAB2, n);
    o->cx = o->cx1 = 0;
    int lastmx = 0, lastmy = 0;
    for (int by = 0; by < nby; by++)
        for (int bx = 0; bx < nbx; bx++) {
            const int bi = by * nbx + bx, bt = o->bts[bi];
            if (!bt) continue;
            const int x1 = o->sxy[0][bi], y1 = o->sxy[1][bi], x2 = o->sxy[2][bi], y2 = o->sxy[3][bi];
            if ((bt - 1) & 1) {
                emit(o, ORC_CX_SXY + 0, x1 - bx * 16);
                emit(o, ORC_CX_SXY + 1, y1 - by * 16);
                emit(o, ORC_CX_SXY + 2, x2 - 1 - bx * 16);
                emit(o, ORC_CX_SXY + 3, y2 - 1 - by * 16);
            }
            if ((bt - 1) & 2) {
                if (bi > 0 && o->mvs[0][bi] == lastmx && o->mvs[1][bi] == lastmy)
                    emit(o, ORC_CX_BOOL, 1);
                else {
                    emit(o, ORC_CX_BOOL, 0);
                    emit(o, ORC_CX_MV + 0, o->mvs[0][bi] + 256);
                    emit(o, ORC_CX_MV + 1, o->mvs[1][bi] + 256);
                    lastmx = o->mvs[0][bi];
                    lastmy = o->mvs[1][bi];
                }
                continue;
            }
            /* pixel runs over the sub-rect in its own raster order (screencap.cpp:1044-1064, 1216-1244) */
            int lastptype = 0, lasti = 0, started = 0, ptype = 0, runi = 0;
            n = 0;
            for (int y = y1; y < y2; y++)
                for (int x = x1; x < x2; x++) {
                    const int i = y * stride + x * 3;
                    const int notedge = x > 0 && y > 0;
                    if (started && n < 255 && fits_p(ptype, s + i, o->prev + i, s + lasti, off, notedge))
                        n++;
                    else {
                        if (started) {
                            emit_pixel(o, ptype, lastptype, s + runi);
                            lastptype = ptype;
                            emit(o, ORC_CX_NTAB + ptype, n);
                            o->cx1 = ((unsigned)(s[lasti + 1] >> 2) << 6) & 0xFC0;
                            o->cx = s[lasti + 2] >> 2;
                        }
                        ptype = ptype_p(s + i, o->prev + i, off, notedge);
                        runi = i;
                        n = 1;
                        started = 1;
                    }
                    lasti = i;
                }
            emit_pixel(o, ptype, lastptype, s + runi);
            emit(o, ORC_CX_NTAB + ptype, n);
            o->cx1 = ((unsigned)(s[lasti + 1] >> 2) << 6) & 0xFC0;
            o->cx = s[lasti + 2] >> 2;
        }
}

/* stages B + C over the collected events */
static int finish_stream(Orc* o, unsigned char* dst) {
    if (o->capfq < o->nev) {
        o->capfq = o->capev;
        o->fq = (orc_freq*)realloc(o->fq, o->capfq * sizeof(orc_freq));
    }
    for (size_t i = 0; i < o->nev; i++) o->fq[i] = models_encode(&o->m, o->ev[i]);
    size_t out = 0;
    for (size_t b = 0; b < o->nev; b += RANS_BLOCK) {
        size_t len = o->nev - b < RANS_BLOCK ? o->nev - b : RANS_BLOCK;
        out += rans_block(o->fq + b, (int)len, dst + out, o->ranstmp);
    }
    return (int)out;
}

static void do_loss(Orc* o, unsigned char* s) { /* DoLoss screencap.cpp:201-220, CMD_DOLOSS :852-861 */
    if (o->loss_mask != -1) {
        int n = o->Y * o->stride / 4;
        uint32_t* p = (uint32_t*)s;
        for (int i = 0; i < n; i++) p[i] = (p[i] & (uint32_t)o->loss_mask) | (uint32_t)o->corr_mask;
    }
    if (o->X & 3) {
        int pad = o->stride - o->X * 3;
        for (int y = 0; y < o->Y; y++) memset(s + y * o->stride + o->X * 3, 0, (size_t)pad);
    }
}

static int is_flat(const Orc* o, const unsigned char* s) { /* IsFlat screencap.cpp:1436-1444 */
    if (o->X & 3)
        return !memcmp(s, s + 3, (size_t)(o->X - 1) * 3) && !memcmp(s, s + o->stride, (size_t)(o->Y - 1) * o->stride);
    return !memcmp(s, s + 3, (size_t)o->X * o->Y * 3 - 3);
}

int orc_compress(void* h, unsigned char* src, unsigned char* dst, int dst_cap, int* ftype, int loss) {
    Orc* o = (Orc*)h;
    (void)dst_cap;
    if (loss != o->loss) set_loss(o, loss); /* ScreenCodec::CompressFrame screencap.cpp:1635-1638 */
    if (!o->created) create_codec(o, 4);    /* the encoder always writes v4, screencap.cpp:1646-1648 */
    unsigned char* s = src;
    if (o->bpp == 4) { /* RGB32 -> RGB24, alpha dropped (screencap.cpp:1652-1664) */
        for (int y = 0; y < o->Y; y++) {
            const unsigned char* in = src + (size_t)y * o->X * 4;
            unsigned char* out = o->work + (size_t)y * o->stride;
            for (int x = 0; x < o->X; x++) {
                out[0] = in[0]; out[1] = in[1]; out[2] = in[2];
                in += 4;
                out += 3;
            }
        }
        s = o->work;
    }
    const size_t fsz = (size_t)o->Y * o->stride;
    o->nev = 0;
    /* CScreenCapt::CompressFrame screencap.cpp:1456-1518 */
    if (is_flat(o, s)) {
        *ftype = 0;
        if (!(o->last_was_flat && !memcmp(s, o->last_flat_clr, 3))) {
            memcpy(o->prev, s, fsz);
            models_renew(&o->m);
            memcpy(o->last_flat_clr, s, 3);
        }
        dst[0] = (unsigned char)(1 + (o->version - 1) * 16);
        memcpy(dst + 1, s, 3);
        o->last_was_flat = 1;
        return 4;
    }
    o->last_was_flat = 0;
    if (o->fn && *ftype) {
        *ftype = 1;
        o->fn++;
        do_loss(o, s);
        if (!memcmp(s, o->prev, fsz)) { /* CMD_CMPPREV screencap.cpp:845-851, 1113-1116 */
            dst[0] = 0;
            return 1;
        }
        dst[0] = 1;
        int xx1 = 0, xx2 = 0;
        decide_blocks(o, s, &xx1, &xx2);
        events_p(o, s, xx1, xx2);
        int n = finish_stream(o, dst + 1);
        memcpy(o->prev, s, fsz);
        return n + 1;
    }
    *ftype = 0;
    o->fn++;
    dst[0] = (unsigned char)(2 + (o->version - 1) * 16);
    do_loss(o, s);
    models_renew(&o->m); /* RenewI screencap.cpp:343 */
    events_i(o, s);
    int n = finish_stream(o, dst + 1);
    memcpy(o->prev, s, fsz);
    return n + 1;
}

/* ------------------------------------------------------------------------------------------
 * decoder
 * ---------------------------------------------------------------------------------------- */
static int dec_color(Orc* o, int id) { /* UseANS::decodeC screencap.h:318-333 */
    ColorCtx* x = &o->m.color[id];
    int c;
    if (x->kind >= 4) {
        orc_freq iv;
        c = cc_find(x, (int)(o->rd.x & (PROB_SCALE - 1)));
        cc_encode_counted(x, c, &iv);
        rdec_advance(&o->rd, iv.cum, iv.freq);
    } else {
        c = *o->rd.p++;
        cc_update_raw(x, c, o->m.f0);
    }
    rdec_count(&o->rd);
    return c;
}

static int dec_fixed(Orc* o, int id) { /* UseANS::decodeF screencap.h:346-359 */
    FixedCtx* f = &o->m.fx[id - ORC_CX_NTAB];
    int c = table_find(f->cum, f->nsym, (int)(o->rd.x & (PROB_SCALE - 1)));
    orc_freq iv = fx_encode(f, c);
    rdec_advance(&o->rd, iv.cum, iv.freq);
    rdec_count(&o->rd);
    return c;
}

static int dec_bool(Orc* o) { /* screencap.h:411-421 */
    int flag = (o->rd.x & (PROB_SCALE - 1)) >= PROB_SCALE / 2;
    rdec_advance(&o->rd, flag ? PROB_SCALE / 2 : 0, PROB_SCALE / 2);
    rdec_count(&o->rd);
    return flag;
}

static void dec_rgb(Orc* o, int* r, int* g, int* b) { /* DecodeRGB screencap.cpp:662-679 */
    *r = dec_color(o, 0 * 4096 + (int)(o->cx + o->cx1));
    MAKECX1(o);
    o->cx = (unsigned)*r >> 2;
    *g = dec_color(o, 1 * 4096 + (int)(o->cx + o->cx1));
    MAKECX1(o);
    o->cx = (unsigned)*g >> 2;
    *b = dec_color(o, 2 * 4096 + (int)(o->cx + o->cx1));
    MAKECX1(o);
    o->cx = (unsigned)*b >> 2;
}

static void decode_i(Orc* o, unsigned char* d) { /* DecompressI screencap.cpp:414-498 */
    const int X = o->X, Y = o->Y, stride = o->stride, off = -stride - 3;
    int r, g, b;
    models_renew(&o->m);
    o->cx = o->cx1 = 0;
    int i = 0, k = 0, lasti = 0, n, ptype = 0, lastptype;
    while (k < X + 1) {
        dec_rgb(o, &r, &g, &b);
        n = dec_fixed(o, ORC_CX_NTAB + 0);
        for (int j = 0; j < n; j++) {
            d[i] = (unsigned char)r; d[i + 1] = (unsigned char)g; d[i + 2] = (unsigned char)b;
            k++;
            lasti = i;
            i += 3;
            if ((i % stride) >= X * 3) i = (i / stride + 1) * stride;
        }
    }
    int x = (i % stride) / 3, y = i / stride;
    while (y < Y) {
        lastptype = ptype;
        ptype = dec_fixed(o, ORC_CX_PTYPE + lastptype);
        if (!ptype) dec_rgb(o, &r, &g, &b);
        n = dec_fixed(o, ORC_CX_NTAB + ptype);
        i = y * stride + x * 3;
        while (n-- > 0) {
            switch (ptype) {
            case 0: d[i] = (unsigned char)r; d[i + 1] = (unsigned char)g; d[i + 2] = (unsigned char)b; break;
            case 1: d[i] = d[lasti]; d[i + 1] = d[lasti + 1]; d[i + 2] = d[lasti + 2]; break;
            case 2: d[i] = d[i + off + 3]; d[i + 1] = d[i + off + 4]; d[i + 2] = d[i + off + 5]; break;
            case 4:
                d[i] = (unsigned char)((int)d[lasti] + (int)d[i + off + 3] - (int)d[i + off]);
                d[i + 1] = (unsigned char)((int)d[lasti + 1] + (int)d[i + off + 4] - (int)d[i + off + 1]);
                d[i + 2] = (unsigned char)((int)d[lasti + 2] + (int)d[i + off + 5] - (int)d[i + off + 2]);
                break;
            case 5: d[i] = d[i + off]; d[i + 1] = d[i + off + 1]; d[i + 2] = d[i + off + 2]; break;
            }
            lasti = i;
            x++;
            i += 3;
            if (x >= X) {
                x = 0;
                y++;
                i = y * stride;
            }
        }
        o->cx = d[lasti + 1] >> 2;
        MAKECX1(o);
        o->cx = d[lasti + 2] >> 2;
    }
    memcpy(o->prev, d, (size_t)Y * stride);
}

static void decode_p(Orc* o, const unsigned char* src, unsigned char* d) { /* DecompressP screencap.cpp:1275-1432 */
    const int X = o->X, Y = o->Y, stride = o->stride, nbx = o->nbx, nby = o->nby, off = -stride - 3;
    if (!(src[0] & 1)) {
        memcpy(d, o->prev, (size_t)Y * stride);
        return;
    }
    o->rd.p = src + 1;
    o->rd.ndec = 0;
    rdec_init(&o->rd);
    int t = dec_fixed(o, ORC_CX_XX);
    int xx1 = (dec_fixed(o, ORC_CX_XX) << 8) + t;
    t = dec_fixed(o, ORC_CX_XX);
    int xx2 = (dec_fixed(o, ORC_CX_XX) << 8) + t;
    memset(o->bts, 0, (size_t)nbx * nby);
    for (int x = xx1; x <= xx2;) {
        int c = dec_fixed(o, ORC_CX_BT);
        int n = dec_fixed(o, ORC_CX_NTAB2);
        for (int i = 0; i < n; i++) o->bts[x++] = (unsigned char)c;
    }
    o->cx = o->cx1 = 0;
    int lastmx = 0, lastmy = 0;
    for (int by = 0; by < nby; by++)
        for (int bx = 0; bx < nbx; bx++) {
            int x1 = bx * 16, y1 = by * 16, x2 = x1 + 16, y2 = y1 + 16;
            if (x2 > X) x2 = X;
            if (y2 > Y) y2 = Y;
            const int bi = by * nbx + bx, bt = o->bts[bi];
            if (!bt || ((bt - 1) & 1))
                for (int y = y1; y < y2; y++) memcpy(d + y * stride + x1 * 3, o->prev + y * stride + x1 * 3, (size_t)(x2 - x1) * 3);
            if (!bt) continue;
            if ((bt - 1) & 1) {
                x1 = dec_fixed(o, ORC_CX_SXY + 0) + bx * 16;
                y1 = dec_fixed(o, ORC_CX_SXY + 1) + by * 16;
                x2 = dec_fixed(o, ORC_CX_SXY + 2) + bx * 16 + 1;
                y2 = dec_fixed(o, ORC_CX_SXY + 3) + by * 16 + 1;
            }
            if ((bt - 1) & 2) {
                int mx = lastmx, my = lastmy;
                if (!dec_bool(o)) {
                    mx = dec_fixed(o, ORC_CX_MV + 0) - 256;
                    my = dec_fixed(o, ORC_CX_MV + 1) - 256;
                }
                lastmx = mx;
                lastmy = my;
                for (int y = y1; y < y2; y++)
                    memcpy(d + y * stride + x1 * 3, o->prev + (y + my) * stride + (x1 + mx) * 3, (size_t)(x2 - x1) * 3);
                continue;
            }
            int x = x1, y = y1, ptype = 0, lastptype;
            while (y < y2) {
                int r = 0, g = 0, b = 0, i = y * stride + x * 3;
                lastptype = ptype;
                ptype = dec_fixed(o, ORC_CX_PTYPE + lastptype);
                if (!ptype) dec_rgb(o, &r, &g, &b);
                int n = dec_fixed(o, ORC_CX_NTAB + ptype);
                for (int c = 0; c < n; c++) {
                    switch (ptype) {
                    case 1: r = d[i - 3]; g = d[i - 2]; b = d[i - 1]; break;
                    case 2: r = d[i + off + 3]; g = d[i + off + 4]; b = d[i + off + 5]; break;
                    case 3: r = o->prev[i]; g = o->prev[i + 1]; b = o->prev[i + 2]; break;
                    case 4:
                        r = (int)d[i - 3] + (int)d[i + off + 3] - (int)d[i + off];
                        g = (int)d[i - 2] + (int)d[i + off + 4] - (int)d[i + off + 1];
                        b = (int)d[i - 1] + (int)d[i + off + 5] - (int)d[i + off + 2];
                        break;
                    case 5: r = d[i + off]; g = d[i + off + 1]; b = d[i + off + 2]; break;
                    }
                    d[i] = (unsigned char)r; d[i + 1] = (unsigned char)g; d[i + 2] = (unsigned char)b;
                    i += 3;
                    x++;
                    if (x >= x2) {
                        x = x1;
                        y++;
                        i = y * stride + x * 3;
                    }
                }
                o->cx = ((unsigned)g & 255) >> 2;
                MAKECX1(o);
                o->cx = ((unsigned)b & 255) >> 2;
            }
        }
    memcpy(o->prev, d, (size_t)Y * stride);
}

int orc_decompress(void* h, unsigned char* src, int src_len, unsigned char* dst, int pitch, int ftype) {
    Orc* o = (Orc*)h;
    (void)src_len;
    if (!o->created) { /* ScreenCodec::DecompressFrame screencap.cpp:1695-1702 */
        if (ftype > 0) return 0;
        int version = (src[0] >> 4) + 1;
        if (version != 3 && version != 4) return -version; /* v2 (range coder) is out of scope */
        create_codec(o, version);
    }
    unsigned char* d = o->work; /* RGB24 staging, then repack (screencap.cpp:1704-1739) */
    const int X = o->X, Y = o->Y, stride = o->stride;
    if (X & 3) /* screencap.cpp:1524-1528 */
        for (int y = 0; y < Y; y++) memset(d + y * stride + X * 3, 0, (size_t)(stride - X * 3));
    o->fn++;
    if (ftype) {
        o->last_was_flat = 0;
        decode_p(o, src, d);
    } else {
        int alg = src[0] & 0x0F;
        if (alg == 1) { /* flat frame screencap.cpp:1537-1553 */
            for (int y = 0; y < Y; y++)
                for (int x = 0; x < X; x++) memcpy(d + y * stride + x * 3, src + 1, 3);
            if (!(o->last_was_flat && !memcmp(o->last_flat_clr, src + 1, 3))) {
                memcpy(o->prev, d, (size_t)Y * stride);
                models_renew(&o->m);
            }
            o->last_was_flat = 1;
            memcpy(o->last_flat_clr, src + 1, 3);
        } else {
            o->last_was_flat = 0;
            o->rd.p = src + 1;
            o->rd.ndec = 0;
            rdec_init(&o->rd);
            decode_i(o, d);
        }
    }
    for (int y = 0; y < Y; y++) {
        const unsigned char* in = d + (size_t)y * stride;
        unsigned char* out = dst + (size_t)y * pitch;
        if (o->bpp == 4)
            for (int x = 0; x < X; x++) {
                out[0] = in[0]; out[1] = in[1]; out[2] = in[2]; out[3] = 255;
                in += 3;
                out += 4;
            }
        else
            memcpy(out, in, (size_t)X * 3);
    }
    return 1;
}
