/* oracle/scpr_oracle.h -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
 *
 * C-ABI of the plain-C CPU restatement of ScreenPressor's v4 (and v3) encode/decode path.
 * Same call shape as oracle/ref_capi.cpp so tests can swap one for the other.
 */
#ifndef SCPR_ORACLE_H
#define SCPR_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Event = (context id << 16) | symbol, in bitstream order.  Context ids (shared with the CUDA
 * path, DESIGN.md "event format"): */
enum {
    ORC_CX_COLOR = 0,       /* 0..12287 : cntab[ch][cx], id = ch*4096 + cx (screencap.h:436)  */
    ORC_CX_NTAB = 12288,    /* +0..5    : ntab[ptype], 256 symbols            (screencap.h:437) */
    ORC_CX_NTAB2 = 12294,   /* block-type run lengths, 256 symbols            (screencap.h:438) */
    ORC_CX_XX = 12295,      /* changed-block index bytes, 256 symbols         (screencap.h:443) */
    ORC_CX_BT = 12296,      /* block types, 5 symbols                         (screencap.h:439) */
    ORC_CX_SXY = 12297,     /* +0..3    : sub-rect paddings, 16 symbols       (screencap.h:440) */
    ORC_CX_MV = 12301,      /* +0..1    : motion vector x / y, 512 symbols    (screencap.h:441) */
    ORC_CX_PTYPE = 12303,   /* +0..5    : pixel type given previous type, 6   (screencap.h:442) */
    ORC_CX_BOOL = 12309,    /* fixed p=1/2 flag                               (screencap.h:407) */
    ORC_NUM_CX = 12310
};

typedef struct {
    uint16_t freq, cum; /* freq == 0: raw byte `cum` (ransmt.h:125-128) */
} orc_freq;

void* orc_create(int width, int height, int bits_per_pixel, int loss, int threads); /* threads: I-frame row bands of the reference with that many workers */
void orc_destroy(void* h);
/* *ftype in: 0 = I, 1 = P request; out: actual.  Returns bytes written. */
int orc_compress(void* h, unsigned char* src, unsigned char* dst, int dst_cap, int* ftype, int loss);
/* 1 ok, 0 refused (P before I), -v bad/unsupported version */
int orc_decompress(void* h, unsigned char* src, int src_len, unsigned char* dst, int pitch, int ftype);

/* Differential hooks: stage outputs of the most recent orc_compress call. */
size_t orc_last_events(void* h, const uint32_t** ev);
size_t orc_last_freqs(void* h, const orc_freq** fq);
/* Stage A side tables of the most recent P frame: bts[nbx*nby], sxy[4][nbx*nby], mvs[2][nbx*nby]. */
const uint8_t* orc_last_bts(void* h);
const int* orc_last_sxy(void* h, int k);
const int* orc_last_mvs(void* h, int k);

/* Coverage: number of colour-model promotions kind `from` -> `to` seen so far (process-wide). */
unsigned long orc_transition_count(int from, int to);

/* Stand-alone stage entry points, used to test kernels in isolation. */
/* Replay `n` events through a fresh (RenewI) model set; writes n intervals. f0 = 32 (v4) / 64 (v3). */
void orc_replay_events(const uint32_t* ev, size_t n, orc_freq* out, int f0);
/* Encode intervals as concatenated independent rANS blocks of 131072 (ransmt.h:38,116-134). */
size_t orc_rans_encode(const orc_freq* fq, size_t n, unsigned char* dst);

#ifdef __cplusplus
}
#endif
#endif
