/* include/screencodec_b200.h -- the reference's `ScreenCodec` class (reference screencap.h:519-541), backed by libscpr_b200.so.
 *
 * Header-only C++ facade over the C ABI of scpr_c.h.  The VfW layer of the reference (`CodecInst`, screenpressor.h:15) holds one
 * `ScreenCodec sc;` and calls exactly five things on it:
 *     sc.Init(&params)                                         screenpressor.cpp:381 (CompressBegin), :575 (DecompressBegin)
 *     sc.CompressFrame(in, out, outBufSz, ftype, loss)         screenpressor.cpp:425
 *     sc.DecompressFrame(in, size, out, stride, ftype)         screenpressor.cpp:620, BadVersionException caught at :621-636
 *     sc.Deinit()                                              screenpressor.cpp:444, :647
 *     sc.CrashHappened()                                       drvproc.cpp's SEH handler
 * plus the types `CodecParameters` (screencap.h:49-55) and `BadVersionException` (screencap.h:86-90).  With this header in place of
 * screencap.h those call sites compile unchanged; tests/cpp/vfw_caller.cpp is a CodecInst-shaped caller that proves it.
 *
 * Same conventions as the reference: ftype in = 0 (I) / 1 (P) requested, out = the type coded; CompressFrame returns the byte
 * count (0 after a crash or an error, as the reference's `if (crashed) return 0`, screencap.cpp:1634); DecompressFrame returns 1,
 * or 0 for a P frame before any I frame; an undecodable stream generation throws BadVersionException(v).
 * Differences (all loud): errors of the GPU path are kept in last_error(); without a CUDA device Init() marks the object crashed.
 */
#ifndef SCREENCODEC_B200_H
#define SCREENCODEC_B200_H

#include "scpr_c.h"

#ifndef SCPR_FACADE_NO_WIN_TYPES /* the VfW sources get these from <windows.h> */
typedef unsigned char BYTE;
typedef unsigned short WORD;
typedef unsigned int uint;
#endif

/* reference screencap.h:49-55 -- same members, same order */
struct CodecParameters {
    uint width, height;
    BYTE bits_per_pixel;
    WORD redmask, greenmask, bluemask;
    uint high_range_x, high_range_y, low_range_x, low_range_y;
    uint loss;
};

/* reference screencap.h:86-90 */
class BadVersionException {
public:
    BadVersionException(int v) : version(v) {}
    int version;
};

class ScreenCodec {
    scpr_codec* c;
    bool crashed;
    int device;
    int last_rc;

    ScreenCodec(const ScreenCodec&);
    ScreenCodec& operator=(const ScreenCodec&);

public:
    ScreenCodec() : c(0), crashed(false), device(0), last_rc(0) {}
    ~ScreenCodec() { Deinit(); }

    /* which GPU the next Init() uses (no reference equivalent; default 0) */
    void SetDevice(int ordinal) { device = ordinal; }

    /* reference screencap.cpp:1565-1584 */
    void Init(CodecParameters* p) {
        Deinit();
        scpr_params q;
        q.width = p->width;
        q.height = p->height;
        q.bits_per_pixel = p->bits_per_pixel;
        q.redmask = p->redmask;
        q.greenmask = p->greenmask;
        q.bluemask = p->bluemask;
        q.high_range_x = p->high_range_x;
        q.high_range_y = p->high_range_y;
        q.low_range_x = p->low_range_x;
        q.low_range_y = p->low_range_y;
        q.loss = p->loss;
        last_rc = scpr_create(&q, device, &c);
        if (last_rc != SCPR_OK) {
            c = 0;
            crashed = true;
        } else
            crashed = false;
    }

    /* reference screencap.cpp:1619-1629 */
    void Deinit() {
        if (c) {
            scpr_destroy(c);
            c = 0;
        }
    }

    /* reference screencap.cpp:1632-1692; frame type 0-I, 1-P */
    int CompressFrame(BYTE* pSrc, BYTE* pDst, int dstLength, int& ftype, int loss) {
        if (crashed || !c) return 0;
        last_rc = scpr_compress_frame(c, pSrc, pDst, dstLength, &ftype, loss);
        return last_rc < 0 ? 0 : last_rc;
    }

    /* reference screencap.cpp:1695-1743 */
    int DecompressFrame(BYTE* pSrc, int srcLength, BYTE* pDst, int pitch, int ftype) {
        if (!c) return 0;
        last_rc = scpr_decompress_frame(c, pSrc, srcLength, pDst, pitch, ftype);
        if (last_rc < 0 && last_rc > -256) throw BadVersionException(-last_rc); /* screencap.cpp:1589-1590, caught at screenpressor.cpp:621 */
        return last_rc < 0 ? 0 : last_rc;
    }

    void CrashHappened() { crashed = true; }

    /* diagnostics the reference does not have */
    int last_status() const { return last_rc; }
    const char* last_error() const { return scpr_last_error(); }
};

/* CheckCode(conf.email, conf.regcode) (screenpressor.cpp:370): licence check of the commercial build, a no-op in the open-source one
 * (NOPROTECT, screencap.cpp:252-297). */
inline int CheckCode(const char*, const char*) { return 1; }

#endif
