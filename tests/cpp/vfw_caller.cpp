// tests/cpp/vfw_caller.cpp -- a CodecInst-shaped C++ caller of include/screencodec_b200.h (TEST HARNESS, host C++ only).
//
// Drives the facade exactly as the reference's VfW layer drives its ScreenCodec member:
//   CompressBegin   -> fills CodecParameters (ranges 256/256/8/8) and sc.Init(&params)         screenpressor.cpp:343-384
//   Compress        -> keyframe decision, quality -> loss, sc.CompressFrame(in, out, outBufSz, ftype, loss),
//                      AVIIF_KEYFRAME / npframes bookkeeping                                    screenpressor.cpp:392-437
//   CompressEnd     -> sc.Deinit()                                                              screenpressor.cpp:441-447
//   DecompressBegin -> sc.Init(&params) with loss 0                                             screenpressor.cpp:560-577
//   Decompress      -> InferFrameType, stride, sc.DecompressFrame(...), catch (BadVersionException)  screenpressor.cpp:579-640
//   DecompressEnd   -> sc.Deinit()                                                              screenpressor.cpp:644-650
// The few VfW structs it needs are declared here with the fields those functions touch.
//
// usage: vfw_caller W H BPP N KF_INTERVAL QUALITY frames.raw out.stream out.index out.decoded
//   frames.raw : N frames, rows top to bottom, pitch (W*BPP/8 + 3) & ~3
//   out.index  : N x {uint32 size, uint32 flags}   (flags = AVIIF_KEYFRAME for key frames)
//   out.decoded: the N frames decoded again from out.stream through Decompress
// `vfw_caller --nodevice` checks the no-GPU behaviour: Init marks the object crashed, CompressFrame returns 0.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "screencodec_b200.h"

typedef unsigned int DWORD;
typedef int LONG;
enum { ICERR_OK = 0, ICERR_BADFORMAT = -2 };
enum { ICCOMPRESS_KEYFRAME = 1, ICDECOMPRESS_NOTKEYFRAME = 0x08000000, AVIIF_KEYFRAME = 0x10 };

struct BITMAPINFOHEADER {
    DWORD biSize;
    LONG biWidth, biHeight;
    WORD biPlanes, biBitCount;
    DWORD biCompression, biSizeImage;
};
struct ICCOMPRESS {
    DWORD dwFlags;
    BITMAPINFOHEADER* lpbiOutput;
    void* lpOutput;
    BITMAPINFOHEADER* lpbiInput;
    void* lpInput;
    DWORD* lpckid;
    DWORD* lpdwFlags;
    LONG lFrameNum;
    DWORD dwFrameSize, dwQuality;
};
struct ICDECOMPRESS {
    DWORD dwFlags;
    BITMAPINFOHEADER* lpbiInput;
    void* lpInput;
    BITMAPINFOHEADER* lpbiOutput;
    void* lpOutput;
};

class CodecInst {
    ScreenCodec sc;
    int kf_interval, npframes;
    bool force_interval, force_loss, decompressing;
    int conf_loss;
    WORD rmask, gmask, bmask;
    DWORD size_image;

public:
    int bad_version;
    CodecInst(int interval) : kf_interval(interval), npframes(0), force_interval(true), force_loss(false), decompressing(false), conf_loss(0),
                              rmask(0x7C00), gmask(0x3E0), bmask(0x1F), size_image(0), bad_version(0) {}

    DWORD CompressBegin(BITMAPINFOHEADER* in) {
        CompressEnd();
        npframes = 0;
        CheckCode("", "");
        CodecParameters params;
        params.width = in->biWidth;
        params.height = in->biHeight;
        params.bits_per_pixel = (BYTE)in->biBitCount;
        params.redmask = rmask;
        params.greenmask = gmask;
        params.bluemask = bmask;
        params.high_range_x = 256;
        params.high_range_y = 256;
        params.low_range_x = 8;
        params.low_range_y = 8;
        params.loss = conf_loss;
        sc.Init(&params);
        return ICERR_OK;
    }
    DWORD Compress(ICCOMPRESS* ic) {
        BYTE* const in = (BYTE*)ic->lpInput;
        BYTE* const out = (BYTE*)ic->lpOutput;
        int ftype = 1;
        const bool forced_kf = force_interval && (npframes + 1 >= kf_interval);
        const bool host_kf = !force_interval && (ic->dwFlags & ICCOMPRESS_KEYFRAME);
        if (host_kf || forced_kf) ftype = 0;
        const DWORD outBufSz = std::max(ic->lpbiOutput->biSizeImage, ic->lpbiInput->biSizeImage);
        int loss = conf_loss;
        if (!force_loss) {
            const DWORD quality = std::min(ic->dwQuality, (DWORD)10000);
            loss = std::min((int)((10000 - quality) / 2000), 4);
        }
        const int sz = sc.CompressFrame(in, out, (int)outBufSz, ftype, loss);
        if (!ftype) {
            *ic->lpdwFlags = AVIIF_KEYFRAME;
            npframes = 0;
        } else {
            *ic->lpdwFlags = 0;
            npframes++;
        }
        ic->lpbiOutput->biSizeImage = sz;
        return ICERR_OK;
    }
    DWORD CompressEnd() {
        sc.Deinit();
        return ICERR_OK;
    }

    DWORD DecompressBegin(BITMAPINFOHEADER* in) {
        DecompressEnd();
        CodecParameters params;
        params.width = in->biWidth;
        params.height = in->biHeight;
        params.bits_per_pixel = (BYTE)in->biBitCount;
        params.redmask = rmask;
        params.greenmask = gmask;
        params.bluemask = bmask;
        params.high_range_x = 256;
        params.high_range_y = 256;
        params.low_range_x = 8;
        params.low_range_y = 8;
        params.loss = 0;
        sc.Init(&params);
        size_image = (DWORD)(((in->biWidth * in->biBitCount / 8 + 3) & ~3) * in->biHeight);
        decompressing = true;
        return ICERR_OK;
    }
    static int InferFrameType(BYTE first_byte, DWORD data_size) {
        switch (first_byte) {
            case 0: return 1;
            case 1: return data_size <= 4 ? 0 : 1;
            case 0x02:
            case 0x11:
            case 0x12: return 0;
        }
        return -1;
    }
    DWORD Decompress(ICDECOMPRESS* ic) {
        try {
            if (!decompressing) {
                const DWORD r = DecompressBegin(ic->lpbiInput);
                if (r != ICERR_OK) return r;
            }
            ic->lpbiOutput->biSizeImage = size_image;
            BYTE* const in = (BYTE*)ic->lpInput;
            BYTE* out = (BYTE*)ic->lpOutput;
            int ftype = 0;
            if (ic->dwFlags & ICDECOMPRESS_NOTKEYFRAME) ftype = 1;
            const int inferred = InferFrameType(in[0], ic->lpbiInput->biSizeImage);
            if (inferred >= 0) ftype = inferred;
            const int bits = ic->lpbiInput->biBitCount;
            const int stride = (ic->lpbiInput->biWidth * bits / 8 + 3) & (~3);
            sc.DecompressFrame(in, (int)ic->lpbiInput->biSizeImage, out, stride, ftype);
        } catch (BadVersionException bve) {
            bad_version = bve.version;
            return (DWORD)ICERR_BADFORMAT;
        }
        return ICERR_OK;
    }
    DWORD DecompressEnd() {
        sc.Deinit();
        decompressing = false;
        return ICERR_OK;
    }
    ScreenCodec& codec() { return sc; }
};

static int fail(const char* what) {
    fprintf(stderr, "vfw_caller: %s\n", what);
    return 2;
}

int main(int argc, char** argv) {
    if (argc == 2 && !strcmp(argv[1], "--nodevice")) {
        // without a CUDA device: Init() must fail loudly (object crashed, CompressFrame returns 0) -- never a CPU path
        CodecInst ci(500);
        BITMAPINFOHEADER bi = {40, 64, 64, 1, 32, 0, 64 * 64 * 4};
        ci.CompressBegin(&bi);
        std::vector<BYTE> in(64 * 64 * 4, 7), out(64 * 64 * 6);
        int ftype = 1;
        const int sz = ci.codec().CompressFrame(in.data(), out.data(), (int)out.size(), ftype, 0);
        printf("status %d size %d error %s\n", ci.codec().last_status(), sz, ci.codec().last_error());
        return sz == 0 && ci.codec().last_status() == SCPR_E_NODEVICE ? 0 : 1;
    }
    if (argc != 11) return fail("usage: vfw_caller W H BPP N KF_INTERVAL QUALITY frames.raw out.stream out.index out.decoded");
    const int W = atoi(argv[1]), H = atoi(argv[2]), BPP = atoi(argv[3]), N = atoi(argv[4]), KF = atoi(argv[5]);
    const DWORD quality = (DWORD)atoi(argv[6]);
    const size_t stride = ((size_t)W * BPP / 8 + 3) & ~(size_t)3, fb = stride * H;
    std::vector<BYTE> frames(fb * N);
    FILE* f = fopen(argv[7], "rb");
    if (!f || fread(frames.data(), 1, frames.size(), f) != frames.size()) return fail("cannot read the frames");
    fclose(f);

    // ---- capture side: ICM_COMPRESS_BEGIN, N x ICM_COMPRESS, ICM_COMPRESS_END
    CodecInst enc(KF);
    BITMAPINFOHEADER bin = {40, W, H, 1, (WORD)BPP, 0, (DWORD)fb};
    BITMAPINFOHEADER bout = {40, W, H, 1, (WORD)BPP, 0x52504353 /* 'SCPR' */, (DWORD)((size_t)W * H * 6)};  // CompressGetSize
    if (enc.CompressBegin(&bin) != ICERR_OK) return fail("CompressBegin");
    std::vector<BYTE> obuf((size_t)W * H * 6 + 64);
    std::vector<BYTE> stream;
    std::vector<DWORD> index;
    for (int i = 0; i < N; i++) {
        DWORD flags = 0, ckid = 0;
        bout.biSizeImage = (DWORD)((size_t)W * H * 6);
        ICCOMPRESS ic = {0, &bout, obuf.data(), &bin, frames.data() + fb * i, &ckid, &flags, i, 0, quality};
        if (enc.Compress(&ic) != ICERR_OK) return fail("Compress");
        if (bout.biSizeImage == 0) {
            fprintf(stderr, "frame %d: CompressFrame returned 0 (status %d: %s)\n", i, enc.codec().last_status(), enc.codec().last_error());
            return 3;
        }
        stream.insert(stream.end(), obuf.begin(), obuf.begin() + bout.biSizeImage);
        index.push_back(bout.biSizeImage);
        index.push_back(flags);
    }
    enc.CompressEnd();

    // ---- playback side: N x ICM_DECOMPRESS (the first one opens the decoder), ICM_DECOMPRESS_END
    CodecInst dec(KF);
    std::vector<BYTE> decoded(fb * N), cur(fb);
    size_t pos = 0;
    for (int i = 0; i < N; i++) {
        BITMAPINFOHEADER din = {40, W, H, 1, (WORD)BPP, 0x52504353, index[2 * i]};
        BITMAPINFOHEADER dout = {40, W, H, 1, (WORD)BPP, 0, 0};
        std::vector<BYTE> chunk(stream.begin() + pos, stream.begin() + pos + index[2 * i]);
        chunk.resize(chunk.size() + 16);
        ICDECOMPRESS id = {(index[2 * i + 1] & AVIIF_KEYFRAME) ? 0u : (DWORD)ICDECOMPRESS_NOTKEYFRAME, &din, chunk.data(), &dout, cur.data()};
        if (dec.Decompress(&id) != ICERR_OK) return fail("Decompress");
        if (dec.codec().last_status() != 1) {
            fprintf(stderr, "frame %d: DecompressFrame status %d: %s\n", i, dec.codec().last_status(), dec.codec().last_error());
            return 4;
        }
        memcpy(decoded.data() + fb * i, cur.data(), fb);
        pos += index[2 * i];
    }
    // a stream generation nobody can decode must surface as BadVersionException -> ICERR_BADFORMAT (screenpressor.cpp:621-636)
    {
        BYTE bogus[8] = {0x92, 0, 0, 0, 0, 0, 0, 0};  // version nibble 9
        BITMAPINFOHEADER din = {40, W, H, 1, (WORD)BPP, 0x52504353, 8};
        BITMAPINFOHEADER dout = {40, W, H, 1, (WORD)BPP, 0, 0};
        CodecInst d2(KF);
        ICDECOMPRESS id = {0, &din, bogus, &dout, cur.data()};
        const DWORD r = d2.Decompress(&id);
        printf("bogus stream: result %d version %d\n", (int)r, d2.bad_version);
        d2.DecompressEnd();
    }
    dec.DecompressEnd();

    FILE* o = fopen(argv[8], "wb");
    if (!o || fwrite(stream.data(), 1, stream.size(), o) != stream.size()) return fail("cannot write the stream");
    fclose(o);
    o = fopen(argv[9], "wb");
    if (!o || fwrite(index.data(), 4, index.size(), o) != index.size()) return fail("cannot write the index");
    fclose(o);
    o = fopen(argv[10], "wb");
    if (!o || fwrite(decoded.data(), 1, decoded.size(), o) != decoded.size()) return fail("cannot write the decoded frames");
    fclose(o);
    printf("frames %d bytes %zu\n", N, stream.size());
    return 0;
}
