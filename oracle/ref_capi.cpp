// oracle/ref_capi.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C-ABI wrapper around the UNMODIFIED reference codec core, compiled from the sources where they
// lie under /root/reference (never copied into this repo) into oracle/_ref/libscpr_ref.so.
// It exposes ScreenCodec::{Init,CompressFrame,DecompressFrame,Deinit} (screencap.h:519-541) so
// tests and bench.py's cpu_baseline / --impl reference legs can drive the real reference.
// It also supplies the three symbols the excluded VfW files normally provide
// (drvproc.cpp:191-197 Set/GetThreadLocalInt, screencap.cpp:222 hmoduleSCPR, logging.cpp logF).
#include "screencap.h"

FILE* logF = NULL;
HMODULE hmoduleSCPR = NULL;
int g_shim_nproc = 1;

static thread_local int t_tls_int = 0;
void SetThreadLocalInt(int v) { t_tls_int = v; }
int GetThreadLocalInt() { return t_tls_int; }

extern "C" {

// nthreads: value reported as dwNumberOfProcessors when the squad is created on the first
// CompressFrame (screencap.cpp:1458-1462). 1 = canonical bitstream.
void* ref_create(int width, int height, int bits_per_pixel, int loss, int nthreads) {
    CodecParameters p;
    memset(&p, 0, sizeof(p));
    p.width = width;
    p.height = height;
    p.bits_per_pixel = (BYTE)bits_per_pixel;
    p.redmask = 0x7C00; p.greenmask = 0x3E0; p.bluemask = 0x1F;
    // screenpressor.cpp:374-379
    p.high_range_x = 256; p.high_range_y = 256; p.low_range_x = 8; p.low_range_y = 8;
    p.loss = loss;
    g_shim_nproc = nthreads < 1 ? 1 : nthreads;
    ScreenCodec* sc = new ScreenCodec();
    sc->Init(&p);
    return sc;
}

void ref_set_threads(int nthreads) { g_shim_nproc = nthreads < 1 ? 1 : nthreads; }

void ref_destroy(void* h) { delete (ScreenCodec*)h; }

// returns bytes written; *ftype in: 0=I 1=P request, out: actual
int ref_compress(void* h, unsigned char* src, unsigned char* dst, int dst_cap, int* ftype, int loss) {
    int ft = *ftype;
    int n = ((ScreenCodec*)h)->CompressFrame(src, dst, dst_cap, ft, loss);
    *ftype = ft;
    return n;
}

// returns 1 ok, 0 refused, -version on BadVersionException
int ref_decompress(void* h, unsigned char* src, int src_len, unsigned char* dst, int pitch, int ftype) {
    try {
        return ((ScreenCodec*)h)->DecompressFrame(src, src_len, dst, pitch, ftype);
    } catch (BadVersionException& e) {
        return -e.version;
    }
}

}  // extern "C"
