// oracle/ref_capi.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C-ABI wrapper around the UNMODIFIED reference codec core, compiled from the sources where they
// lie under /root/reference (never copied into this repo) into oracle/_ref/libscpr_ref.so.
// It exposes ScreenCodec::{Init,CompressFrame,DecompressFrame,Deinit} (screencap.h:519-541) so
// tests and bench.py's cpu_baseline / --impl reference legs can drive the real reference.
// It also supplies the three symbols the excluded VfW files normally provide
// (drvproc.cpp:191-197 Set/GetThreadLocalInt, screencap.cpp:222 hmoduleSCPR, logging.cpp logF).
#include <atomic>   // (standard headers first: the windows.h shim defines min / max macros)
#include <chrono>
#include <thread>
#include <vector>

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

// ---- the reference with several clips in flight: one clip per host thread ---------------------------------------------------
// bench.py's gops_in_flight leg sets the GPU (all GOP chains of a batch of clips in one launch) beside the reference given the
// same courtesy: n_clips copies of one clip decoded by n_threads host threads, every thread with its own ScreenCodec and frame
// buffer, no Python in between.  Returns the wall-clock seconds from the moment all threads are ready until the last is done.
extern "C" double ref_decode_many(int width, int height, int bits_per_pixel, const unsigned char* stream, const unsigned* sizes,
                                  const unsigned char* ftypes, int n, int n_clips, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > n_clips) n_threads = n_clips;
    size_t total = 0;
    for (int i = 0; i < n; i++) total += sizes[i];
    std::vector<unsigned char> padded(stream, stream + total);
    padded.resize(total + 64, 0);  // the decoder's refill may read a few bytes past the end of a frame
    std::atomic<int> ready(0), next(0);
    std::atomic<bool> go(false);
    std::vector<std::thread> pool;
    const int pitch = ((width * bits_per_pixel / 8) + 3) & ~3;
    for (int t = 0; t < n_threads; t++)
        pool.emplace_back([&, t]() {
            std::vector<unsigned char> out((size_t)pitch * height);
            ready++;
            while (!go.load()) std::this_thread::yield();
            for (;;) {
                const int k = next++;
                if (k >= n_clips) break;
                void* h = ref_create(width, height, bits_per_pixel, 0, 1);
                size_t pos = 0;
                for (int i = 0; i < n; i++) {
                    ref_decompress(h, padded.data() + pos, (int)sizes[i], out.data(), pitch, ftypes[i]);
                    pos += sizes[i];
                }
                ref_destroy(h);
            }
        });
    while (ready.load() < n_threads) std::this_thread::yield();
    const auto t0 = std::chrono::steady_clock::now();
    go.store(true);
    for (auto& th : pool) th.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
