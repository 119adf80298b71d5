// multi.cu -- one clip cut by contiguous frame ranges across the GPUs of a box, driven from one host process
// (BASELINE north_star: "partitioned across the 8xB200 box by contiguous frame ranges ... with host-side bitstream
// concatenation").  Host C++ only: every range is encoded / decoded by an ordinary codec object on its own
// device and host thread; nothing here touches pixels.
//
// What makes the concatenated stream byte-identical to a single codec's:
//   * cuts only where the reference starts a GOP: a requested keyframe that is not a single-colour frame (a flat frame is
//     coded as 4 bytes without restarting the frame counter, screencap.cpp:1488-1511);
//   * the reference's motion-vector array mvs[] is never cleared, not even by an I frame (screencap.cpp:96-97, 715-735): range
//     k + 1 must start its in-order motion-vector resolve with the array range k left behind.  It travels device to device
//     (cudaMemcpyPeerAsync, 8 bytes per 16x16 block) between the resolve kernels of neighbouring ranges -- the codec's
//     mvs hooks -- so only those resolves run one after the other; frame scan, motion search, typing, model replay and rANS
//     of all ranges overlap.
// Decoding needs no hand-off: a GOP is self-contained.
#include <string.h>

#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "codec.h"

using namespace scpr;

namespace {

struct Range {
    int first, count, device;
};

// is frame f a single-colour frame?  (host check of a handful of candidate cut frames)
bool host_frame_is_flat(const scpr_params& p, const uint8_t* fr) {
    const int bpp = p.bits_per_pixel / 8;
    const size_t pitch = bpp == 3 ? (((size_t)p.width * 3 + 3) & ~(size_t)3) : (size_t)p.width * bpp;
    for (uint32_t y = 0; y < p.height; y++) {
        const uint8_t* row = fr + (size_t)y * pitch;
        for (uint32_t x = 0; x < p.width; x++)
            for (int k = 0; k < (bpp == 4 ? 3 : bpp); k++)
                if (row[(size_t)x * bpp + k] != fr[k]) return false;
    }
    return true;
}

// GOP-aligned contiguous ranges, one per device at most, balanced by frame count (the same plan as shard.assign_ranges of the
// Python test harness).  cuts[f] = a range may start at frame f.
std::vector<Range> plan_ranges(int n, const std::vector<uint8_t>& cuts, const int* devices, int n_dev) {
    std::vector<int> gs, gc;  // GOP starts and lengths
    for (int f = 0; f < n; f++)
        if (f == 0 || cuts[f]) gs.push_back(f);
    for (size_t k = 0; k < gs.size(); k++) gc.push_back((k + 1 < gs.size() ? gs[k + 1] : n) - gs[k]);
    const int G = (int)gs.size(), parts = n_dev < G ? n_dev : G;
    std::vector<Range> out;
    int g = 0;
    for (int r = 0; r < parts; r++) {
        const int first = gs[g];
        const double target = (double)n * (r + 1) / parts;
        int cnt = 0;
        while (g < G - (parts - 1 - r)) {  // take GOPs while the range end stays closer to the ideal cut, leaving one for every later range
            const double end = gs[g] + gc[g], d_end = end > target ? end - target : target - end;
            const double d_start = gs[g] > target ? gs[g] - target : target - gs[g];
            if (cnt && d_end > d_start) break;
            cnt += gc[g];
            g++;
        }
        out.push_back(Range{first, cnt, devices[r]});
    }
    if (g < G) out.back().count = n - out.back().first;  // the remainder goes to the last range
    return out;
}

// joins whatever threads were started, also when starting a later one throws (a joinable std::thread must not be destroyed)
struct Pool {
    std::vector<std::thread> t;
    ~Pool() {
        for (auto& x : t)
            if (x.joinable()) x.join();
    }
};

// hand-off of mvs[] between neighbouring ranges
struct Relay {
    std::mutex m;
    std::condition_variable cv;
    std::vector<char> ready;     // range k has resolved: its mvs[] is final
    std::vector<char> failed;
    std::vector<scpr_codec*> codec;
};
struct HookCtx {
    Relay* relay;
    int k;
    int err = 0;
};

void hook_wait(void* user) {  // range k: right before its in-order resolve is enqueued
    HookCtx* h = (HookCtx*)user;
    if (h->k == 0) return;
    Relay& r = *h->relay;
    {
        std::unique_lock<std::mutex> lk(r.m);
        r.cv.wait(lk, [&] { return r.ready[h->k - 1] || r.failed[h->k - 1]; });
        if (r.failed[h->k - 1]) {
            h->err = 1;
            return;
        }
    }
    scpr_codec* me = r.codec[h->k];
    scpr_codec* prev = r.codec[h->k - 1];
    // device to device, ordered on this range's stream before the resolve kernel (the predecessor has synchronised)
    if (cudaMemcpyPeerAsync(me->mvs.p, me->device, prev->mvs.p, prev->device, (size_t)me->g.nb * sizeof(int2), me->st) != cudaSuccess) h->err = 1;
}
void hook_ready(void* user) {  // range k: its resolve has completed
    HookCtx* h = (HookCtx*)user;
    Relay& r = *h->relay;
    {
        std::lock_guard<std::mutex> lk(r.m);
        r.ready[h->k] = 1;
    }
    r.cv.notify_all();
}

}  // namespace

// ---- a standing set of codec objects, one per device ordinal given (created once: device workspaces and model states stay) ----
struct scpr_multi {
    scpr_params p;
    std::vector<int> devices;
    std::vector<scpr_codec*> enc, dec;   // one of each per entry of `devices`
};

namespace {

int64_t multi_compress(scpr_multi* m, const uint8_t* frames, int n, const uint8_t* keyflags, uint8_t* dst, size_t dst_cap, uint32_t* sizes,
                       uint8_t* ftypes, int* range_first, int* n_ranges) {
    const scpr_params* p = &m->p;
    const int bpp = p->bits_per_pixel / 8;
    const size_t pitch = bpp == 3 ? (((size_t)p->width * 3 + 3) & ~(size_t)3) : (size_t)p->width * bpp;
    const size_t fb = pitch * p->height;
    std::vector<uint8_t> cuts(n, 0);
    for (int f = 1; f < n; f++) cuts[f] = keyflags[f] && !host_frame_is_flat(*p, frames + (size_t)f * fb);
    const std::vector<Range> ranges = plan_ranges(n, cuts, m->devices.data(), (int)m->devices.size());
    const int R = (int)ranges.size();
    if (n_ranges) *n_ranges = R;
    if (range_first)
        for (int k = 0; k < R; k++) range_first[k] = ranges[k].first;
    Relay relay;
    relay.ready.assign(R, 0);
    relay.failed.assign(R, 0);
    relay.codec.assign(R, nullptr);
    std::vector<HookCtx> ctx(R);
    std::vector<std::unique_ptr<uint8_t[]>> out(R);
    std::vector<size_t> out_cap(R, 0);
    std::vector<int64_t> used(R, 0);
    for (int k = 0; k < R; k++) {
        relay.codec[k] = m->enc[k];  // range k runs on the k-th device given (plan_ranges hands them out in order)
        const int r = scpr_reset(relay.codec[k]);
        if (r < 0) {
            for (int q = 0; q < k; q++) scpr_set_mvs_hooks(relay.codec[q], nullptr, nullptr, nullptr);  // they point into this frame
            return r;
        }
        ctx[k].relay = &relay;
        ctx[k].k = k;
        scpr_set_mvs_hooks(relay.codec[k], hook_wait, hook_ready, &ctx[k]);
        // worst case per frame is W*H*6 (CompressGetSize) but never more than the caller's whole buffer; the pages of an
        // untouched new[] are not committed
        size_t cap = (size_t)ranges[k].count * scpr_max_compressed_size(p);
        if (cap > dst_cap) cap = dst_cap;
        out_cap[k] = cap;
        out[k].reset(new uint8_t[cap + 16]);
    }
    Pool pool;
    for (int k = 0; k < R; k++)
        pool.t.emplace_back([&, k]() {
            const Range& rg = ranges[k];
            std::vector<uint8_t> keys(keyflags + rg.first, keyflags + rg.first + rg.count);
            keys[0] = 1;
            // a fresh codec treats its first frame as a keyframe anyway; for k > 0 the cut guarantees the reference does too
            int64_t r = scpr_compress_clip(relay.codec[k], frames + (size_t)rg.first * fb, rg.count, keys.data(), out[k].get(), out_cap[k],
                                           sizes + rg.first, ftypes + rg.first);
            if (r >= 0 && ctx[k].err) r = SCPR_E_CUDA;
            used[k] = r;
            if (r < 0) {
                {
                    std::lock_guard<std::mutex> lk(relay.m);
                    relay.failed[k] = 1;
                }
                relay.cv.notify_all();
            }
        });
    for (auto& t : pool.t) t.join();
    for (int k = 0; k < R; k++) scpr_set_mvs_hooks(relay.codec[k], nullptr, nullptr, nullptr);
    int64_t total = 0, err = 0;
    for (int k = 0; k < R; k++) {
        if (used[k] < 0 && !err) err = used[k];
        if (used[k] > 0) total += used[k];
    }
    if (!err && (size_t)total > dst_cap) {
        set_error("destination too small: need %lld bytes", (long long)total);
        err = SCPR_E_DSTSIZE;
    }
    if (!err) {  // host-side concatenation in frame order
        size_t pos = 0;
        for (int k = 0; k < R; k++) {
            memcpy(dst + pos, out[k].get(), (size_t)used[k]);
            pos += (size_t)used[k];
        }
    }
    return err ? err : total;
}

int multi_decompress(scpr_multi* m, const uint8_t* stream, const uint32_t* sizes, const uint8_t* ftypes, int n, uint8_t* frames, int pitch) {
    const scpr_params* p = &m->p;
    // a range may start at any I frame that is not a flat frame (a flat frame's successors can be P frames on older models)
    std::vector<uint8_t> cuts(n, 0);
    std::vector<size_t> off(n + 1, 0);
    for (int f = 0; f < n; f++) off[f + 1] = off[f] + sizes[f];
    for (int f = 1; f < n; f++) cuts[f] = ftypes[f] == 0 && sizes[f] >= 1 && (stream[off[f]] & 0x0F) == 2;
    const std::vector<Range> ranges = plan_ranges(n, cuts, m->devices.data(), (int)m->devices.size());
    const int R = (int)ranges.size();
    std::vector<int> res(R, 1);
    Pool pool;
    const size_t fb = (size_t)pitch * p->height;
    for (int k = 0; k < R; k++)
        pool.t.emplace_back([&, k]() {
            const Range& rg = ranges[k];
            scpr_codec* c = m->dec[k];
            int r = scpr_reset(c);
            if (r >= 0) r = scpr_decompress_clip(c, stream + off[rg.first], sizes + rg.first, ftypes + rg.first, rg.count, frames + (size_t)rg.first * fb, pitch);
            res[k] = r;
        });
    for (auto& t : pool.t) t.join();
    for (int k = 0; k < R; k++)
        if (res[k] != 1) return res[k];
    return 1;
}

}  // namespace

extern "C" {

int scpr_plan_ranges(const uint8_t* cuts, int n, int n_ranges_max, int* first, int* count) {
    if (!cuts || n < 1 || n_ranges_max < 1 || !first || !count) return SCPR_E_PARAM;
    try {
        std::vector<uint8_t> c(cuts, cuts + n);
        std::vector<int> dev(n_ranges_max, 0);
        const std::vector<Range> r = plan_ranges(n, c, dev.data(), n_ranges_max);
        for (size_t k = 0; k < r.size(); k++) {
            first[k] = r[k].first;
            count[k] = r[k].count;
        }
        return (int)r.size();
    } catch (...) {
        set_error("out of memory");
        return SCPR_E_PARAM;
    }
}

int scpr_multi_create(const scpr_params* p, const int* devices, int n_dev, scpr_multi** out) {
    if (!p || !devices || n_dev < 1 || !out) return SCPR_E_PARAM;
    *out = nullptr;
    try {
        if (p->bits_per_pixel == 16) {
            set_error("frame-range splitting of 16 bpp clients is not built");
            return SCPR_E_UNSUPPORTED;
        }
        std::unique_ptr<scpr_multi> m(new scpr_multi());
        m->p = *p;
        m->devices.assign(devices, devices + n_dev);
        for (int k = 0; k < n_dev; k++) {
            scpr_codec *e = nullptr, *d = nullptr;
            int r = scpr_create(p, devices[k], &e);
            if (r >= 0) {
                m->enc.push_back(e);
                r = scpr_create(p, devices[k], &d);
            }
            if (r >= 0) m->dec.push_back(d);
            if (r < 0) {
                for (auto c : m->enc) scpr_destroy(c);
                for (auto c : m->dec) scpr_destroy(c);
                return r;
            }
        }
        *out = m.release();
        return SCPR_OK;
    } catch (...) {
        set_error("out of memory");
        return SCPR_E_PARAM;
    }
}

void scpr_multi_destroy(scpr_multi* m) {
    if (!m) return;
    for (auto c : m->enc) scpr_destroy(c);
    for (auto c : m->dec) scpr_destroy(c);
    delete m;
}

int64_t scpr_multi_compress_clip(scpr_multi* m, const uint8_t* frames, int n, const uint8_t* keyflags, uint8_t* dst, size_t dst_cap, uint32_t* sizes,
                                 uint8_t* ftypes, int* range_first, int* n_ranges) {
    if (!m || !frames || !keyflags || !dst || !sizes || !ftypes || n < 0) return SCPR_E_PARAM;
    if (n == 0) return 0;
    try {
        return multi_compress(m, frames, n, keyflags, dst, dst_cap, sizes, ftypes, range_first, n_ranges);
    } catch (...) {
        set_error("out of memory");
        return SCPR_E_PARAM;
    }
}

int scpr_multi_decompress_clip(scpr_multi* m, const uint8_t* stream, const uint32_t* sizes, const uint8_t* ftypes, int n, uint8_t* frames, int pitch) {
    if (!m || !stream || !sizes || !ftypes || !frames || n < 0) return SCPR_E_PARAM;
    if (n == 0) return 1;
    try {
        return multi_decompress(m, stream, sizes, ftypes, n, frames, pitch);
    } catch (...) {
        set_error("out of memory");
        return SCPR_E_PARAM;
    }
}

// one-shot forms: create the set, run, destroy
int64_t scpr_compress_clip_multi(const scpr_params* p, const int* devices, int n_dev, const uint8_t* frames, int n, const uint8_t* keyflags,
                                 uint8_t* dst, size_t dst_cap, uint32_t* sizes, uint8_t* ftypes, int* range_first, int* n_ranges) {
    if (!p || !devices || n_dev < 1 || !frames || !keyflags || !dst || !sizes || !ftypes || n < 0) return SCPR_E_PARAM;
    if (n == 0) return 0;
    scpr_multi* m = nullptr;
    const int r = scpr_multi_create(p, devices, n_dev, &m);
    if (r < 0) return r;
    const int64_t used = scpr_multi_compress_clip(m, frames, n, keyflags, dst, dst_cap, sizes, ftypes, range_first, n_ranges);
    scpr_multi_destroy(m);
    return used;
}

int scpr_decompress_clip_multi(const scpr_params* p, const int* devices, int n_dev, const uint8_t* stream, const uint32_t* sizes,
                               const uint8_t* ftypes, int n, uint8_t* frames, int pitch) {
    if (!p || !devices || n_dev < 1 || !stream || !sizes || !ftypes || !frames || n < 0) return SCPR_E_PARAM;
    if (n == 0) return 1;
    scpr_multi* m = nullptr;
    int r = scpr_multi_create(p, devices, n_dev, &m);
    if (r < 0) return r;
    r = scpr_multi_decompress_clip(m, stream, sizes, ftypes, n, frames, pitch);
    scpr_multi_destroy(m);
    return r;
}

}  // extern "C"
