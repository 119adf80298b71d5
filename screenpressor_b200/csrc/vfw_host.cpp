// vfw_host.cpp -- host side above the codec object: the policy of the reference's VfW layer and the AVI container
// its streams live in.  Plain C++ (no CUDA here); the frames still go through scpr_compress_frame /
// scpr_decompress_frame, i.e. through the kernels.
//
// Mirrors (reference):
//   CodecInst::CompressBegin / Compress      screenpressor.cpp:343-384, 392-437  (keyframe policy, quality -> loss)
//   CodecInst::InferFrameType / Decompress   screenpressor.cpp:579-620
//   CodecInst::CanCompress / CompressGetFormat  screenpressor.cpp:276-339         (16 bpp masks travel in the format header)
//   Configuration defaults                   conf.h:7, 20-21                      (interval 500, forced; loss 0, forced)
// and writes / reads the RIFF AVI layout VfW hosts produce for it: one 'vids' stream, handler and biCompression 'SCPR'
// (screenpressor.h:6), '00dc' chunks, AVIIF_KEYFRAME in 'idx1' exactly where Compress set it.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/scpr_c.h"

namespace scpr {
void set_error(const char* fmt, ...);
}

static const uint32_t AVIIF_KEYFRAME = 0x10;
static inline uint32_t fcc(const char* s) { return (uint32_t)(uint8_t)s[0] | ((uint32_t)(uint8_t)s[1] << 8) | ((uint32_t)(uint8_t)s[2] << 16) | ((uint32_t)(uint8_t)s[3] << 24); }

struct scpr_session {
    scpr_codec* enc = nullptr;
    scpr_codec* dec = nullptr;
    scpr_params p;
    scpr_policy pol;
    int device = 0;
    int npframes = 0;  // P frames since the last keyframe (CodecInst::npframes)
};

struct AviIndexEntry {
    uint64_t off;   // file offset of the chunk's data
    uint32_t size;
    uint32_t flags;
};

struct scpr_avi {
    FILE* f = nullptr;
    bool writing = false;
    scpr_avi_info info;
    std::vector<AviIndexEntry> idx;
    uint64_t movi_pos = 0;   // file offset of the 'movi' fourcc
    uint32_t max_chunk = 0;
    size_t hdr_bytes = 0;
    bool io_failed = false;  // writer: a chunk could not be written; close reports it instead of leaving a silent truncated file
};

static void put32(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static void put16(uint8_t* p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static uint32_t get32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint32_t get16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

// header block: RIFF/AVI, LIST hdrl { avih, LIST strl { strh, strf } }, LIST movi.  Sizes and counts are patched on close.
static std::vector<uint8_t> build_header(const scpr_avi_info& in, uint32_t frames, uint32_t max_chunk, uint32_t movi_bytes, uint32_t riff_bytes) {
    const uint32_t nextra = in.bits_per_pixel == 16 ? 12 : 0;  // three mask DWORDs (CompressGetFormat, screenpressor.cpp:318-336)
    const uint32_t strf = 40 + nextra, strl = 4 + (8 + 56) + (8 + strf), hdrl = 4 + (8 + 56) + (8 + strl);
    std::vector<uint8_t> h(12 + 8 + hdrl + 12, 0);
    uint8_t* p = h.data();
    put32(p, fcc("RIFF")); put32(p + 4, riff_bytes); put32(p + 8, fcc("AVI ")); p += 12;
    put32(p, fcc("LIST")); put32(p + 4, hdrl); put32(p + 8, fcc("hdrl")); p += 12;
    put32(p, fcc("avih")); put32(p + 4, 56); p += 8;
    const uint32_t stride = ((in.width * in.bits_per_pixel / 8) + 3) & ~3u;
    put32(p + 0, in.fps_num ? (uint32_t)(1000000.0 * in.fps_den / in.fps_num + 0.5) : 0);  // dwMicroSecPerFrame
    put32(p + 4, 0);                   // dwMaxBytesPerSec
    put32(p + 12, 0x10);               // dwFlags = AVIF_HASINDEX
    put32(p + 16, frames);             // dwTotalFrames
    put32(p + 24, 1);                  // dwStreams
    put32(p + 28, max_chunk);          // dwSuggestedBufferSize
    put32(p + 32, in.width); put32(p + 36, in.height);
    p += 56;
    put32(p, fcc("LIST")); put32(p + 4, strl); put32(p + 8, fcc("strl")); p += 12;
    put32(p, fcc("strh")); put32(p + 4, 56); p += 8;
    put32(p + 0, fcc("vids")); put32(p + 4, fcc("SCPR"));
    put32(p + 20, in.fps_den); put32(p + 24, in.fps_num);   // dwScale, dwRate
    put32(p + 32, frames);                                  // dwLength
    put32(p + 36, max_chunk); put32(p + 40, 0xFFFFFFFFu);   // dwSuggestedBufferSize, dwQuality
    put16(p + 52, in.width); put16(p + 54, in.height);      // rcFrame right, bottom
    p += 56;
    put32(p, fcc("strf")); put32(p + 4, strf); p += 8;
    put32(p + 0, strf); put32(p + 4, in.width); put32(p + 8, in.height);
    put16(p + 12, 1); put16(p + 14, in.bits_per_pixel);
    put32(p + 16, fcc("SCPR")); put32(p + 20, stride * in.height);
    if (nextra) { put32(p + 40, in.redmask); put32(p + 44, in.greenmask); put32(p + 48, in.bluemask); }
    p += strf;
    put32(p, fcc("LIST")); put32(p + 4, movi_bytes); put32(p + 8, fcc("movi"));
    return h;
}

extern "C" {

// ---- policy --------------------------------------------------------------------------------------------------------
int scpr_quality_to_loss(uint32_t quality) {  // screenpressor.cpp:410-422
    if (quality > 10000) quality = 10000;
    const uint32_t l = (10000 - quality) / 2000;
    return (int)(l < 4 ? l : 4);
}

int scpr_infer_frame_type(uint8_t first_byte, uint32_t data_size) {  // screenpressor.cpp:579-589; 0 = I, 1 = P, -1 = unknown
    switch (first_byte) {
    case 0: return 1;
    case 1: return data_size <= 4 ? 0 : 1;
    case 0x02:
    case 0x11:
    case 0x12: return 0;
    }
    return -1;
}

void scpr_policy_default(scpr_policy* p) {  // Configuration::Configuration, conf.h:20-21
    if (!p) return;
    p->force_interval = 1;
    p->kf_interval = 500;
    p->force_loss = 1;
    p->conf_loss = 0;
}

int scpr_session_create(const scpr_params* p, int device, const scpr_policy* pol, scpr_session** out) {
    if (!p || !out) return SCPR_E_PARAM;
    scpr_session* s = new scpr_session();
    s->p = *p;
    s->device = device;
    if (pol) s->pol = *pol; else scpr_policy_default(&s->pol);
    *out = s;
    return SCPR_OK;
}

void scpr_session_destroy(scpr_session* s) {
    if (!s) return;
    scpr_destroy(s->enc);
    scpr_destroy(s->dec);
    delete s;
}

// CodecInst::Compress (screenpressor.cpp:392-437): keyframe decision, quality -> loss, CompressFrame, npframes bookkeeping
int scpr_session_compress(scpr_session* s, const uint8_t* src, uint8_t* dst, int dst_cap, int host_keyframe, uint32_t quality, int* is_key) {
    if (!s || !src || !dst) return SCPR_E_PARAM;
    if (!s->enc) {
        const int r = scpr_create(&s->p, s->device, &s->enc);  // CompressBegin: sc.Init(&params), npframes = 0
        if (r < 0) return r;
        s->npframes = 0;
    }
    int ftype = 1;
    const bool forced_kf = s->pol.force_interval && (s->npframes + 1 >= s->pol.kf_interval);
    const bool host_kf = !s->pol.force_interval && host_keyframe;
    if (host_kf || forced_kf) ftype = 0;
    const int loss = s->pol.force_loss ? s->pol.conf_loss : scpr_quality_to_loss(quality);
    const int sz = scpr_compress_frame(s->enc, src, dst, dst_cap, &ftype, loss);
    if (sz < 0) return sz;
    if (!ftype) s->npframes = 0; else s->npframes++;
    if (is_key) *is_key = !ftype;   // AVIIF_KEYFRAME
    return sz;
}

// CodecInst::Decompress (screenpressor.cpp:592-620): the frame type comes from the data when it can be inferred
int scpr_session_decompress(scpr_session* s, const uint8_t* src, int src_len, uint8_t* dst, int pitch, int not_keyframe) {
    if (!s || !src || !dst || src_len <= 0) return SCPR_E_PARAM;
    if (!s->dec) {
        const int r = scpr_create(&s->p, s->device, &s->dec);
        if (r < 0) return r;
    }
    int ftype = not_keyframe ? 1 : 0;
    const int inferred = scpr_infer_frame_type(src[0], (uint32_t)src_len);
    if (inferred >= 0) ftype = inferred;
    return scpr_decompress_frame(s->dec, src, src_len, dst, pitch, ftype);
}

// ---- AVI -----------------------------------------------------------------------------------------------------------
int scpr_avi_create(const char* path, const scpr_avi_info* info, scpr_avi** out) {
    if (!path || !info || !out) return SCPR_E_PARAM;
    if (info->bits_per_pixel != 16 && info->bits_per_pixel != 24 && info->bits_per_pixel != 32) return SCPR_E_PARAM;
    FILE* f = fopen(path, "wb");
    if (!f) {
        scpr::set_error("cannot create %s", path);
        return SCPR_E_PARAM;
    }
    scpr_avi* a = new scpr_avi();
    a->f = f;
    a->writing = true;
    a->info = *info;
    a->info.frames = 0;
    const std::vector<uint8_t> h = build_header(a->info, 0, 0, 4, 0);
    a->hdr_bytes = h.size();
    a->movi_pos = h.size() - 4;
    fwrite(h.data(), 1, h.size(), f);
    *out = a;
    return SCPR_OK;
}

int scpr_avi_write_frame(scpr_avi* a, const uint8_t* data, uint32_t len, int is_key) {
    if (!a || !a->writing || (!data && len)) return SCPR_E_PARAM;
    const uint64_t pos = (uint64_t)ftell(a->f);
    if (pos + len + 16 + 16ull * (a->idx.size() + 1) >= 0xFFFF0000ull) {
        scpr::set_error("AVI 1.0 files end at 4 GB");
        return SCPR_E_DSTSIZE;
    }
    uint8_t ch[8];
    put32(ch, fcc("00dc")); put32(ch + 4, len);
    bool io = fwrite(ch, 1, 8, a->f) == 8;
    if (io && len) io = fwrite(data, 1, len, a->f) == len;
    if (io && (len & 1)) io = fputc(0, a->f) != EOF;  // chunks are word aligned
    if (!io) {
        a->io_failed = true;
        scpr::set_error("AVI write failed (disk full?)");
        return SCPR_E_PARAM;
    }
    try {
        a->idx.push_back(AviIndexEntry{pos + 8, len, is_key ? AVIIF_KEYFRAME : 0u});
    } catch (...) {
        return SCPR_E_PARAM;
    }
    if (len > a->max_chunk) a->max_chunk = len;
    return SCPR_OK;
}

static int avi_finish_write(scpr_avi* a) {
    const long at = ftell(a->f);
    if (at < 0 || a->io_failed) {
        scpr::set_error("AVI file is incomplete: an earlier write failed");
        return SCPR_E_PARAM;
    }
    const uint64_t movi_end = (uint64_t)at;
    std::vector<uint8_t> ix(8 + 16 * a->idx.size());
    put32(ix.data(), fcc("idx1")); put32(ix.data() + 4, (uint32_t)(16 * a->idx.size()));
    for (size_t i = 0; i < a->idx.size(); i++) {
        uint8_t* p = ix.data() + 8 + 16 * i;
        put32(p, fcc("00dc")); put32(p + 4, a->idx[i].flags);
        put32(p + 8, (uint32_t)(a->idx[i].off - 8 - a->movi_pos));  // offset of the chunk header relative to 'movi'
        put32(p + 12, a->idx[i].size);
    }
    if (fwrite(ix.data(), 1, ix.size(), a->f) != ix.size()) {
        scpr::set_error("AVI index write failed (disk full?)");
        return SCPR_E_PARAM;
    }
    const uint64_t total = (uint64_t)ftell(a->f);
    const std::vector<uint8_t> h = build_header(a->info, (uint32_t)a->idx.size(), a->max_chunk, (uint32_t)(movi_end - a->movi_pos),
                                                (uint32_t)(total - 8));
    if (fseek(a->f, 0, SEEK_SET) != 0 || fwrite(h.data(), 1, h.size(), a->f) != h.size() || fflush(a->f) != 0) {
        scpr::set_error("AVI header rewrite failed");
        return SCPR_E_PARAM;
    }
    return SCPR_OK;
}

int scpr_avi_close(scpr_avi* a) {
    if (!a) return SCPR_E_PARAM;
    int r = SCPR_OK;
    try {
        if (a->writing) r = avi_finish_write(a);
    } catch (...) {
        r = SCPR_E_PARAM;
    }
    if (a->f && fclose(a->f) != 0 && r == SCPR_OK && a->writing) r = SCPR_E_PARAM;
    delete a;
    return r;
}

static int avi_open_impl(const char* path, scpr_avi** out, scpr_avi_info* info);
int scpr_avi_open(const char* path, scpr_avi** out, scpr_avi_info* info) {
    try {  // no exception crosses the C ABI (a hostile file must not turn into std::bad_alloc in the caller)
        return avi_open_impl(path, out, info);
    } catch (...) {
        scpr::set_error("out of memory while reading %s", path ? path : "(null)");
        return SCPR_E_PARAM;
    }
}
static int avi_open_impl(const char* path, scpr_avi** out, scpr_avi_info* info) {
    if (!path || !out) return SCPR_E_PARAM;
    FILE* f = fopen(path, "rb");
    if (!f) {
        scpr::set_error("cannot open %s", path);
        return SCPR_E_PARAM;
    }
    scpr_avi* a = new scpr_avi();
    a->f = f;
    memset(&a->info, 0, sizeof(a->info));
    uint8_t h[12];
    bool ok = fread(h, 1, 12, f) == 12 && get32(h) == fcc("RIFF") && get32(h + 8) == fcc("AVI ");
    uint64_t idx_pos = 0, idx_len = 0, movi_end = 0, fsize = 0;
    bool have_strf = false;
    // walk the top-level chunks; descend into hdrl / strl, remember movi and idx1
    std::vector<std::pair<uint64_t, uint64_t>> todo;  // (start, end) ranges to scan
    if (ok) {
        fseek(f, 0, SEEK_END);
        fsize = (uint64_t)ftell(f);
        todo.push_back({12, fsize});
        while (!todo.empty()) {
            auto [pos, end] = todo.back();
            todo.pop_back();
            while (pos + 8 <= end) {
                uint8_t ch[12];
                fseek(f, (long)pos, SEEK_SET);
                if (fread(ch, 1, 8, f) != 8) break;
                const uint32_t id = get32(ch), sz = get32(ch + 4);
                const uint64_t body = pos + 8;
                if (id == fcc("LIST")) {
                    if (fread(ch + 8, 1, 4, f) != 4) break;
                    const uint32_t lt = get32(ch + 8);
                    if (lt == fcc("movi")) {
                        a->movi_pos = body;
                        movi_end = body + sz < fsize ? body + sz : fsize;  // sizes come from the file: never past its end
                    } else if (lt == fcc("hdrl") || lt == fcc("strl"))
                        todo.push_back({body + 4, body + sz});
                } else if (id == fcc("avih") && sz >= 40) {
                    uint8_t b[56] = {0};
                    if (fread(b, 1, sz < 56 ? sz : 56, f) < 40) break;
                    a->info.frames = get32(b + 16);
                } else if (id == fcc("strh") && sz >= 36) {
                    uint8_t b[56] = {0};
                    if (fread(b, 1, sz < 56 ? sz : 56, f) < 36) break;
                    if (get32(b) == fcc("vids") && !have_strf) {
                        a->info.fourcc = get32(b + 4);
                        a->info.fps_den = get32(b + 20);
                        a->info.fps_num = get32(b + 24);
                    }
                } else if (id == fcc("strf") && sz >= 40 && !have_strf) {
                    uint8_t b[52] = {0};
                    if (fread(b, 1, sz < 52 ? sz : 52, f) < 40) break;
                    a->info.width = get32(b + 4);
                    a->info.height = get32(b + 8);
                    a->info.bits_per_pixel = get16(b + 14);
                    a->info.fourcc = get32(b + 16);
                    if (a->info.bits_per_pixel == 16) {  // masks follow the header (BI_BITFIELDS layout), else 5-5-5
                        a->info.redmask = sz >= 52 ? get32(b + 40) : 0x7C00;
                        a->info.greenmask = sz >= 52 ? get32(b + 44) : 0x3E0;
                        a->info.bluemask = sz >= 52 ? get32(b + 48) : 0x1F;
                    }
                    have_strf = true;
                } else if (id == fcc("idx1")) {
                    idx_pos = body;
                    idx_len = body + sz <= fsize ? sz : (body < fsize ? fsize - body : 0);
                }
                pos = body + sz + (sz & 1);
            }
        }
    }
    ok = ok && have_strf && a->movi_pos;
    if (ok && idx_len >= 16) {
        std::vector<uint8_t> ix(idx_len);
        fseek(f, (long)idx_pos, SEEK_SET);
        ok = fread(ix.data(), 1, idx_len, f) == idx_len;
        // offsets are relative to the 'movi' fourcc in most files, absolute in some: test the first entry
        uint64_t base = a->movi_pos;
        if (ok) {
            uint8_t ch[4];
            fseek(f, (long)(base + get32(ix.data() + 8)), SEEK_SET);
            if (fread(ch, 1, 4, f) != 4 || get32(ch) != get32(ix.data())) base = 0;
        }
        for (size_t i = 0; ok && i + 16 <= idx_len; i += 16) {
            const uint32_t id = get32(ix.data() + i);
            if ((id >> 16) != (fcc("00dc") >> 16) && (id >> 16) != (fcc("00db") >> 16)) continue;  // video chunks of stream 0
            if ((id & 0xFFFF) != (fcc("00dc") & 0xFFFF)) continue;
            const uint64_t off = base + get32(ix.data() + i + 8) + 8;
            const uint32_t len = get32(ix.data() + i + 12);
            if (off + len > fsize) continue;  // entry points outside the file
            a->idx.push_back(AviIndexEntry{off, len, get32(ix.data() + i + 4)});
        }
    } else if (ok) {  // no index: scan the movi list; frame types are then inferred from the data by the decoder
        uint64_t pos = a->movi_pos + 4;
        while (pos + 8 <= movi_end) {
            uint8_t ch[12];
            fseek(f, (long)pos, SEEK_SET);
            if (fread(ch, 1, 8, f) != 8) break;
            const uint32_t id = get32(ch), sz = get32(ch + 4);
            if (id == fcc("LIST")) {  // 'rec ' lists group the chunks of one interleave period: their frames are inside
                pos += 12;
                continue;
            }
            if ((id == fcc("00dc") || id == fcc("00db")) && pos + 8 + sz <= fsize) a->idx.push_back(AviIndexEntry{pos + 8, sz, 0});
            pos += 8 + (uint64_t)sz + (sz & 1);
        }
    }
    if (!ok) {
        scpr::set_error("%s is not an AVI file with a video stream", path);
        fclose(f);
        delete a;
        return SCPR_E_PARAM;
    }
    a->info.frames = (uint32_t)a->idx.size();
    if (info) *info = a->info;
    *out = a;
    return SCPR_OK;
}

int64_t scpr_avi_read_frame(scpr_avi* a, uint32_t i, uint8_t* buf, size_t cap, int* is_key) {
    if (!a || a->writing || i >= a->idx.size()) return SCPR_E_PARAM;
    const AviIndexEntry& e = a->idx[i];
    if (is_key) *is_key = (e.flags & AVIIF_KEYFRAME) != 0;
    if (!buf) return e.size;
    if (cap < e.size) return SCPR_E_DSTSIZE;
    fseek(a->f, (long)e.off, SEEK_SET);
    if (fread(buf, 1, e.size, a->f) != e.size) return SCPR_E_PARAM;
    return e.size;
}

}  // extern "C"
