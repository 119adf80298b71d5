// codec.cu -- host side of the C ABI (include/scpr_c.h): codec object, frame-type policy, GOP
// ("chain") bookkeeping, workspace management and kernel orchestration.
//
// Mirrors, on the host, only the control decisions of the reference:
//   ScreenCodec::Init / CompressFrame / DecompressFrame   screencap.cpp:1565-1743
//   CScreenCapt::CompressFrame (flat / I / P choice)      screencap.cpp:1456-1518
// All pixel, model and entropy work is done by the kernels in frame_scan.cu, pframe.cu, iframe.cu,
// models.cu, rans.cu and decode.cu.  There is no CPU implementation of any of it in this library.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <vector>

#include "../../include/scpr_c.h"
#include "codec.h"

namespace scpr {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

StageTimer::StageTimer(cudaStream_t s) : on(getenv("SCPR_TIMING") != nullptr), st(s) { mark("start"); }
void StageTimer::mark(const char* name) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.push_back(e);
    names.push_back(name);
}
void StageTimer::report(const char* what) {
    if (!on) return;
    cudaEventSynchronize(ev.back());
    float tot = 0;
    cudaEventElapsedTime(&tot, ev.front(), ev.back());
    fprintf(stderr, "[scpr timing] %s total %.3f ms:", what, tot);
    for (size_t i = 1; i < ev.size(); i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ev[i - 1], ev[i]);
        fprintf(stderr, " %s=%.3f", names[i], ms);
    }
    fprintf(stderr, "\n");
    for (auto e : ev) cudaEventDestroy(e);
    ev.clear();
}

int DBuf::ensure(size_t bytes) {
    if (bytes <= cap) return SCPR_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        return SCPR_E_CUDA;
    }
    cap = want;
    return SCPR_OK;
}
void DBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

}  // namespace scpr

using namespace scpr;

#define CK(call) SCPR_CUDA_CHECK(call)
#define TRY(expr)                 \
    do {                          \
        int r__ = (expr);         \
        if (r__ < 0) return r__;  \
    } while (0)

static int ensure_states(scpr_codec* c, int n) {
    if (n <= c->n_states) return SCPR_OK;
    // keep the open chain's state: allocate a larger pool and copy the live state over
    DBuf nb;
    TRY(nb.ensure((size_t)n * model_state_bytes()));
    if (c->states.p)
        CK(cudaMemcpyAsync((uint8_t*)nb.p + (size_t)c->cur_state * model_state_bytes(),
                           (uint8_t*)c->states.p + (size_t)c->cur_state * model_state_bytes(), model_state_bytes(),
                           cudaMemcpyDeviceToDevice, c->st));
    CK(cudaStreamSynchronize(c->st));
    c->states.release();
    c->states = nb;
    c->n_states = n;
    return SCPR_OK;
}

extern "C" {

const char* scpr_last_error(void) { return g_err; }

size_t scpr_max_compressed_size(const scpr_params* p) { return (size_t)p->width * p->height * 6; }

int scpr_create(const scpr_params* p, int device, scpr_codec** out) {
    if (!p || !out) return SCPR_E_PARAM;
    *out = nullptr;
    if (p->width < 3 || p->height < 2 || p->width > 65535 || p->height > 65535) {
        set_error("unsupported frame size %ux%u", p->width, p->height);
        return SCPR_E_PARAM;
    }
    if (p->bits_per_pixel != 16 && p->bits_per_pixel != 24 && p->bits_per_pixel != 32) {
        set_error("bits_per_pixel %d: 16, 24 or 32 expected", (int)p->bits_per_pixel);  // BadVersionException(48), screencap.cpp:1607-1609
        return SCPR_E_PARAM;
    }
    if (p->bits_per_pixel == 16 && (!p->redmask || !p->greenmask || !p->bluemask)) {
        set_error("16 bpp needs the three channel masks");
        return SCPR_E_PARAM;
    }
    if (p->loss > 5) {
        set_error("loss must be 0..5 bits");
        return SCPR_E_PARAM;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        set_error("no CUDA device: libscpr_b200 has no CPU path");
        return SCPR_E_NODEVICE;
    }
    if (device < 0 || device >= ndev) return SCPR_E_PARAM;
    CK(cudaSetDevice(device));
    scpr_codec* c = new scpr_codec();
    c->p = *p;
    c->device = device;
    c->loss = (int)p->loss;
    Geo& g = c->g;
    g.X = (int)p->width;
    g.Y = (int)p->height;
    g.bpp = p->bits_per_pixel == 32 ? 4 : 3;
    g.pitch = g.bpp == 4 ? g.X * 4 : ((g.X * 3 + 3) & ~3);
    if (p->bits_per_pixel == 16) {  // ScreenCodec::Init, screencap.cpp:1575-1583
        c->rgb16 = true;
        c->m16.rmask = p->redmask; c->m16.gmask = p->greenmask; c->m16.bmask = p->bluemask;
        while (!((1u << c->m16.rshift) & p->redmask)) c->m16.rshift++;
        while (!((1u << c->m16.gshift) & p->greenmask)) c->m16.gshift++;
        while (!((1u << c->m16.bshift) & p->bluemask)) c->m16.bshift++;
    }
    g.nbx = (g.X + 15) / 16;
    g.nby = (g.Y + 15) / 16;
    g.nb = g.nbx * g.nby;
    g.frame_bytes = (size_t)g.pitch * g.Y;
    if (g.nb > 65536) {  // xx1/xx2 are two bytes each (screencap.cpp:1145-1150)
        delete c;
        set_error("more than 65536 blocks per frame");
        return SCPR_E_PARAM;
    }
    int r = c->prev.ensure(g.frame_bytes);
    if (r >= 0) r = c->mvs.ensure((size_t)g.nb * sizeof(int2));
    if (r >= 0) r = c->dec_mvs.ensure((size_t)g.nb * sizeof(int2));
    if (r < 0) {
        scpr_destroy(c);
        return r;
    }
    cudaMemset(c->prev.p, 0, g.frame_bytes);              // prev = calloc (screencap.cpp:89)
    cudaMemset(c->mvs.p, 0, (size_t)g.nb * sizeof(int2));  // mvs = calloc (screencap.cpp:96-97)
    *out = c;
    return SCPR_OK;
}

void scpr_destroy(scpr_codec* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->st);
    DBuf* all[] = {&c->prev, &c->mvs, &c->dec_mvs, &c->states, &c->frames, &c->blkinfo, &c->summary, &c->chg_list, &c->hdr,
                   &c->ftype, &c->blocks, &c->pframes, &c->runs, &c->bts_rle, &c->ihdr, &c->desc, &c->exit_tab, &c->entry,
                   &c->starts, &c->chunk_cnt, &c->frame_ev_off, &c->events, &c->intervals, &c->sorted, &c->seg_off,
                   &c->chunk_hist, &c->chunk_base, &c->chains, &c->rblks, &c->scratch, &c->out, &c->dec_ws, &c->dec_stream,
                   &c->dec_desc, &c->dec_frames, &c->dec_state, &c->dec_prev, &c->cands, &c->sorted_sym, &c->summary2, &c->raw16,
                   &c->dec24};
    for (DBuf* b : all) b->release();
    if (c->dec_progress) cudaFreeHost((void*)c->dec_progress);
    for (cudaEvent_t e : c->copy_ev) cudaEventDestroy(e);
    if (c->copy_st) cudaStreamDestroy(c->copy_st);
    if (c->aux_st) {
        cudaStreamDestroy(c->aux_st);
        cudaEventDestroy(c->aux_fork);
        cudaEventDestroy(c->aux_join);
    }
    delete c;
}

int scpr_reset(scpr_codec* c) {
    if (!c) return SCPR_E_PARAM;
    CK(cudaSetDevice(c->device));
    c->fn = 0;
    c->last_was_flat = false;
    c->have_models = false;
    c->loss = (int)c->p.loss;
    c->dec_created = false;
    c->dec_last_was_flat = false;
    CK(cudaMemsetAsync(c->prev.p, 0, c->g.frame_bytes, c->st));
    CK(cudaMemsetAsync(c->mvs.p, 0, (size_t)c->g.nb * sizeof(int2), c->st));
    if (c->dec_prev.p) CK(cudaMemsetAsync(c->dec_prev.p, 0, (size_t)c->dec_prev_pitch * c->g.Y, c->st));
    return SCPR_OK;
}

int scpr_set_stream(scpr_codec* c, void* cuda_stream) {
    if (!c) return SCPR_E_PARAM;
    c->st = (cudaStream_t)cuda_stream;
    return SCPR_OK;
}

uint64_t scpr_kernel_launches(const scpr_codec* c) { return c ? c->launches : 0; }

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// batch encoder: n device-resident frames -> host bitstreams
// ------------------------------------------------------------------------------------------------
// hooks: bit 0 = this batch opens a compress call (fire mvs_wait), bit 1 = it closes one (fire mvs_ready)
static int64_t encode_batch(scpr_codec* c, const uint8_t* d_frames, int n, const uint8_t* keyflags, uint8_t* dst,
                            size_t dst_cap, uint32_t* sizes, uint8_t* ftypes_out, int hooks = 3) {
    const Geo& g = c->g;
    cudaStream_t st = c->st;
    CK(cudaSetDevice(c->device));
    if (n <= 0) return 0;
    StageTimer tm(st);
    // host wall clock between the synchronisation points (SCPR_TIMING): what a per-frame caller pays besides the kernels
    const auto hw0 = std::chrono::steady_clock::now();
    double hw[6] = {0, 0, 0, 0, 0, 0};
    auto hmark = [&](int k) { hw[k] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - hw0).count(); };

    // ---- pass 1: frame scan + changed-block lists for every frame ------------------------------
    TRY(c->blkinfo.ensure((size_t)n * g.nb * 4));
    TRY(c->summary.ensure((size_t)n * sizeof(FrameSummary)));
    TRY(c->chg_list.ensure((size_t)n * g.nb * 4));
    TRY(c->hdr.ensure((size_t)n * sizeof(PFrameHdr)));
    TRY(c->ftype.ensure((size_t)n));
    CK(cudaMemsetAsync(c->summary.p, 0, (size_t)n * sizeof(FrameSummary), st));
    CK(cudaMemsetAsync(c->ftype.p, FT_P, (size_t)n, st));
    launch_frame_scan(d_frames, (const uint8_t*)c->prev.p, n, g, (uint32_t*)c->blkinfo.p, (FrameSummary*)c->summary.p, st,
                      &c->launches);
    if (c->loss) {
        // lossy mode: the flat test above ran on the raw frames (IsFlat precedes DoLoss); now mask the
        // non-flat frames in place (d_frames is the codec's own buffer in lossy mode) and redo the differencing on the masked pixels
        launch_apply_loss(const_cast<uint8_t*>(d_frames), n, g, (const FrameSummary*)c->summary.p, c->loss, st, &c->launches);
        TRY(c->summary2.ensure((size_t)n * sizeof(FrameSummary)));
        CK(cudaMemsetAsync(c->summary2.p, 0, (size_t)n * sizeof(FrameSummary), st));
        launch_frame_scan(d_frames, (const uint8_t*)c->prev.p, n, g, (uint32_t*)c->blkinfo.p, (FrameSummary*)c->summary2.p, st,
                          &c->launches);
    }
    tm.mark("scan");
    launch_compact_changed((const uint32_t*)c->blkinfo.p, (const uint8_t*)c->ftype.p, n, g, (uint32_t*)c->chg_list.p,
                           (PFrameHdr*)c->hdr.p, st, &c->launches);
    tm.mark("compact");
    std::vector<FrameSummary> summary(n);
    std::vector<PFrameHdr> hdr(n);
    CK(cudaMemcpyAsync(summary.data(), c->summary.p, (size_t)n * sizeof(FrameSummary), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(hdr.data(), c->hdr.p, (size_t)n * sizeof(PFrameHdr), cudaMemcpyDeviceToHost, st));
    if (c->loss) {  // "changed" comes from the masked pass, flatness and the flat colour from the raw one
        std::vector<FrameSummary> s2(n);
        CK(cudaMemcpyAsync(s2.data(), c->summary2.p, (size_t)n * sizeof(FrameSummary), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int f = 0; f < n; f++) summary[f].changed = s2[f].changed;
    }
    CK(cudaStreamSynchronize(st));
    hmark(0);

    // ---- host plan: frame types and chains (CScreenCapt::CompressFrame, screencap.cpp:1456-1518) --
    std::vector<uint8_t> ftype(n);
    std::vector<int> pframes, iframes;
    std::vector<int> chain_of(n, -1);
    std::vector<ChainDesc> chains;
    std::vector<int> chain_first(0);
    int open_chain = -1;  // index into `chains` of the chain coded P frames append to
    int next_state = 0;
    auto new_state = [&]() {
        // any slot except the one holding a live chain of this batch; chain k gets its own
        int s = next_state++;
        return s;
    };
    // the persistent state (if any) is chain "-1": it is referenced lazily by the first P frame
    const int persistent_state = c->cur_state;
    // The plan works on copies of the codec's host state; they are committed only after the last synchronisation of
    // the call succeeded (a failed call must not leave fn / last_was_flat / have_models ahead of the stream).
    unsigned l_fn = c->fn;
    bool l_last_was_flat = c->last_was_flat, l_have_models = c->have_models;
    uint8_t l_last_flat_clr[3] = {c->last_flat_clr[0], c->last_flat_clr[1], c->last_flat_clr[2]};
    for (int f = 0; f < n; f++) {
        const bool flat = !summary[f].notflat;
        const uint8_t clr[3] = {(uint8_t)summary[f].pixel0, (uint8_t)(summary[f].pixel0 >> 8), (uint8_t)(summary[f].pixel0 >> 16)};
        if (flat) {
            ftype[f] = FT_FLAT;
            if (!(l_last_was_flat && !memcmp(clr, l_last_flat_clr, 3))) {
                memcpy(l_last_flat_clr, clr, 3);
                ChainDesc cd = {0, 0, -1, 1};  // RenewI with no events of its own (screencap.cpp:1490-1494)
                chains.push_back(cd);
                chain_first.push_back(f);
                open_chain = (int)chains.size() - 1;
                l_have_models = true;
            }
            l_last_was_flat = true;
            continue;
        }
        l_last_was_flat = false;
        if (l_fn && !keyflags[f]) {
            l_fn++;
            if (!summary[f].changed) {
                ftype[f] = FT_PSAME;
                continue;
            }
            ftype[f] = FT_P;
            pframes.push_back(f);
            if (open_chain < 0) {  // continue the chain left open by the previous call
                if (!l_have_models) {
                    // reachable through scpr_import_range_state with a small (full = 0) blob whose range does not start
                    // on a frame that renews the models: there is no model state to continue
                    set_error("P frame without model state: the codec was handed a frame-range state without models (use a full blob, or cut the "
                              "range on a frame that is coded as an I frame)");
                    return SCPR_E_PARAM;
                }
                ChainDesc cd = {0, 0, -2, 0};
                chains.push_back(cd);
                chain_first.push_back(f);
                open_chain = (int)chains.size() - 1;
            }
            chain_of[f] = open_chain;
        } else {
            l_fn++;
            ftype[f] = FT_I;
            iframes.push_back(f);
            ChainDesc cd = {0, 0, -1, 1};
            chains.push_back(cd);
            chain_first.push_back(f);
            open_chain = (int)chains.size() - 1;
            chain_of[f] = open_chain;
            l_have_models = true;
        }
    }
    // From here on device state that later frames depend on is modified in place (mvs[] by the resolve, the open chain's
    // models by the replay).  If the call fails after this point the codec cannot continue the stream it was writing: the
    // guard below then makes the next coded frame an I frame with fresh models, so whatever the caller did receive stays
    // decodable (the frames of the failed call are lost to the stream, as with the reference when its caller drops a frame).
    struct Poison {
        scpr_codec* c;
        bool armed = true;
        ~Poison() {
            if (armed) {
                c->fn = 0;
                c->have_models = false;
                c->last_was_flat = false;
            }
        }
    } poison{c};
    // model states: the continued chain keeps the persistent slot, every new chain gets its own
    {
        int need = (int)chains.size() + 1;
        TRY(ensure_states(c, need));
        for (size_t k = 0; k < chains.size(); k++) {
            if (chains[k].state == -2)
                chains[k].state = persistent_state;
            else {
                int s = new_state();
                if (s == persistent_state) s = new_state();
                chains[k].state = s;
            }
        }
    }
    const int new_cur_state = chains.empty() ? c->cur_state : chains.back().state;

    // ---- stage A --------------------------------------------------------------------------------
    int total_blocks = 0;
    for (int f : pframes) {
        hdr[f].chg_off = total_blocks;
        total_blocks += hdr[f].n_changed;
    }
    const int n_p = (int)pframes.size(), n_i = (int)iframes.size();
    CK(cudaMemcpyAsync(c->ftype.p, ftype.data(), (size_t)n, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->hdr.p, hdr.data(), (size_t)n * sizeof(PFrameHdr), cudaMemcpyHostToDevice, st));
    TRY(c->frame_ev_off.ensure((size_t)(n + 1) * 4));
    PWork pw;
    memset(&pw, 0, sizeof(pw));
    pw.frames = d_frames; pw.prev0 = (const uint8_t*)c->prev.p; pw.n = n; pw.g = g;
    pw.ftype = (const uint8_t*)c->ftype.p; pw.blkinfo = (const uint32_t*)c->blkinfo.p; pw.chg_list = (const uint32_t*)c->chg_list.p;
    pw.hdr = (PFrameHdr*)c->hdr.p; pw.total_blocks = total_blocks; pw.n_pframes = n_p; pw.mvs = (int2*)c->mvs.p;
    pw.frame_ev_off = (const uint32_t*)c->frame_ev_off.p;
    pw.tm = &tm;
    if (n_p) {
        TRY(c->pframes.ensure((size_t)n_p * 4));
        TRY(c->blocks.ensure((size_t)(total_blocks + 1) * sizeof(ChgBlock)));
        TRY(c->runs.ensure((size_t)(total_blocks + 1) * 256 * 2));
        TRY(c->bts_rle.ensure((size_t)n_p * 2 * g.nb * 4));
        TRY(c->cands.ensure((size_t)n_p * (2 * MV_MAXC + 2) * 4));
        CK(cudaMemcpyAsync(c->pframes.p, pframes.data(), (size_t)n_p * 4, cudaMemcpyHostToDevice, st));
        pw.blocks = (ChgBlock*)c->blocks.p; pw.pframes = (const int*)c->pframes.p; pw.runs = (uint16_t*)c->runs.p;
        pw.bts_rle = (uint32_t*)c->bts_rle.p;
        pw.cands = (int*)c->cands.p; pw.ncands = (int*)c->cands.p + (size_t)n_p * MV_MAXC;
        pw.cands0 = (int*)c->cands.p + (size_t)n_p * (MV_MAXC + 1); pw.ncands0 = (int*)c->cands.p + (size_t)n_p * (2 * MV_MAXC + 1);
        pw.pre_resolve = (hooks & 1) ? c->mvs_wait : nullptr;
        pw.post_resolve = (hooks & 2) ? c->mvs_ready : nullptr;
        pw.hook_user = c->mvs_user;
        pw.in_hook = &c->in_hook;
        launch_p_stage_a(pw, st, &c->launches);
        tm.mark("p_stage_a");
    } else {  // no motion search in this batch: mvs[] passes through unchanged
        c->in_hook = true;
        if ((hooks & 1) && c->mvs_wait) c->mvs_wait(c->mvs_user);
        if ((hooks & 2) && c->mvs_ready) c->mvs_ready(c->mvs_user);
        c->in_hook = false;
    }
    IWork iw;
    memset(&iw, 0, sizeof(iw));
    std::vector<IFrameHdr> ihdr(n_i);
    const long total_px = (long)g.X * g.Y;
    const int nchunks = (int)((total_px - (g.X + 1) + ICHUNK - 1) / ICHUNK);
    if (n_i) {
        for (int k = 0; k < n_i; k++) {
            ihdr[k].frame = iframes[k];
            ihdr[k].n_hdr_ev = ihdr[k].n_ev = ihdr[k].pad = 0;
        }
        TRY(c->ihdr.ensure((size_t)n_i * sizeof(IFrameHdr)));
        TRY(c->desc.ensure((size_t)n_i * total_px * 2));
        TRY(c->exit_tab.ensure((size_t)n_i * nchunks * 256));
        TRY(c->entry.ensure((size_t)n_i * nchunks * 2));
        TRY(c->starts.ensure((size_t)n_i * nchunks * ICHUNK * 2));
        TRY(c->chunk_cnt.ensure((size_t)n_i * nchunks * 16));
        CK(cudaMemcpyAsync(c->ihdr.p, ihdr.data(), (size_t)n_i * sizeof(IFrameHdr), cudaMemcpyHostToDevice, st));
        iw.frames = d_frames; iw.g = g; iw.hdr = (IFrameHdr*)c->ihdr.p; iw.n_iframes = n_i; iw.nchunks = nchunks;
        iw.desc = (uint16_t*)c->desc.p; iw.exit_tab = (uint8_t*)c->exit_tab.p; iw.entry = (uint16_t*)c->entry.p;
        iw.starts = (uint16_t*)c->starts.p; iw.chunk_cnt = (uint32_t*)c->chunk_cnt.p;
        iw.frame_ev_off = (const uint32_t*)c->frame_ev_off.p;
        iw.tm = &tm;
        iw.bands = c->threads_layout;
        launch_i_stage_a(iw, st, &c->launches);
        tm.mark("i_stage_a");
    }
    if (n_p) CK(cudaMemcpyAsync(hdr.data(), c->hdr.p, (size_t)n * sizeof(PFrameHdr), cudaMemcpyDeviceToHost, st));
    if (n_i) CK(cudaMemcpyAsync(ihdr.data(), c->ihdr.p, (size_t)n_i * sizeof(IFrameHdr), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    hmark(1);

    // ---- event layout, chains, rANS blocks ----------------------------------------------------
    std::vector<uint32_t> frame_ev_off(n + 1, 0), frame_nev(n, 0);
    for (int f : pframes) frame_nev[f] = hdr[f].n_ev;
    for (int k = 0; k < n_i; k++) frame_nev[iframes[k]] = ihdr[k].n_ev;
    uint64_t tot = 0;
    for (int f = 0; f < n; f++) {
        frame_ev_off[f] = (uint32_t)tot;
        tot += frame_nev[f];
    }
    frame_ev_off[n] = (uint32_t)tot;
    if (tot >= 0xFFFF0000ull) {
        set_error("batch too large: %llu events", (unsigned long long)tot);
        return SCPR_E_PARAM;
    }
    const uint32_t total_ev = (uint32_t)tot;
    for (size_t k = 0; k < chains.size(); k++) {
        const int f0 = chain_first[k];
        const int f1 = k + 1 < chains.size() ? chain_first[k + 1] : n;
        chains[k].ev_off = frame_ev_off[f0];
        chains[k].n_ev = frame_ev_off[f1] - frame_ev_off[f0];
    }
    std::vector<RansBlk> rblks;
    uint64_t scratch_bytes = 0;
    for (int f = 0; f < n; f++)
        for (uint32_t b = 0; b < frame_nev[f]; b += RANS_BLOCK) {
            RansBlk rb;
            rb.iv_off = frame_ev_off[f] + b;
            rb.len = frame_nev[f] - b < (uint32_t)RANS_BLOCK ? frame_nev[f] - b : (uint32_t)RANS_BLOCK;
            rb.scratch = (uint32_t)scratch_bytes;
            rb.size = 0;
            rb.frame = (uint32_t)f;
            rb.out_off = 0;
            scratch_bytes += 2 * (uint64_t)rb.len + 4;
            rblks.push_back(rb);
        }
    if (scratch_bytes >= 0xFFFF0000ull) {
        set_error("batch too large: %llu scratch bytes", (unsigned long long)scratch_bytes);
        return SCPR_E_PARAM;
    }
    const int n_chains = (int)chains.size(), n_rb = (int)rblks.size();
    CK(cudaMemcpyAsync(c->frame_ev_off.p, frame_ev_off.data(), (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st));
    TRY(c->events.ensure((size_t)(total_ev + 1) * 4));
    TRY(c->intervals.ensure((size_t)(total_ev + 1) * 4));
    pw.events = (uint32_t*)c->events.p; pw.intervals = (uint32_t*)c->intervals.p;
    iw.events = (uint32_t*)c->events.p;
    tm.mark("plan2");
    if (n_p) launch_p_emit(pw, st, &c->launches);
    if (n_i) launch_i_emit(iw, st, &c->launches);
    tm.mark("emit");

    if (n_chains) {
        TRY(c->sorted.ensure((size_t)(total_ev + 1) * 4));
        TRY(c->sorted_sym.ensure((size_t)(total_ev + 1) * 2));
        TRY(c->seg_off.ensure((size_t)n_chains * (NUM_CX + 1) * 4));
        size_t hist_entries = 0;
        for (auto& cd : chains) hist_entries += replay_hist_entries(cd.n_ev);
        TRY(c->chunk_hist.ensure((hist_entries + 1) * 4));
        std::vector<uint32_t> cb(sort_chunk_words(chains.data(), n_chains) + 4);
        const size_t words = build_sort_chunks(chains.data(), n_chains, cb.data());
        TRY(c->chunk_base.ensure((words + 1) * 4));
        TRY(c->chains.ensure((size_t)n_chains * sizeof(ChainDesc)));
        CK(cudaMemcpyAsync(c->chunk_base.p, cb.data(), words * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(c->chains.p, chains.data(), (size_t)n_chains * sizeof(ChainDesc), cudaMemcpyHostToDevice, st));
        ReplayWork rw;
        memset(&rw, 0, sizeof(rw));
        rw.events = (const uint32_t*)c->events.p; rw.intervals = (uint32_t*)c->intervals.p;
        rw.chains = (const ChainDesc*)c->chains.p; rw.n_chains = n_chains; rw.h_chains = chains.data();
        rw.states = (uint8_t*)c->states.p; rw.f0 = 32;
        rw.sorted = (uint32_t*)c->sorted.p; rw.sorted_sym = (uint16_t*)c->sorted_sym.p; rw.seg_off = (uint32_t*)c->seg_off.p; rw.chunk_hist = (uint32_t*)c->chunk_hist.p;
        rw.chunk_base = (const uint32_t*)c->chunk_base.p; rw.total_events = total_ev; rw.tm = &tm;
        if (!c->aux_st) {
            CK(cudaStreamCreateWithFlags(&c->aux_st, cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&c->aux_fork, cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c->aux_join, cudaEventDisableTiming));
        }
        rw.aux = c->aux_st; rw.fork = c->aux_fork; rw.join = c->aux_join;
        launch_replay(rw, st, &c->launches);
        tm.mark("replay");
        CK(cudaStreamSynchronize(st));  // `cb` and `chains` are host vectors read by the async copies above
        hmark(2);
    }
    if (n_rb) {
        TRY(c->rblks.ensure((size_t)n_rb * sizeof(RansBlk)));
        TRY(c->scratch.ensure((size_t)scratch_bytes + 16));
        CK(cudaMemcpyAsync(c->rblks.p, rblks.data(), (size_t)n_rb * sizeof(RansBlk), cudaMemcpyHostToDevice, st));
        launch_rans((const uint32_t*)c->intervals.p, (RansBlk*)c->rblks.p, n_rb, (uint8_t*)c->scratch.p, st, &c->launches);
        tm.mark("rans");
        CK(cudaMemcpyAsync(rblks.data(), c->rblks.p, (size_t)n_rb * sizeof(RansBlk), cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(st));
    hmark(3);

    // ---- output layout -----------------------------------------------------------------------------
    std::vector<uint64_t> frame_out(n + 1, 0);
    {
        uint64_t pos = 0;
        size_t rb = 0;
        for (int f = 0; f < n; f++) {
            frame_out[f] = pos;
            uint64_t sz = 0;
            switch (ftype[f]) {
            case FT_FLAT: sz = 4; break;
            case FT_PSAME: sz = 1; break;
            default:
                sz = 1;
                while (rb < rblks.size() && rblks[rb].frame == (uint32_t)f) {
                    rblks[rb].out_off = (uint32_t)(pos + sz);
                    sz += rblks[rb].size;
                    rb++;
                }
            }
            pos += sz;
            if (sizes) sizes[f] = (uint32_t)sz;
            if (ftypes_out) ftypes_out[f] = (ftype[f] == FT_P || ftype[f] == FT_PSAME) ? 1 : 0;
        }
        frame_out[n] = pos;
        if (pos > dst_cap) {
            set_error("destination too small: need %llu bytes, have %zu", (unsigned long long)pos, dst_cap);
            return SCPR_E_DSTSIZE;
        }
        if (pos >= 0xFFFF0000ull) {
            set_error("batch output too large");
            return SCPR_E_PARAM;
        }
    }
    const uint64_t out_bytes = frame_out[n];
    if (n_rb) {
        TRY(c->out.ensure((size_t)out_bytes + 16));
        CK(cudaMemcpyAsync(c->rblks.p, rblks.data(), (size_t)n_rb * sizeof(RansBlk), cudaMemcpyHostToDevice, st));
        launch_assemble((const RansBlk*)c->rblks.p, n_rb, (const uint8_t*)c->scratch.p, (uint8_t*)c->out.p, st, &c->launches);
        CK(cudaMemcpyAsync(dst, c->out.p, (size_t)out_bytes, cudaMemcpyDeviceToHost, st));
    }
    // the last frame of the batch is the next call's previous frame (memcpy(prev, ...) of the reference)
    CK(cudaMemcpyAsync(c->prev.p, d_frames + (size_t)(n - 1) * g.frame_bytes, g.frame_bytes, cudaMemcpyDeviceToDevice, st));
    tm.mark("assemble+d2h");
    CK(cudaStreamSynchronize(st));
    hmark(4);
    static const bool host_wall = getenv("SCPR_HOSTWALL") != nullptr;  // the same without the stage events (no stage serialisation)
    if (tm.on || host_wall)
        fprintf(stderr, "[scpr timing] encode_batch host wall: scan+sync %.3f | stage A+sync %.3f | replay+sync %.3f | rans+sync %.3f | out+sync %.3f ms\n",
                hw[0], hw[1] - hw[0], hw[2] ? hw[2] - hw[1] : 0.0, hw[3] - (hw[2] ? hw[2] : hw[1]), hw[4] - hw[3]);
    tm.report("encode_batch");
    if (tm.on) mv_stats_report();
    for (int f = 0; f < n; f++) {
        uint8_t* o = dst + frame_out[f];
        switch (ftype[f]) {
        case FT_FLAT:
            o[0] = 0x31;  // 1 + (version-1)*16, screencap.cpp:1495
            o[1] = (uint8_t)summary[f].pixel0;
            o[2] = (uint8_t)(summary[f].pixel0 >> 8);
            o[3] = (uint8_t)(summary[f].pixel0 >> 16);
            break;
        case FT_PSAME: o[0] = 0; break;    // screencap.cpp:1113-1116
        case FT_P: o[0] = 1; break;        // screencap.cpp:1117
        case FT_I: o[0] = 0x32; break;     // 2 + (version-1)*16, screencap.cpp:1509
        }
    }
    // commit the host state: the stream the caller now holds and the codec agree
    poison.armed = false;
    c->fn = l_fn;
    c->last_was_flat = l_last_was_flat;
    c->have_models = l_have_models;
    memcpy(c->last_flat_clr, l_last_flat_clr, 3);
    c->cur_state = new_cur_state;
    // keep what the debug hooks need
    c->dbg_n = n;
    c->dbg_frame_ev_off = frame_ev_off;
    c->dbg_ftype = ftype;
    c->dbg_hdr = hdr;
    return (int64_t)out_bytes;
}

extern "C" {

int64_t scpr_compress_clip_dev(scpr_codec* c, const uint8_t* d_frames, int n, const uint8_t* keyflags, uint8_t* dst,
                               size_t dst_cap, uint32_t* sizes, uint8_t* ftypes) {
    if (!c || !d_frames || !keyflags || !dst || n < 0) return SCPR_E_PARAM;
    if ((reinterpret_cast<uintptr_t>(d_frames) & 15) != 0) {  // the frame scan reads 128 bits at a time
        set_error("device frames must be 16-byte aligned");
        return SCPR_E_PARAM;
    }
    if (!c->rgb16 && c->loss && n > 0) {  // the loss mask is applied in place: on a copy, the caller's frames stay as they are
        CK(cudaSetDevice(c->device));
        TRY(c->frames.ensure((size_t)n * c->g.frame_bytes));
        CK(cudaMemcpyAsync(c->frames.p, d_frames, (size_t)n * c->g.frame_bytes, cudaMemcpyDeviceToDevice, c->st));
        d_frames = (const uint8_t*)c->frames.p;
    }
    if (c->rgb16 && n > 0) {  // device frames of 2*X bytes per row -> the RGB24 image the codec works on
        CK(cudaSetDevice(c->device));
        TRY(c->frames.ensure((size_t)n * c->g.frame_bytes));
        launch_unpack16(d_frames, (uint8_t*)c->frames.p, n, c->g, c->m16, c->st, &c->launches);
        d_frames = (const uint8_t*)c->frames.p;
    }
    return encode_batch(c, d_frames, n, keyflags, dst, dst_cap, sizes, ftypes);
}

int64_t scpr_compress_clip(scpr_codec* c, const uint8_t* frames, int n, const uint8_t* keyflags, uint8_t* dst, size_t dst_cap,
                           uint32_t* sizes, uint8_t* ftypes) {
    if (!c || !frames || !keyflags || !dst || n < 0) return SCPR_E_PARAM;
    if (n == 0) return 0;
    CK(cudaSetDevice(c->device));
    const size_t fb = c->g.frame_bytes;
    const size_t in_fb = c->rgb16 ? (size_t)c->g.X * 2 * c->g.Y : fb;  // bytes per frame on the caller's side
    TRY(c->frames.ensure((size_t)n * fb));
    if (c->rgb16) TRY(c->raw16.ensure((size_t)n * in_fb));
    uint8_t* const up = c->rgb16 ? (uint8_t*)c->raw16.p : (uint8_t*)c->frames.p;
    // Host frames: the clip is encoded as a few sub-batches (any frame boundary is a valid cut -- previous frame,
    // open model chain, mvs[] and frame counter carry over exactly as between calls) so that the upload of
    // sub-batch k+1 on a copy stream overlaps the kernels of sub-batch k.  Pinned host memory makes the
    // copies asynchronous; pageable memory degrades to copy-then-compute.
    const int per = n > 48 ? (n + 7) / 8 : n;
    const int nsub = (n + per - 1) / per;
    if (!c->copy_st) CK(cudaStreamCreateWithFlags(&c->copy_st, cudaStreamNonBlocking));
    while ((int)c->copy_ev.size() < nsub) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->copy_ev.push_back(e);
    }
    for (int k = 0; k < nsub; k++) {
        const int f0 = k * per, m = n - f0 < per ? n - f0 : per;
        CK(cudaMemcpyAsync(up + (size_t)f0 * in_fb, frames + (size_t)f0 * in_fb, (size_t)m * in_fb, cudaMemcpyHostToDevice, c->copy_st));
        CK(cudaEventRecord(c->copy_ev[k], c->copy_st));
    }
    int64_t used = 0;
    for (int k = 0; k < nsub; k++) {
        const int f0 = k * per, m = n - f0 < per ? n - f0 : per;
        CK(cudaStreamWaitEvent(c->st, c->copy_ev[k], 0));
        if (c->rgb16)
            launch_unpack16(up + (size_t)f0 * in_fb, (uint8_t*)c->frames.p + (size_t)f0 * fb, m, c->g, c->m16, c->st, &c->launches);
        const int64_t r = encode_batch(c, (const uint8_t*)c->frames.p + (size_t)f0 * fb, m, keyflags + f0, dst + used, dst_cap - (size_t)used,
                                       sizes ? sizes + f0 : nullptr, ftypes ? ftypes + f0 : nullptr, (k == 0 ? 1 : 0) | (k == nsub - 1 ? 2 : 0));
        if (r < 0) {
            cudaStreamSynchronize(c->copy_st);
            return r;
        }
        used += r;
    }
    return used;
}

int scpr_compress_frame(scpr_codec* c, const uint8_t* src, uint8_t* dst, int dst_cap, int* ftype, int loss) {
    if (!c || !src || !dst || !ftype || dst_cap <= 0) return SCPR_E_PARAM;
    if (loss < 0 || loss > 5) return SCPR_E_PARAM;
    c->loss = loss;  // SetupLossMask on change, screencap.cpp:1635-1638
    const uint8_t key = *ftype ? 0 : 1;
    uint32_t size = 0;
    uint8_t ft = 0;
    const int64_t r = scpr_compress_clip(c, src, 1, &key, dst, (size_t)dst_cap, &size, &ft);
    if (r < 0) return (int)r;
    *ftype = ft;
    return (int)size;
}

// ---- encoder state hand-off between frame ranges (DESIGN.md "Multi-GPU") ---------------------
// Blob: RangeHdr, mvs[] (nb x int2), then -- full blobs only -- the previous frame and the open chain's ModelState.
struct RangeHdr {
    uint32_t magic, nb, full, fn;
    uint8_t last_was_flat, last_flat_clr[3];
    uint32_t have_models;
    uint64_t frame_bytes, state_bytes;
};
static const uint32_t RANGE_MAGIC = 0x52435053u;  // "SPCR"

size_t scpr_range_state_size(const scpr_codec* c, int full) {
    if (!c) return 0;
    size_t n = sizeof(RangeHdr) + (size_t)c->g.nb * sizeof(int2);
    if (full) n += c->g.frame_bytes + model_state_bytes();
    return n;
}

int64_t scpr_export_range_state(scpr_codec* c, uint8_t* blob, size_t cap, int full) {
    if (!c || !blob) return SCPR_E_PARAM;
    const size_t need = scpr_range_state_size(c, full);
    if (cap < need) return SCPR_E_DSTSIZE;
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->st));
    RangeHdr h;
    memset(&h, 0, sizeof(h));
    h.magic = RANGE_MAGIC;
    h.nb = (uint32_t)c->g.nb;
    h.full = full ? 1u : 0u;
    h.fn = c->fn;
    h.last_was_flat = c->last_was_flat;
    memcpy(h.last_flat_clr, c->last_flat_clr, 3);
    h.have_models = c->have_models && c->n_states > 0;
    h.frame_bytes = c->g.frame_bytes;
    h.state_bytes = model_state_bytes();
    memcpy(blob, &h, sizeof(h));
    uint8_t* p = blob + sizeof(h);
    CK(cudaMemcpy(p, c->mvs.p, (size_t)c->g.nb * sizeof(int2), cudaMemcpyDeviceToHost));
    p += (size_t)c->g.nb * sizeof(int2);
    if (full) {
        CK(cudaMemcpy(p, c->prev.p, c->g.frame_bytes, cudaMemcpyDeviceToHost));
        p += c->g.frame_bytes;
        if (h.have_models)
            CK(cudaMemcpy(p, (const uint8_t*)c->states.p + (size_t)c->cur_state * model_state_bytes(), model_state_bytes(),
                          cudaMemcpyDeviceToHost));
        else
            memset(p, 0, model_state_bytes());
    }
    return (int64_t)need;
}

int scpr_import_range_state(scpr_codec* c, const uint8_t* blob, size_t len) {
    if (!c || !blob || len < sizeof(RangeHdr)) return SCPR_E_PARAM;
    RangeHdr h;
    memcpy(&h, blob, sizeof(h));
    if (h.magic != RANGE_MAGIC || h.nb != (uint32_t)c->g.nb || len < scpr_range_state_size(c, (int)h.full) ||
        (h.full && (h.frame_bytes != c->g.frame_bytes || h.state_bytes != model_state_bytes()))) {
        set_error("range state blob does not match this codec");
        return SCPR_E_PARAM;
    }
    CK(cudaSetDevice(c->device));
    const uint8_t* p = blob + sizeof(h);
    if (c->in_hook) {
        // inside the mvs_wait hook of a running compress call: frame types are already planned and kernels that do
        // not read mvs[] are in flight -- only the vectors are taken, ordered on the stream before the resolve
        CK(cudaMemcpyAsync(c->mvs.p, p, (size_t)c->g.nb * sizeof(int2), cudaMemcpyHostToDevice, c->st));
        CK(cudaStreamSynchronize(c->st));  // the blob is the caller's memory
        return SCPR_OK;
    }
    CK(cudaStreamSynchronize(c->st));
    c->fn = h.fn;
    c->last_was_flat = h.last_was_flat != 0;
    memcpy(c->last_flat_clr, h.last_flat_clr, 3);
    CK(cudaMemcpy(c->mvs.p, p, (size_t)c->g.nb * sizeof(int2), cudaMemcpyHostToDevice));
    p += (size_t)c->g.nb * sizeof(int2);
    // a small blob carries no models: a P frame that would continue a chain is refused (encode_batch) instead of
    // being coded on whatever the state pool holds
    if (!h.full) c->have_models = false;
    if (h.full) {
        CK(cudaMemcpy(c->prev.p, p, c->g.frame_bytes, cudaMemcpyHostToDevice));
        p += c->g.frame_bytes;
        c->have_models = h.have_models != 0;
        if (h.have_models) {
            TRY(ensure_states(c, 1));
            CK(cudaMemcpy((uint8_t*)c->states.p + (size_t)c->cur_state * model_state_bytes(), p, model_state_bytes(),
                          cudaMemcpyHostToDevice));
        }
    }
    return SCPR_OK;
}

int scpr_set_mvs_hooks(scpr_codec* c, void (*wait)(void*), void (*ready)(void*), void* user) {
    if (!c) return SCPR_E_PARAM;
    c->mvs_wait = wait;
    c->mvs_ready = ready;
    c->mvs_user = user;
    return SCPR_OK;
}

int scpr_set_threads_layout(scpr_codec* c, int n_threads) {
    if (!c || n_threads < 1) return SCPR_E_PARAM;
    // the reference indexes a per-thread table that is sized by block rows (screencap.cpp:1462, 366-367): more bands than block
    // rows is undefined behaviour there; more bands than pixel rows makes no sense anywhere
    if (n_threads > c->g.nby || n_threads > c->g.Y) {
        set_error("threads layout %d exceeds the %d block rows of the frame", n_threads, c->g.nby);
        return SCPR_E_PARAM;
    }
    c->threads_layout = n_threads;
    return SCPR_OK;
}

int64_t scpr_debug_events(scpr_codec* c, int frame, uint32_t* ev, uint32_t* iv, size_t cap) {
    if (!c || frame < 0 || frame >= c->dbg_n) return SCPR_E_PARAM;
    const uint32_t off = c->dbg_frame_ev_off[frame], cnt = c->dbg_frame_ev_off[frame + 1] - off;
    const size_t m = cnt < cap ? cnt : cap;
    if (m && ev) CK(cudaMemcpy(ev, (const uint32_t*)c->events.p + off, m * 4, cudaMemcpyDeviceToHost));
    if (m && iv) CK(cudaMemcpy(iv, (const uint32_t*)c->intervals.p + off, m * 4, cudaMemcpyDeviceToHost));
    return cnt;
}

int scpr_debug_blocks(scpr_codec* c, int frame, uint8_t* bts, int32_t* sxy4, int32_t* mv2) {
    if (!c || frame < 0 || frame >= c->dbg_n) return SCPR_E_PARAM;
    const Geo& g = c->g;
    memset(bts, 0, (size_t)g.nb);
    if (c->dbg_ftype[frame] != FT_P) return 0;
    const PFrameHdr& h = c->dbg_hdr[frame];
    std::vector<ChgBlock> blocks(h.n_changed);
    if (h.n_changed)
        CK(cudaMemcpy(blocks.data(), (const ChgBlock*)c->blocks.p + h.chg_off, (size_t)h.n_changed * sizeof(ChgBlock),
                      cudaMemcpyDeviceToHost));
    for (auto& b : blocks) {
        const int by = (int)b.bi / g.nbx, bx = (int)b.bi - by * g.nbx;
        bts[b.bi] = b.bt;
        if (sxy4) {
            sxy4[4 * b.bi + 0] = bx * 16 + (int)((b.info >> 4) & 15);
            sxy4[4 * b.bi + 1] = by * 16 + (int)((b.info >> 8) & 15);
            sxy4[4 * b.bi + 2] = bx * 16 + (int)((b.info >> 12) & 15) + 1;
            sxy4[4 * b.bi + 3] = by * 16 + (int)((b.info >> 16) & 15) + 1;
        }
        if (mv2 && b.bt >= 3) {
            mv2[2 * b.bi] = b.mx;
            mv2[2 * b.bi + 1] = b.my;
        }
    }
    return h.n_changed;
}

float scpr_debug_frame_scan(scpr_codec* c, int mode, const uint8_t* d_frames, const uint8_t* d_prev, int n, int reps, uint32_t* blkinfo, uint32_t* summary) {
    if (!c || !d_frames || n <= 0 || reps <= 0 || mode < 0 || mode > 2) return (float)SCPR_E_PARAM;
    const Geo& g = c->g;
    if (cudaSetDevice(c->device) != cudaSuccess) return (float)SCPR_E_CUDA;
    if (c->blkinfo.ensure((size_t)n * g.nb * 4) < 0 || c->summary.ensure((size_t)n * sizeof(FrameSummary)) < 0)
        return (float)SCPR_E_CUDA;
    const uint8_t* prev = d_prev ? d_prev : (const uint8_t*)c->prev.p;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    bool ok = launch_frame_scan_mode(mode, d_frames, prev, n, g, (uint32_t*)c->blkinfo.p, (FrameSummary*)c->summary.p, c->st, &c->launches);  // warm-up
    cudaMemsetAsync(c->summary.p, 0, (size_t)n * sizeof(FrameSummary), c->st);
    cudaEventRecord(e0, c->st);
    for (int r = 0; ok && r < reps; r++)
        ok = launch_frame_scan_mode(mode, d_frames, prev, n, g, (uint32_t*)c->blkinfo.p, (FrameSummary*)c->summary.p, c->st, &c->launches);
    cudaEventRecord(e1, c->st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (!ok) {
        set_error("the TMA frame scan cannot take this geometry");
        return (float)SCPR_E_UNSUPPORTED;
    }
    if (cudaGetLastError() != cudaSuccess) return (float)SCPR_E_CUDA;
    if (blkinfo && cudaMemcpy(blkinfo, c->blkinfo.p, (size_t)n * g.nb * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return (float)SCPR_E_CUDA;
    if (summary && cudaMemcpy(summary, c->summary.p, (size_t)n * sizeof(FrameSummary), cudaMemcpyDeviceToHost) != cudaSuccess) return (float)SCPR_E_CUDA;
    return ms / reps;
}

float scpr_bench_frame_scan(scpr_codec* c, const uint8_t* d_frames, int n, int reps) {
    return scpr_debug_frame_scan(c, 0, d_frames, nullptr, n, reps, nullptr, nullptr);
}

}  // extern "C"
