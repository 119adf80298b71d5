// codec.h -- the codec object behind the C ABI (host side only).
#pragma once
#include <vector>

#include "../../include/scpr_c.h"
#include "kernels.cuh"

namespace scpr {

// optional per-stage timing (SCPR_TIMING=1): CUDA events on the codec's stream, printed to stderr
struct StageTimer {
    bool on;
    cudaStream_t st;
    std::vector<cudaEvent_t> ev;
    std::vector<const char*> names;
    explicit StageTimer(cudaStream_t s);
    void mark(const char* name);
    void report(const char* what);
};

// grow-only device buffer
struct DBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes);
    void release();
};

}  // namespace scpr

struct scpr_codec {
    scpr_params p;
    int device = 0;
    cudaStream_t st = 0;
    cudaStream_t copy_st = nullptr;         // uploads of host frames, overlapped with the kernels (scpr_compress_clip)
    std::vector<cudaEvent_t> copy_ev;
    cudaStream_t aux_st = nullptr;          // second lane for independent kernels of one call (fixed-table replay beside the colour replay)
    cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
    scpr::Geo g;                     // geometry of the resident frames (16 bpp clients: the RGB24 image the codec works on)
    bool rgb16 = false;              // the caller's frames are 16 bpp (converted on the device on the way in and out)
    scpr::Rgb16 m16 = {0, 0, 0, 0, 0, 0};
    scpr::DBuf raw16, dec24;         // 16 bpp staging: uploaded input frames / decoded RGB24 frames
    uint64_t launches = 0;

    // ---- encoder state carried between calls (CScreenCapt members, screencap.h:445-463) ----------
    int loss = 0;                    // bits of loss applied to non-flat frames (SetupLossMask, screencap.cpp:127-139)
    unsigned fn = 0;                 // coded (non-flat) frames so far
    bool last_was_flat = false;
    uint8_t last_flat_clr[3] = {0, 0, 0};
    bool have_models = false;
    scpr::DBuf prev;                 // previous frame, native pixel format
    scpr::DBuf mvs;                  // int2 per block, never cleared
    scpr::DBuf states;               // pool of ModelState; states[cur_state] belongs to the open chain
    int n_states = 0, cur_state = 0;

    // frame-range pipelining (scpr_set_mvs_hooks): called on the compress call's thread around the in-order MV resolve
    void (*mvs_wait)(void*) = nullptr;
    void (*mvs_ready)(void*) = nullptr;
    void* mvs_user = nullptr;
    bool in_hook = false;
    int threads_layout = 1;          // I-frame run breaks of the reference running with this many worker threads (scpr_set_threads_layout)

    // ---- encoder workspaces ---------------------------------------------------------------------
    scpr::DBuf summary2;
    scpr::DBuf frames, blkinfo, summary, chg_list, hdr, ftype, blocks, pframes, runs, bts_rle, cands;
    scpr::DBuf ihdr, desc, exit_tab, entry, starts, chunk_cnt;
    scpr::DBuf frame_ev_off, events, intervals, sorted, sorted_sym, seg_off, chunk_hist, chunk_base, chains, rblks, scratch, out;

    // ---- decoder state and workspaces -----------------------------------------------------------
    bool dec_created = false;        // a codec exists (first I frame seen), screencap.cpp:1698-1702
    int dec_version = 4;
    scpr::DBuf dec_state;            // pool of ModelState; dec_state[dec_cur_state] belongs to the open chain
    int dec_n_states = 0, dec_cur_state = 0;
    bool dec_last_was_flat = false;
    uint8_t dec_last_flat_clr[3] = {0, 0, 0};
    scpr::DBuf dec_prev;             // last decoded frame, output format
    int dec_prev_pitch = 0;
    scpr::DBuf dec_mvs, dec_ws, dec_stream, dec_desc, dec_frames;
    volatile int* dec_progress = nullptr;   // mapped host memory: per chain, frames below this index are complete
    int dec_progress_cap = 0;

    // ---- debug hooks -----------------------------------------------------------------------------
    int dbg_n = 0;
    std::vector<uint32_t> dbg_frame_ev_off;
    std::vector<uint8_t> dbg_ftype;
    std::vector<scpr::PFrameHdr> dbg_hdr;
};
