// kernels.cuh -- launcher declarations and the small descriptor structs shared by the stages.
#pragma once
#include "scpr_dev.cuh"

namespace scpr {

struct StageTimer;

// frame types decided by the host plan (CScreenCapt::CompressFrame, reference screencap.cpp:1456-1518)
enum : uint8_t {
    FT_FLAT = 0,   // 0x31 + colour, 4 bytes
    FT_I = 1,      // 0x32 + coded stream
    FT_P = 2,      // 0x01 + coded stream
    FT_PSAME = 3   // 0x00, identical to the previous frame
};

// per-frame descriptor (device copy of the host plan + counters filled by kernels)
struct PFrameHdr {
    int n_changed;      // changed blocks in this frame (k_compact_changed)
    int xx1, xx2;       // corners of the bounding box of changed blocks, as block indices
    int chg_off;        // first slot of this frame in the batch-wide ChgBlock array (host plan)
    uint32_t n_hdr_ev;  // events before the first block: 4 XX + (BT,BN) pairs (k_p_count)
    uint32_t n_ev;      // total events of the frame
    uint32_t pad[2];
};

constexpr int MV_MAXC = 64;  // candidate vectors per P frame (k_mv_cands): slot k lives in lane k & 31, half k >> 5

// one changed 16x16 block of a P frame
struct ChgBlock {
    uint32_t bi;        // block index in the frame
    uint32_t info;      // blkinfo word (changed sub-rect)
    int16_t fmx, fmy;   // F(bi): first hit of the fixed-order motion search (k_mv_search)
    int16_t mx, my;     // final motion vector (k_mv_resolve)
    uint8_t has_f;      // F(bi) exists
    uint8_t bt;         // final block type 1..4 (reference bts[])
    uint8_t rep;        // MV equals the previously coded MV of this frame (encodeBool(true))
    uint8_t fidx;       // index of F(bi) in the frame's candidate list, 0xFF = not listed (k_mv_cands)
    uint8_t pad[3];
    uint32_t mmask;     // bit k: candidate k of the frame reproduces this block (k_mv_prematch); candidates 32..63 in mmask2
    uint32_t mmask2;
    int32_t prev_nonmv; // previous pixel-coded changed block of the frame (slot index), -1 if none
    uint32_t n_runs;    // pixel runs of a pixel-coded block
    uint32_t n_ev;      // events this block contributes
    uint32_t ev_off;    // offset of its first event inside the frame's events
};

// one I frame of the batch
struct IFrameHdr {
    int frame;          // index in the batch
    uint32_t n_hdr_ev;  // events of pixel 0 + first row + pixel (0,1)
    uint32_t n_ev;      // total events
    uint32_t pad;
};

constexpr int ICHUNK = 2048;  // pixels per I-frame classification chunk

// rANS block descriptor (one independent stream, ransmt.h:116-134)
struct RansBlk {
    uint32_t iv_off;    // first interval (batch-wide index)
    uint32_t len;       // <= RANS_BLOCK
    uint32_t scratch;   // byte offset of this block's scratch area (2*len+4 bytes)
    uint32_t size;      // bytes produced (k_rans_encode)
    uint32_t frame;     // owning frame
    uint32_t out_off;   // byte offset in the batch output (k_assemble)
};

// ---- stage A --------------------------------------------------------------------------------
void launch_frame_scan(const uint8_t* frames, const uint8_t* prev0, int n, const Geo& g, uint32_t* blkinfo,
                       FrameSummary* summary, cudaStream_t st, uint64_t* launches);
bool launch_frame_scan_mode(int mode, const uint8_t* frames, const uint8_t* prev0, int n, const Geo& g, uint32_t* blkinfo,
                            FrameSummary* summary, cudaStream_t st, uint64_t* launches);
bool frame_scan_tma_usable(const uint8_t* frames, const uint8_t* prev0, const Geo& g);
bool launch_frame_scan_tma(const uint8_t* frames, const uint8_t* prev0, int n, const Geo& g, uint32_t* blkinfo, FrameSummary* summary,
                           cudaStream_t st, uint64_t* launches);
void launch_apply_loss(uint8_t* frames, int n, const Geo& g, const FrameSummary* summary, int loss, cudaStream_t st, uint64_t* launches);
// 16 bpp <-> RGB24 (channel masks of CodecParameters; shifts = position of each mask's lowest set bit, screencap.cpp:1575-1583)
struct Rgb16 {
    uint32_t rmask, gmask, bmask;
    int rshift, gshift, bshift;
};
void launch_unpack16(const uint8_t* src16, uint8_t* dst24, int n, const Geo& g, const Rgb16& m, cudaStream_t st, uint64_t* launches);
void launch_pack16(const uint8_t* src24, uint8_t* dst16, int n, const Geo& g, int out_pitch, const Rgb16& m, cudaStream_t st, uint64_t* launches);
void launch_compact_changed(const uint32_t* blkinfo, const uint8_t* ftype, int n, const Geo& g, uint32_t* chg_list,
                            PFrameHdr* hdr, cudaStream_t st, uint64_t* launches);

// P frames: motion search, in-order resolve, pixel runs, event emission
struct PWork {
    const uint8_t* frames; const uint8_t* prev0; int n; Geo g;
    const uint8_t* ftype; const uint32_t* blkinfo; const uint32_t* chg_list;
    PFrameHdr* hdr; ChgBlock* blocks; int total_blocks;
    const int* pframes; int n_pframes;     // batch indices of P-coded frames, ascending
    int2* mvs;                             // persistent per-block MV array (reference mvs[], never cleared)
    int* cands0; int* ncands0;             // per P frame: distinct F vectors of that frame alone
    int* cands; int* ncands;               // per P frame: up to 32 distinct F vectors (packed x | y<<16) and their count
    uint16_t* runs;                        // per changed block: 256 x (ptype<<8 | n)
    uint32_t* bts_rle;                     // per P frame: scratch for the block-type RLE, 2*nb entries
    const uint32_t* frame_ev_off;          // per batch frame: first event (batch-wide index)
    uint32_t* events; uint32_t* intervals;
    struct StageTimer* tm;
    // host-side hooks around the in-order resolve (frame-range pipelining, include/scpr_c.h scpr_set_mvs_hooks)
    void (*pre_resolve)(void*); void (*post_resolve)(void*); void* hook_user; bool* in_hook;
};
void launch_p_stage_a(const PWork& w, cudaStream_t st, uint64_t* launches);   // search, resolve, runs, counts
void launch_p_emit(const PWork& w, cudaStream_t st, uint64_t* launches);      // events
void mv_stats_report();

// I frames: pixel classification, run segmentation, event emission
struct IWork {
    const uint8_t* frames; Geo g;
    IFrameHdr* hdr; int n_iframes; int nchunks;
    int bands;             // row bands that each start a new run (the reference's worker-thread count, 1 = canonical stream)
    uint16_t* desc;        // per pixel (type<<8 | len): the run that would start here
    uint8_t* exit_tab;     // per chunk: entry offset -> entry offset of the next chunk (256 entries)
    uint16_t* entry;       // per chunk: offset of the first run start
    uint16_t* starts;      // per chunk: run start offsets (ICHUNK entries)
    uint32_t* chunk_cnt;   // per chunk: {n_runs, n_ev, last_type, ev_off}
    const uint32_t* frame_ev_off;
    uint32_t* events;
    struct StageTimer* tm;
};
void launch_i_stage_a(const IWork& w, cudaStream_t st, uint64_t* launches);
void launch_i_emit(const IWork& w, cudaStream_t st, uint64_t* launches);

// ---- stage B: model replay ----------------------------------------------------------------------
struct ModelState;  // models.cu
size_t model_state_bytes();
struct ChainDesc {
    uint32_t ev_off, n_ev;  // events of the chain (batch-wide indices), all its frames back to back
    int state;              // index of the model state it starts from / leaves behind
    int renew;              // 1: RenewI before the first event (I frame or recoloured flat frame)
};
struct ReplayWork {
    const uint32_t* events; uint32_t* intervals;
    const ChainDesc* chains; int n_chains;     // device array
    const ChainDesc* h_chains;                 // host copy
    uint8_t* states;                           // n_states * model_state_bytes()
    int f0;                                    // Cx6 start frequency: 32 for v4 (screencap.cpp:1614)
    // sort workspace
    uint32_t* sorted; uint16_t* sorted_sym; uint32_t* seg_off; uint32_t* chunk_hist; const uint32_t* chunk_base;
    uint32_t total_events;
    struct StageTimer* tm;
    // optional second stream: the 21 fixed tables replay there while the colour contexts replay on the main stream (disjoint
    // contexts, disjoint interval slots); fork / join are events owned by the caller
    cudaStream_t aux; cudaEvent_t fork, join;
};
size_t replay_hist_entries(uint32_t n_ev);     // u32 entries of chunk_hist needed for a chain of n_ev events
size_t build_sort_chunks(const ChainDesc* chains, int n_chains, uint32_t* out);  // host image of chunk_base; returns words
size_t sort_chunk_words(const ChainDesc* chains, int n_chains);
void launch_replay(const ReplayWork& w, cudaStream_t st, uint64_t* launches);
void launch_renew_state(uint8_t* state, cudaStream_t st, uint64_t* launches);

// ---- stage C: rANS ---------------------------------------------------------------------------------
void launch_rans(const uint32_t* intervals, RansBlk* blks, int n_blks, uint8_t* scratch, cudaStream_t st, uint64_t* launches);
void launch_assemble(const RansBlk* blks, int n_blks, const uint8_t* scratch, uint8_t* out, cudaStream_t st, uint64_t* launches);

}  // namespace scpr
