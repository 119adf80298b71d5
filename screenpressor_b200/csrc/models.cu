// models.cu -- stage B: replay of the adaptive models over a GOP's events, parallel across contexts.
//
// Replaces the serial main-thread walk ec.encodeX(...) -> model.encode() of the reference
// (screencap.h:311-317, 339-344; ans_contexts.cpp:34-50; ans_contexts.h:1063-1091).  Which context a
// symbol is coded in never depends on model state, so the events of a chain (one
// GOP = everything between two RenewI) are stably partitioned by context id and each context's
// subsequence is replayed on its own:
//   * the 21 fixed tables: one warp per table, epoch-wise -- the interval table is frozen between
//     rescales and the rescale instant depends only on the event count, so an epoch is parallel
//     table look-ups + a shared-memory histogram + one table rebuild (warp scan);
//   * the 12288 colour contexts: one thread per context walking the Cx1..Cx7 state machine.
// Intervals are scattered back to bitstream order for the rANS stage.
#include <stdlib.h>

#include "codec.h"
#include "models.cuh"

namespace scpr {

constexpr int SORT_CHUNK = 8192;

size_t model_state_bytes() { return sizeof(ModelState); }
size_t replay_hist_entries(uint32_t n_ev) { return (size_t)((n_ev + SORT_CHUNK - 1) / SORT_CHUNK) * NUM_CX; }

// ---- stable partition by context id ----------------------------------------------------------------
// per chain: chunks of SORT_CHUNK events; hist[chunk][ctx] (u32) -> exclusive prefix over chunks;
// seg_off[ctx] = exclusive prefix over contexts; rank inside a chunk by in-order warp match.
struct SortChunkDesc {
    uint32_t ev_begin, len;   // batch-wide event range
    uint32_t hist_off;        // row offset (in u32 entries) of this chunk's histogram
    uint32_t chain;
};

__global__ void __launch_bounds__(256) k_sort_hist(const uint32_t* __restrict__ events, const SortChunkDesc* __restrict__ chunks,
                                                   uint32_t* __restrict__ hist) {
    // two 16-bit counters per word (a chunk holds <= 8192 events, so a half never carries over)
    __shared__ uint32_t s_h[NUM_CX / 2];
    const SortChunkDesc cd = chunks[blockIdx.x];
    for (int i = threadIdx.x; i < NUM_CX / 2; i += 256) s_h[i] = 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < cd.len; i += 256) {
        const uint32_t ctx = events[cd.ev_begin + i] >> 16;
        atomicAdd(&s_h[ctx >> 1], 1u << (16 * (ctx & 1)));
    }
    __syncthreads();
    uint32_t* row = hist + cd.hist_off;
    for (int i = threadIdx.x; i < NUM_CX; i += 256) row[i] = (s_h[i >> 1] >> (16 * (i & 1))) & 0xFFFFu;
}

// column scan: thread per (context, chain).  hist rows of a chain are contiguous.
__global__ void k_sort_scan_cols(uint32_t* __restrict__ hist, const ChainDesc* __restrict__ chains,
                                 const uint32_t* __restrict__ chain_hist_off, uint32_t* __restrict__ totals) {
    const int ctx = blockIdx.x * blockDim.x + threadIdx.x;
    const int ch = blockIdx.y;
    if (ctx >= NUM_CX) return;
    const uint32_t nchunks = (chains[ch].n_ev + SORT_CHUNK - 1) / SORT_CHUNK;
    uint32_t* col = hist + chain_hist_off[ch] + ctx;
    uint32_t run = 0;
    for (uint32_t k = 0; k < nchunks; k++) {
        const uint32_t v = col[(size_t)k * NUM_CX];
        col[(size_t)k * NUM_CX] = run;
        run += v;
    }
    totals[(size_t)ch * (NUM_CX + 1) + ctx] = run;
}

// exclusive scan over contexts -> seg_off[chain][0..NUM_CX]; one CTA per chain
__global__ void __launch_bounds__(256) k_sort_scan_ctx(uint32_t* __restrict__ seg_off) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_base;
    uint32_t* row = seg_off + (size_t)blockIdx.x * (NUM_CX + 1);
    const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < NUM_CX + 1; i0 += 256) {
        const int i = i0 + threadIdx.x;
        const uint32_t v = i < NUM_CX ? row[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) s_warp[wi] = inc;
        __syncthreads();
        uint32_t off = s_base;
        for (int j = 0; j < wi; j++) off += s_warp[j];
        if (i < NUM_CX + 1) row[i] = off + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (int j = 0; j < 8; j++) t += s_warp[j];
            s_base += t;
        }
        __syncthreads();
    }
}

// placement: one warp per chunk, events taken 32 at a time in order; equal contexts inside a
// warp step are ranked with match_any, across steps with a shared-memory running count.
__global__ void __launch_bounds__(32) k_sort_place(const uint32_t* __restrict__ events, const SortChunkDesc* __restrict__ chunks,
                                                   const uint32_t* __restrict__ hist, const uint32_t* __restrict__ seg_off,
                                                   const ChainDesc* __restrict__ chains, uint32_t* __restrict__ sorted,
                                                   uint16_t* __restrict__ sorted_sym) {
    extern __shared__ uint16_t s_cnt[];  // NUM_CX
    const SortChunkDesc cd = chunks[blockIdx.x];
    const int lane = threadIdx.x;
    for (int i = lane; i < NUM_CX; i += 32) s_cnt[i] = 0;
    __syncwarp();
    const uint32_t* row = hist + cd.hist_off;
    const uint32_t* seg = seg_off + (size_t)cd.chain * (NUM_CX + 1);
    uint32_t* out = sorted + chains[cd.chain].ev_off;
    uint16_t* out_sym = sorted_sym + chains[cd.chain].ev_off;
    for (uint32_t i0 = 0; i0 < cd.len; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool ok = i < cd.len;
        const uint32_t ev = ok ? events[cd.ev_begin + i] : 0xFFFF0000u;
        const uint32_t ctx = ev >> 16;
        const uint32_t m = __match_any_sync(0xFFFFFFFFu, ctx);
        const int leader = __ffs(m) - 1;
        uint32_t base = 0;
        if (ok && lane == leader) {
            base = s_cnt[ctx];
            s_cnt[ctx] = (uint16_t)(base + __popc(m));
        }
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (ok) {
            const uint32_t dst = seg[ctx] + row[ctx] + base + __popc(m & ((1u << lane) - 1));
            out[dst] = cd.ev_begin + i;
            out_sym[dst] = (uint16_t)ev;
        }
        __syncwarp();
    }
}

// ---- RenewI ------------------------------------------------------------------------------------------
__global__ void k_renew(ModelState* st) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < NUM_COLOR_CX) st->color[i].kind = 0;
    if (i < NUM_FIXED_CX) fixed_renew(st->fx[i], fixed_nsym(CX_NTAB + i));
}

void launch_renew_state(uint8_t* state, cudaStream_t st, uint64_t* launches) {
    k_renew<<<(NUM_COLOR_CX + 255) / 256, 256, 0, st>>>(reinterpret_cast<ModelState*>(state));
    ++*launches;
}

// ---- fixed tables: one warp per (table, chain), epoch-wise -----------------------------------------------
__global__ void __launch_bounds__(32) k_replay_fixed(ReplayWork w, const uint32_t* __restrict__ seg_off) {
    __shared__ uint32_t s_cnt[512];
    __shared__ uint16_t s_freq[512], s_cum[512];
    const int lane = threadIdx.x;
    const int t = blockIdx.x, ch = blockIdx.y;
    const ChainDesc cd = w.chains[ch];
    const uint32_t* seg = seg_off + (size_t)ch * (NUM_CX + 1);
    const uint32_t pos0 = cd.ev_off + seg[CX_NTAB + t];
    uint32_t m = seg[CX_NTAB + t + 1] - seg[CX_NTAB + t];
    if (m == 0 && !cd.renew) return;
    FixedState& fs = reinterpret_cast<ModelState*>(w.states + (size_t)cd.state * sizeof(ModelState))->fx[t];
    const int nsym = fixed_nsym(CX_NTAB + t);
    int cntsum;
    if (cd.renew) {  // FixedSizeRansCtx::renew, ans_contexts.h:1114-1131
        const int fr = PROB_SCALE / nsym, c0 = fr - (fr >> 1);
        for (int i = lane; i < nsym; i += 32) {
            s_cnt[i] = c0;
            s_freq[i] = (uint16_t)fr;
            s_cum[i] = (uint16_t)(fr * i);
        }
        cntsum = c0 * nsym;
    } else {
        for (int i = lane; i < nsym; i += 32) {
            s_cnt[i] = fs.cnt[i];
            s_freq[i] = fs.freq[i];
            s_cum[i] = fs.cum[i];
        }
        cntsum = fs.cntsum;
    }
    __syncwarp();
    const int per = (nsym + 31) / 32;  // table entries per lane at a rebuild
    uint32_t pos = pos0;
    while (m > 0) {
        // events until the rescale fires: cntsum + 16k + 16 > 4096 (ans_contexts.h:1074-1075)
        const uint32_t epoch = (uint32_t)((PROB_SCALE - 16 - cntsum) / 16 + 1);
        const uint32_t L = min(m, epoch);
        // an epoch holds at most ~130 events (cntsum >= ~2046 after a renew or rescale): issue all of the
        // lane's loads first so their latencies overlap, then look up / count
        uint32_t idx[5], sym[5];
#pragma unroll
        for (int u = 0; u < 5; u++) {
            const uint32_t i = lane + 32 * u;
            if (i < L) {
                idx[u] = w.sorted[pos + i];
                sym[u] = w.sorted_sym[pos + i];
            }
        }
#pragma unroll
        for (int u = 0; u < 5; u++) {
            const uint32_t i = lane + 32 * u;
            if (i < L) {
                w.intervals[idx[u]] = make_iv(s_freq[sym[u]], s_cum[sym[u]]);
                atomicAdd(&s_cnt[sym[u]], 16u);
            }
        }
        for (uint32_t i = lane + 160; i < L; i += 32) {  // not reached with the reference's constants
            const uint32_t id = w.sorted[pos + i], sy = w.sorted_sym[pos + i];
            w.intervals[id] = make_iv(s_freq[sy], s_cum[sy]);
            atomicAdd(&s_cnt[sy], 16u);
        }
        __syncwarp();
        cntsum += 16 * (int)L;
        pos += L;
        m -= L;
        if (cntsum + 16 > PROB_SCALE) {  // rebuild: freq = cnt, cum = prefix, cnt -= freq >> 1
            uint32_t sum = 0;
            const int b = lane * per;
            for (int j = 0; j < per; j++)
                if (b + j < nsym) sum += s_cnt[b + j];
            uint32_t inc = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d) inc += v;
            }
            uint32_t cf = inc - sum, ns = 0;
            for (int j = 0; j < per; j++)
                if (b + j < nsym) {
                    const uint32_t fr = s_cnt[b + j];
                    s_cum[b + j] = (uint16_t)cf;
                    s_freq[b + j] = (uint16_t)fr;
                    cf += fr;
                    const uint32_t nc = fr - (fr >> 1);
                    s_cnt[b + j] = nc;
                    ns += nc;
                }
            cntsum = (int)__reduce_add_sync(0xFFFFFFFFu, ns);
            __syncwarp();
        }
    }
    for (int i = lane; i < nsym; i += 32) {
        fs.cnt[i] = (uint16_t)s_cnt[i];
        fs.freq[i] = s_freq[i];
        fs.cum[i] = s_cum[i];
    }
    if (lane == 0) {
        fs.cntsum = cntsum;
        fs.nsym = nsym;
    }
}

// ---- colour contexts: one warp per (context, chain) ------------------------------------------------------
// A context's events are replayed in order, but not one at a time:
//  * kinds 4/5 (SmallContext, <= 16 symbols, every event changes the frequencies): up to 32 events per
//    step.  Inside a step the frequency of symbol k seen by event t is f_k + 50 * (#earlier events of the
//    step with symbol k), the running total is tot + 50 t, and `maxpos` only moves when some symbol
//    overtakes it -- so lanes compute their events' intervals independently from ballots / match masks,
//    and a step is cut at the first event that (a) meets a new symbol, (b) triggers the rescale
//    (tot + 100 > 4096 after its update, ans_contexts.h:214) or (c) overtakes maxpos (:212-213).
//  * kinds 6/7 (flat tables): epoch-wise like the fixed tables; kind 6 additionally tracks the number of
//    distinct symbols met (promotion to kind 7 at the 41st, ans_contexts.h:631) with match masks.
//  * everything else (raw kinds 0..3, new symbols, promotions) takes the one-event path on lane 0.
__device__ __forceinline__ int shift_for(int tot) {
    int shift = 0;
    while (tot <= PROB_SCALE / 2) {
        tot <<= 1;
        shift++;
    }
    return shift;
}

__global__ void __launch_bounds__(128) k_replay_color(ReplayWork w, const uint32_t* __restrict__ seg_off) {
    const int lane = threadIdx.x & 31;
    const int ctx = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int ch = blockIdx.y;
    const ChainDesc cd = w.chains[ch];
    const uint32_t* seg = seg_off + (size_t)ch * (NUM_CX + 1);
    const uint32_t m = seg[ctx + 1] - seg[ctx];
    ColorState& x = reinterpret_cast<ModelState*>(w.states + (size_t)cd.state * sizeof(ModelState))->color[ctx];
    if (cd.renew && lane == 0) x.kind = 0;  // Context::renew, ans_contexts.h:1050
    if (m == 0) return;
    __syncwarp();
    const uint32_t pos0 = cd.ev_off + seg[ctx];
    const uint32_t lt = (1u << lane) - 1;
    uint32_t i = 0;
    while (i < m) {
        const int kind = x.kind;
        // this step's events: lane t holds event i + t
        bool valid = i + lane < m;
        uint32_t idx = 0;
        int c = -1;
        if (valid) {
            idx = w.sorted[pos0 + i + lane];
            c = (int)(w.sorted_sym[pos0 + i + lane] & 0xFFu);
        }
        if (kind == 4 || kind == 5) {
            // The context's sorted list (<= 16 symbols, their frequencies, maxpos, the running total) stays in registers from step
            // to step and goes back to the canonical state once, when the steps end (events exhausted, or a symbol the list
            // does not hold yet): a step used to begin and end with a round trip to the ColorState in global memory.
            const int d = x.d;
            int maxpos = x.maxpos;
            const int ls = lane < d ? x.ssym[lane] : -1;            // lane k: symbol k
            int lf = lane < d ? x.sfreq[lane] : 0;                  //         and its frequency
            int tot = x.cntsum;
            bool stepped = false;
            for (;;) {
                if (kind == 4) tot = 256 - d + (int)__reduce_add_sync(0xFFFFFFFFu, (unsigned)lf);  // ans_contexts.h:303
                // position of each event's symbol
                int pos = -1;
                for (int k = 0; k < d; k++)
                    if (__shfl_sync(0xFFFFFFFFu, ls, k) == c) pos = k;
                const uint32_t bad = __ballot_sync(0xFFFFFFFFu, !valid || pos < 0);
                int B = bad ? __ffs(bad) - 1 : 32;                                   // (a) new symbol / end of events
                const int kR = tot + 100 > PROB_SCALE ? 0 : (PROB_SCALE - 100 - tot) / 50 + 1;
                B = min(B, kR + 1);                                                  // (b) rescale after event kR
                if (B <= 0) break;
                const uint32_t same = __match_any_sync(0xFFFFFFFFu, pos);
                const int cnt_same = __popc(same & lt);                          // earlier events with my symbol
                const int fpos = __shfl_sync(0xFFFFFFFFu, lf, pos < 0 ? 0 : pos);
                const int fmax = __shfl_sync(0xFFFFFFFFu, lf, maxpos);
                const uint32_t maxm = __ballot_sync(0xFFFFFFFFu, pos == maxpos);
                const int my_f = fpos + 50 * (cnt_same + 1);
                const int max_f = fmax + 50 * __popc(maxm & (lt | (1u << lane)));
                const uint32_t over = __ballot_sync(0xFFFFFFFFu, lane < B && pos != maxpos && my_f > max_f);
                int newmax = maxpos;
                if (over) {                                                      // (c) maxpos moves after that event
                    const int tO = __ffs(over) - 1;
                    B = min(B, tO + 1);
                    if (tO < B) newmax = __shfl_sync(0xFFFFFFFFu, pos, tO);
                }
                // prefix of the frequencies, and the per-event dynamic part
                int pf = lf;
#pragma unroll
                for (int dd = 1; dd < 16; dd <<= 1) {
                    const int u = __shfl_up_sync(0xFFFFFFFFu, pf, dd);
                    if (lane >= dd) pf += u;
                }
                pf -= lf;  // exclusive
                const int base = __shfl_sync(0xFFFFFFFFu, pf, pos < 0 ? 0 : pos);
                int less = 0;  // earlier events of the step whose symbol lies below mine
                uint32_t inb = 0;
                for (int k = 0; k < d; k++) {
                    const uint32_t mk = __ballot_sync(0xFFFFFFFFu, pos == k && lane < B);
                    if (k < pos) less += __popc(mk & lt);
                    if (lane == k) inb = mk;
                }
                if (lane < B) {
                    const int tt = tot + 50 * lane;
                    const int shift = shift_for(tt);
                    const int bonus = (PROB_SCALE - (tt << shift)) >> shift;
                    const int cum = (c - pos) + base + 50 * less + (maxpos < pos ? bonus : 0);
                    const int fr = fpos + 50 * cnt_same + (pos == maxpos ? bonus : 0);
                    w.intervals[idx] = make_iv((uint32_t)(fr << shift) & 0xFFFFu, (uint32_t)(cum << shift) & 0xFFFFu);
                }
                // state after the step
                int nf = lf + 50 * __popc(inb);
                tot += 50 * B;
                if (tot + 50 > PROB_SCALE) {  // rescale, ans_contexts.h:186-193
                    nf -= nf >> 1;
                    tot = 256 - d + (int)__reduce_add_sync(0xFFFFFFFFu, (unsigned)(lane < d ? nf : 0));
                }
                if (lane < d) lf = nf;
                maxpos = newmax;
                stepped = true;
                i += B;
                if (i >= m) break;
                // the next step's events
                valid = i + lane < m;
                idx = 0;
                c = -1;
                if (valid) {
                    idx = w.sorted[pos0 + i + lane];
                    c = (int)(w.sorted_sym[pos0 + i + lane] & 0xFFu);
                }
            }
            if (stepped) {
                if (lane < d) x.sfreq[lane] = (uint16_t)lf;
                if (lane == 0) {
                    x.maxpos = (uint8_t)maxpos;
                    if (kind == 5) x.cntsum = tot;
                }
                __syncwarp();
                continue;
            }
        } else if (kind == 7 || kind == 6) {
            const int step = kind == 7 ? 16 : (25 << x.fshift);
            const int cntsum = x.cntsum;
            // events until the rescale fires: cntsum + step*k + step > 4096
            const int K = (PROB_SCALE - step - cntsum) / step + 1;
            int B = min(32, min((int)(m - i), K));
            const uint32_t grp = __match_any_sync(0xFFFFFFFFu, valid ? c : -1 - lane);
            const bool leader = (grp & lt) == 0;
            const int cntc = valid ? x.cnt[c] : 1;
            const uint32_t newm = __ballot_sync(0xFFFFFFFFu, valid && kind == 6 && cntc == 0 && leader);
            const int d0 = x.d;
            // kind 6: a new symbol when 40 are already met promotes the context (not counted): cut before it
            const uint32_t promo = __ballot_sync(0xFFFFFFFFu, (newm >> lane & 1) && d0 + __popc(newm & lt) >= 40);
            if (promo) B = min(B, __ffs(promo) - 1);
            if (B > 0) {
                if (lane < B) {
                    w.intervals[idx] = make_iv(x.freq[c], x.cum[c]);
                    if (leader) {  // one writer per distinct symbol of the step
                        const int k = __popc(grp & ((B >= 32 ? 0xFFFFFFFFu : (1u << B) - 1)));
                        int nc = cntc;
                        if (kind == 6 && nc == 0) {  // placeSymbol, ans_contexts.h:632-635
                            const int fr = 1 << x.fshift;
                            nc = fr - (fr >> 1);
                        }
                        x.cnt[c] = (uint16_t)(nc + step * k);
                    }
                }
                const uint32_t inB = B >= 32 ? 0xFFFFFFFFu : (1u << B) - 1;
                const int nd = d0 + __popc(newm & inB);
                int ns = cntsum + step * B;
                __syncwarp();
                if (ns + step > PROB_SCALE) {  // rescale by the whole warp: 8 symbols per lane
                    uint32_t cn[8], sum = 0;
                    const int sh = kind == 6 ? (x.fshift > 0 ? x.fshift - 1 : 0) : 0;
                    for (int j = 0; j < 8; j++) {
                        cn[j] = x.cnt[lane * 8 + j];
                        sum += cn[j] ? cn[j] : (1u << sh);  // kind 6: unmet symbols weigh 1 << sh (ans_contexts.h:747-750)
                    }
                    uint32_t inc = sum;
#pragma unroll
                    for (int dd = 1; dd < 32; dd <<= 1) {
                        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, inc, dd);
                        if (lane >= dd) inc += u;
                    }
                    uint32_t cf = inc - sum, tots = 0;
                    for (int j = 0; j < 8; j++) {
                        const uint32_t fr = cn[j] ? cn[j] : (1u << sh);
                        x.freq[lane * 8 + j] = (uint16_t)fr;
                        x.cum[lane * 8 + j] = (uint16_t)cf;
                        cf += fr;
                        const uint32_t nc = cn[j] - (cn[j] >> 1);  // stays 0 for unmet symbols
                        x.cnt[lane * 8 + j] = (uint16_t)nc;
                        tots += nc;
                    }
                    ns = (int)__reduce_add_sync(0xFFFFFFFFu, tots);
                    if (kind == 6) {
                        const int nfs = x.fshift > 0 ? x.fshift - 1 : 0;
                        const int shft = nfs > 0 ? nfs - 1 : 0;
                        ns = (ns + ((256 - nd) << shft)) & 0xFFFF;
                        __syncwarp();
                        if (lane == 0) x.fshift = (uint8_t)nfs;
                    }
                }
                if (lane == 0) {
                    x.cntsum = kind == 6 ? (ns & 0xFFFF) : ns;
                    x.d = (uint16_t)nd;
                }
                i += B;
                __syncwarp();
                continue;
            }
        }
        // one event on lane 0: raw kinds, new symbols, promotions
        if (lane == 0) {
            uint32_t iv;
            if (x.kind < 4) {  // no statistics yet: the byte goes out raw (screencap.h:313-315)
                cc_update_raw(x, c, w.f0);
                iv = make_iv(0, (uint32_t)c);
            } else
                iv = cc_encode_counted(x, c);
            w.intervals[idx] = iv;
        }
        i += 1;
        __syncwarp();
    }
}

// host-side driver -------------------------------------------------------------------------------------
void launch_replay(const ReplayWork& w, cudaStream_t st, uint64_t* launches) {
    if (w.n_chains == 0) return;
    // chunk descriptors live at the front of chunk_base (host-built, already uploaded by the caller):
    //   chunk_base layout: [n_chunks_total SortChunkDesc][n_chains u32 chain_hist_off]
    uint32_t n_chunks = 0;
    for (int c = 0; c < w.n_chains; c++) n_chunks += (w.h_chains[c].n_ev + SORT_CHUNK - 1) / SORT_CHUNK;
    const SortChunkDesc* chunks = reinterpret_cast<const SortChunkDesc*>(w.chunk_base);
    const uint32_t* chain_hist_off = w.chunk_base + (size_t)n_chunks * (sizeof(SortChunkDesc) / 4);
    uint32_t* hist = w.chunk_hist;
    if (n_chunks) {
        k_sort_hist<<<n_chunks, 256, 0, st>>>(w.events, chunks, hist);
        ++*launches;
    }
    dim3 g1((NUM_CX + 255) / 256, w.n_chains);
    k_sort_scan_cols<<<g1, 256, 0, st>>>(hist, w.chains, chain_hist_off, w.seg_off);
    k_sort_scan_ctx<<<w.n_chains, 256, 0, st>>>(w.seg_off);
    *launches += 2;
    if (n_chunks) {
        k_sort_place<<<n_chunks, 32, NUM_CX * sizeof(uint16_t), st>>>(w.events, chunks, hist, w.seg_off, w.chains, w.sorted, w.sorted_sym);
        ++*launches;
    }
    if (w.tm) w.tm->mark("sort");
    const bool timing = w.tm && w.tm->on;
    static const bool no_aux = []() { const char* e = getenv("SCPR_AUX"); return e && e[0] == '0'; }();  // SCPR_AUX=0: one stream (A/B runs)
    if (w.aux && !timing && !no_aux) {
        // k_replay_fixed is 21 warps per chain with long serial epochs: it runs beside the colour replay instead of in front of it
        cudaEventRecord(w.fork, st);
        cudaStreamWaitEvent(w.aux, w.fork, 0);
        k_replay_fixed<<<dim3(NUM_FIXED_CX, w.n_chains), 32, 0, w.aux>>>(w, w.seg_off);
        cudaEventRecord(w.join, w.aux);
        k_replay_color<<<dim3(NUM_COLOR_CX / 4, w.n_chains), 128, 0, st>>>(w, w.seg_off);
        cudaStreamWaitEvent(st, w.join, 0);
    } else {
        k_replay_fixed<<<dim3(NUM_FIXED_CX, w.n_chains), 32, 0, st>>>(w, w.seg_off);
        if (w.tm) w.tm->mark("fixed");
        k_replay_color<<<dim3(NUM_COLOR_CX / 4, w.n_chains), 128, 0, st>>>(w, w.seg_off);
    }
    *launches += 2;
}

size_t sort_chunk_words(const ChainDesc* chains, int n_chains) {
    size_t n_chunks = 0;
    for (int c = 0; c < n_chains; c++) n_chunks += (chains[c].n_ev + SORT_CHUNK - 1) / SORT_CHUNK;
    return n_chunks * (sizeof(SortChunkDesc) / 4) + (size_t)n_chains;
}

// Builds the host image of chunk_base for launch_replay.  Returns the number of u32 words written.
size_t build_sort_chunks(const ChainDesc* chains, int n_chains, uint32_t* out) {
    size_t n_chunks = 0;
    for (int c = 0; c < n_chains; c++) n_chunks += (chains[c].n_ev + SORT_CHUNK - 1) / SORT_CHUNK;
    SortChunkDesc* cd = reinterpret_cast<SortChunkDesc*>(out);
    uint32_t* hist_off = out + n_chunks * (sizeof(SortChunkDesc) / 4);
    size_t k = 0, row = 0;
    for (int c = 0; c < n_chains; c++) {
        hist_off[c] = (uint32_t)(row * NUM_CX);
        for (uint32_t b = 0; b < chains[c].n_ev; b += SORT_CHUNK) {
            cd[k].ev_begin = chains[c].ev_off + b;
            cd[k].len = chains[c].n_ev - b < (uint32_t)SORT_CHUNK ? chains[c].n_ev - b : (uint32_t)SORT_CHUNK;
            cd[k].hist_off = (uint32_t)(row * NUM_CX);
            cd[k].chain = (uint32_t)c;
            k++;
            row++;
        }
    }
    return n_chunks * (sizeof(SortChunkDesc) / 4) + (size_t)n_chains;
}

}  // namespace scpr
