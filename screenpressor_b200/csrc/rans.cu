// rans.cu -- stage C: one independent rANS stream per block of <= 131072 intervals.
//
// Replaces RansMTCoder::writeBlock and the rans_byte.h primitives it calls
// (reference ransmt.h:116-134; rans_byte.h:47-102): state starts at 1<<23, intervals are consumed
// in reverse, a zero frequency means "store the byte raw", the 32-bit state is flushed little
// endian in front of the block.  Blocks never share state, so each is its own stream.
//
// One warp per stream.  The state recurrence x -> C(s, x) is strictly serial, so everything that
// does not depend on x is taken off that chain: the warp stages tiles of 1024 intervals in shared
// memory with coalesced loads and computes, per interval, the exact 32-bit reciprocal of its
// frequency (Alverson-style: m = ceil(2^(31+s)/f), s = ceil(log2 f); x/f == mulhi(x, m) >> (s-1) for
// x < 2^31), so the serial lane does compare / mul-hi / multiply-add per symbol and no division.
// Output bytes are produced backwards into a shared-memory tile and flushed with coalesced stores.
#include "codec.h"

namespace scpr {

constexpr int RTILE = 1024;

__global__ void __launch_bounds__(32) k_rans_encode(const uint32_t* __restrict__ intervals, RansBlk* __restrict__ blks,
                                                    uint8_t* __restrict__ scratch) {
    __shared__ uint32_t s_iv[RTILE];
    __shared__ uint32_t s_rcp[RTILE];
    __shared__ uint8_t s_out[2 * RTILE + 8];
    const int lane = threadIdx.x;
    RansBlk& blk = blks[blockIdx.x];
    const uint32_t* iv = intervals + blk.iv_off;
    const int len = (int)blk.len;
    uint8_t* const end = scratch + blk.scratch + 2 * (size_t)len + 4;
    uint8_t* gp = end;  // global write cursor, moves down
    uint32_t x = RANS_L;
    for (int hi = len; hi > 0; hi -= RTILE) {
        const int lo = max(0, hi - RTILE), cnt = hi - lo;
        for (int i = lane; i < cnt; i += 32) {
            const uint32_t v = iv[lo + i];
            const uint32_t f = v & 0xFFFFu;
            uint32_t m = 0;
            if (f >= 2) {
                const int s = 32 - __clz(f - 1);  // ceil(log2 f)
                m = (uint32_t)((((unsigned long long)1 << (31 + s)) + f - 1) / f);
            }
            s_iv[i] = v;
            s_rcp[i] = m;
        }
        __syncwarp();
        int nout = 0;
        if (lane == 0) {
            uint8_t* p = s_out + sizeof(s_out);
            for (int i = cnt - 1; i >= 0; i--) {
                const uint32_t v = s_iv[i];
                const uint32_t freq = v & 0xFFFFu, start = v >> 16;
                if (freq) {  // RansEncPut (rans_byte.h:76-84) with RansEncRenorm (:59-71)
                    const uint32_t x_max = ((RANS_L >> PROB_BITS) << 8) * freq;
                    while (x >= x_max) {
                        *--p = (uint8_t)(x & 0xFF);
                        x >>= 8;
                    }
                    const uint32_t q = freq >= 2 ? __umulhi(x, s_rcp[i]) >> (31 - __clz(freq - 1)) : x;  // x / freq
                    x = x + start + q * ((1u << PROB_BITS) - freq);  // (q << 12) + (x - q*freq) + start
                } else
                    *--p = (uint8_t)start;  // raw byte, ransmt.h:127-128
            }
            if (lo == 0) {  // RansEncFlush, rans_byte.h:87-100
                p -= 4;
                p[0] = (uint8_t)x;
                p[1] = (uint8_t)(x >> 8);
                p[2] = (uint8_t)(x >> 16);
                p[3] = (uint8_t)(x >> 24);
            }
            nout = (int)(s_out + sizeof(s_out) - p);
        }
        nout = __shfl_sync(0xFFFFFFFFu, nout, 0);
        __syncwarp();
        gp -= nout;
        const uint8_t* src = s_out + sizeof(s_out) - nout;
        for (int i = lane; i < nout; i += 32) gp[i] = src[i];
        __syncwarp();
    }
    if (lane == 0) blk.size = (uint32_t)(end - gp);
}

// one CTA per block: copy its bytes to their place in the batch output
__global__ void __launch_bounds__(256) k_assemble(const RansBlk* __restrict__ blks, const uint8_t* __restrict__ scratch,
                                                  uint8_t* __restrict__ out) {
    const RansBlk blk = blks[blockIdx.x];
    const uint8_t* src = scratch + blk.scratch + 2 * (size_t)blk.len + 4 - blk.size;
    uint8_t* dst = out + blk.out_off;
    for (uint32_t i = threadIdx.x; i < blk.size; i += 256) dst[i] = src[i];
}

void launch_rans(const uint32_t* intervals, RansBlk* blks, int n_blks, uint8_t* scratch, cudaStream_t st, uint64_t* launches) {
    if (!n_blks) return;
    k_rans_encode<<<n_blks, 32, 0, st>>>(intervals, blks, scratch);
    ++*launches;
}

void launch_assemble(const RansBlk* blks, int n_blks, const uint8_t* scratch, uint8_t* out, cudaStream_t st, uint64_t* launches) {
    if (!n_blks) return;
    k_assemble<<<n_blks, 256, 0, st>>>(blks, scratch, out);
    ++*launches;
}

}  // namespace scpr
