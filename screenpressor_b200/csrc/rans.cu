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
// x < 2^31), so the serial lane does compare / mul-hi / multiply-add per symbol and no division, and its
// renormalisation is branch-free.
// Output bytes are produced backwards into a shared-memory tile and flushed with coalesced stores.
#include "codec.h"

namespace scpr {

constexpr int RTILE = 1024;

__global__ void __launch_bounds__(32) k_rans_encode(const uint32_t* __restrict__ intervals, RansBlk* __restrict__ blks,
                                                    uint8_t* __restrict__ scratch) {
    // per interval, everything the serial lane needs that does not depend on the state, as one 128-bit operand (one shared load,
    // no field extraction on the serial path -- a lone lane pays ~4 cycles per instruction whatever it does):
    //   .x = reciprocal m;  .y = x_max = freq << 19 (0: raw byte);  .z = bias = start (+ 4095 for freq 1, see below; the byte for a raw
    //   one);  .w = (4096 - freq) | shift << 16.
    __shared__ uint4 s_op[RTILE];
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
            const uint32_t f = v & 0xFFFFu, start = v >> 16;
            uint32_t m, sh = 0, bias = start;
            if (f >= 2) {
                const int s = 32 - __clz(f - 1);  // ceil(log2 f)
                m = (uint32_t)((((unsigned long long)1 << (31 + s)) + f - 1) / f);
                sh = (uint32_t)(s - 1);
            } else {
                // freq 1: x / 1 == x.  mulhi(x, 2^32-1) == x - 1, compensated in the bias:
                // x + (start + 4095) + (x - 1) * 4095 == 4096 x + start
                m = 0xFFFFFFFFu;
                bias = start + 4095u;
            }
            s_op[i] = make_uint4(m, f << 19, f ? bias : start, (((1u << PROB_BITS) - f) & 0xFFFFu) | (sh << 16));
        }
        __syncwarp();
        int nout = 0;
        if (lane == 0) {
            int pi = (int)sizeof(s_out);  // write index into s_out, moves down
            uint4 op = s_op[cnt - 1];
#pragma unroll 4
            for (int i = cnt - 1; i >= 0; i--) {
                // the next interval's operand is fetched before this state update (it does not depend on x)
                const uint4 nx = s_op[i > 0 ? i - 1 : 0];
                const uint32_t xmax = op.y;
                if (xmax) {
                    // RansEncRenorm (rans_byte.h:59-71) without branches: the state is below 2^31 and x_max at least
                    // 2^19, so at most two bytes leave; both are stored, the cursor moves by the number that count
                    const uint32_t x8 = x >> 8;
                    const int r1 = x >= xmax, r2 = x8 >= xmax;
                    s_out[pi - 1] = (uint8_t)x;
                    s_out[pi - 2] = (uint8_t)x8;
                    pi -= r1 + r2;
                    x = r2 ? (x >> 16) : (r1 ? x8 : x);
                    // RansEncPut (rans_byte.h:76-84): (q << 12) + (x - q * freq) + start
                    const uint32_t q = __umulhi(x, op.x) >> (op.w >> 16);
                    x = x + op.z + q * (op.w & 0xFFFFu);
                } else
                    s_out[--pi] = (uint8_t)op.z;  // raw byte, ransmt.h:127-128
                op = nx;
            }
            if (lo == 0) {  // RansEncFlush, rans_byte.h:87-100
                pi -= 4;
                s_out[pi] = (uint8_t)x;
                s_out[pi + 1] = (uint8_t)(x >> 8);
                s_out[pi + 2] = (uint8_t)(x >> 16);
                s_out[pi + 3] = (uint8_t)(x >> 24);
            }
            nout = (int)sizeof(s_out) - pi;
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
