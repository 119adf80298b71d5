// rans.cu -- stage C: one independent rANS stream per block of <= 131072 intervals.
//
// Replaces RansMTCoder::writeBlock and the rans_byte.h primitives it calls
// (reference ransmt.h:116-134; rans_byte.h:47-102): state starts at 1<<23, intervals are consumed
// in reverse, a zero frequency means "store the byte raw", the 32-bit state is flushed little
// endian in front of the block.  Blocks never share state, so each is its own stream: one thread
// per block here (a P frame is one short stream; an entropy-heavy I frame is a handful of long
// ones), then k_assemble packs the streams of all frames of the batch into one output buffer.
#include "kernels.cuh"

namespace scpr {

__global__ void __launch_bounds__(64) k_rans_encode(const uint32_t* __restrict__ intervals, RansBlk* __restrict__ blks,
                                                    int n_blks, uint8_t* __restrict__ scratch) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blks) return;
    RansBlk& blk = blks[b];
    const uint32_t* iv = intervals + blk.iv_off;
    uint8_t* const end = scratch + blk.scratch + 2 * (size_t)blk.len + 4;
    uint8_t* p = end;
    uint32_t x = RANS_L;
    for (int i = (int)blk.len - 1; i >= 0; i--) {
        const uint32_t v = iv[i];
        const uint32_t freq = v & 0xFFFFu, start = v >> 16;
        if (freq) {  // RansEncPut, rans_byte.h:76-84 with RansEncRenorm :59-71
            const uint32_t x_max = ((RANS_L >> PROB_BITS) << 8) * freq;
            while (x >= x_max) {
                *--p = (uint8_t)(x & 0xFF);
                x >>= 8;
            }
            x = ((x / freq) << PROB_BITS) + (x % freq) + start;
        } else
            *--p = (uint8_t)start;  // raw byte, ransmt.h:127-128
    }
    p -= 4;  // RansEncFlush, rans_byte.h:87-100
    p[0] = (uint8_t)x;
    p[1] = (uint8_t)(x >> 8);
    p[2] = (uint8_t)(x >> 16);
    p[3] = (uint8_t)(x >> 24);
    blk.size = (uint32_t)(end - p);
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
    k_rans_encode<<<(n_blks + 63) / 64, 64, 0, st>>>(intervals, blks, n_blks, scratch);
    ++*launches;
}

void launch_assemble(const RansBlk* blks, int n_blks, const uint8_t* scratch, uint8_t* out, cudaStream_t st, uint64_t* launches) {
    if (!n_blks) return;
    k_assemble<<<n_blks, 256, 0, st>>>(blks, scratch, out);
    ++*launches;
}

}  // namespace scpr
