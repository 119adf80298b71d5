// tools/microbench.cu -- dependent-chain latencies of the instructions the decoder's chain warp lives on, measured on one
// resident warp (the situation of k_dec_chain's warp 0: nobody else on its scheduler).  Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/microbench tools/microbench.cu && /tmp/microbench
// Output goes to profiles/ (latencies in SM cycles per operation of a chain of N dependent operations).
#include <cstdint>
#include <cstdio>

#define N 4096

__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}

template <int OP>
__global__ void k_lat(uint32_t* out, long long* cyc, uint32_t seed, const uint32_t* gmem) {
    __shared__ uint32_t sm[4096];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (uint32_t)((i * 2654435761u + seed) & 4095u) * 4u;  // next byte offset
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(sm);
    uint32_t x = seed + (OP == 4 ? lane : 0);
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        if (OP == 0) x = lds32(sb + (x & 16380u));                                             // LDS, address from the loaded value
        if (OP == 1) x = __shfl_sync(0xFFFFFFFFu, x + 1, x & 31);                             // IADD + SHFL.IDX
        if (OP == 2) x = __reduce_or_sync(0xFFFFFFFFu, (lane == (int)(x & 31)) ? x + 1 : 0u);  // ISETP + SEL + REDUX.OR
        if (OP == 3) x = __ballot_sync(0xFFFFFFFFu, ((x >> lane) & 1u) != 0u) + 1u;            // SHF + ISETP + VOTE + IADD
        if (OP == 4) x = (uint32_t)__clz((int)(x | 1u)) + x;                                   // FLO + IADD
        if (OP == 5) x = (uint32_t)__popc(x) + x;                                              // POPC + IADD
        if (OP == 6) x = x * 3u + 1u;                                                          // IMAD
        if (OP == 7) x = (x ^ 0x5bd1e995u) & 0x7FFFFFFFu;                                      // LOP3
        if (OP == 8) x = __ldg(gmem + (x & 1023u));                                            // LDG, L1 hit after the first pass
        if (OP == 9) {                                                                          // ballot -> flo -> shfl (today's search tail)
            const uint32_t b = __ballot_sync(0xFFFFFFFFu, lane == (int)(x & 31));
            const int j = 31 - __clz((int)(b | 1u));
            x = __shfl_sync(0xFFFFFFFFu, x + lane + 1, j);
        }
        if (OP == 10) x = __reduce_add_sync(0xFFFFFFFFu, (lane == (int)(x & 31)) ? (x & 0xFFFF) + 1u : 0u);  // REDUX.SUM
        if (OP == 11) x = __reduce_max_sync(0xFFFFFFFFu, (lane == (int)(x & 31)) ? (x & 0xFFFF) + 1u : 0u);  // REDUX.MAX
        if (OP == 12) {                                                                         // select chain of the renormalisation
            const uint32_t x1 = (x << 8) | 0x5au, x2 = (x << 16) | 0x1234u;
            x = x < (1u << 15) ? x2 : (x < (1u << 23) ? x1 : x);
            x = (x >> 9) + 77u;
        }
        if (OP == 13) x = __funnelshift_r(x, seed, x & 31) + 1u;                                // SHF.R + IADD
        if (OP == 14) {                                                                         // two REDUX back to back, both needed
            const bool hit = lane == (int)(x & 31);
            const uint32_t a = __reduce_or_sync(0xFFFFFFFFu, hit ? x + 1 : 0u);
            const uint32_t b = __reduce_or_sync(0xFFFFFFFFu, hit ? (uint32_t)lane : 0u);
            x = a + b;
        }
        if (OP == 15) {  // match.any
            x = __match_any_sync(0xFFFFFFFFu, x & 3) + x;
        }
    }
    const long long t1 = clock64();
    if (lane == 0) {
        out[0] = x;
        cyc[0] = t1 - t0;
    }
}

// smem red followed by a dependent-address load chain: does a fire-and-forget RED delay the next LDS?
__global__ void k_red(uint32_t* out, long long* cyc, uint32_t seed) {
    __shared__ uint32_t sm[4096];
    __shared__ uint32_t cnt[4096];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) {
        sm[i] = (uint32_t)((i * 2654435761u + seed) & 4095u) * 4u;
        cnt[i] = 0;
    }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(sm), cb = (uint32_t)__cvta_generic_to_shared(cnt);
    uint32_t x = seed;
    const long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) {
        x = lds32(sb + (x & 16380u));
        if (lane == 0) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(cb + (x & 16380u)), "r"(16u) : "memory");
    }
    const long long t1 = clock64();
    if (lane == 0) {
        out[0] = x + cnt[5];
        cyc[0] = t1 - t0;
    }
}

int main() {
    uint32_t *d_out, *d_g;
    long long* d_cyc;
    cudaMalloc(&d_out, 64);
    cudaMalloc(&d_cyc, 64);
    cudaMalloc(&d_g, 4096);
    cudaMemset(d_g, 0, 4096);
    const char* names[] = {"LDS (address from loaded value)", "IADD + SHFL.IDX", "ISETP + SEL + REDUX.OR", "SHF + ISETP + VOTE + IADD", "FLO + IADD",
                           "POPC + IADD", "IMAD", "LOP3 (x2 fused)", "LDG.CONSTANT L1 hit", "VOTE + FLO + SHFL (search tail today)",
                           "ISETP + SEL + REDUX.SUM", "ISETP + SEL + REDUX.MAX", "renorm select chain + SHF + IADD", "SHF.R (funnel) + IADD",
                           "2 x REDUX.OR in parallel + IADD", "LOP + MATCH.ANY + IADD"};
#define RUN(OP)                                                                       \
    do {                                                                              \
        for (int r = 0; r < 3; r++) k_lat<OP><<<1, 64>>>(d_out, d_cyc, 12345u + r, d_g); \
        long long c;                                                                  \
        cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);                             \
        printf("%-44s %7.1f cycles per step\n", names[OP], (double)c / N);            \
    } while (0)
    RUN(0); RUN(1); RUN(2); RUN(3); RUN(4); RUN(5); RUN(6); RUN(7); RUN(8); RUN(9); RUN(10); RUN(11); RUN(12); RUN(13); RUN(14); RUN(15);
    for (int r = 0; r < 3; r++) k_red<<<1, 64>>>(d_out, d_cyc, 777u + r);
    long long c;
    cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %7.1f cycles per step\n", "LDS chain + RED.shared by lane 0 each step", (double)c / N);
    printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
