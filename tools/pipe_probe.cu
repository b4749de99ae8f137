// Pipe microbenchmark (nvcc -arch=sm_100a tools/pipe_probe.cu; cited by DESIGN.md section 4): can IMAD (FMA pipe) co-issue with SHF/LOP3 (ALU pipe) on sm_100?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CHAINS 8
#define STEPS 4096
__device__ __forceinline__ uint32_t rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }
__device__ __forceinline__ uint32_t mulhi(uint32_t a, uint32_t m) { uint32_t d; asm("mul.hi.u32 %0,%1,%2;" : "=r"(d) : "r"(a), "r"(m)); return d; }
__device__ __forceinline__ uint32_t madhi(uint32_t a, uint32_t m, uint32_t c) { uint32_t d; asm("mad.hi.u32 %0,%1,%2,%3;" : "=r"(d) : "r"(a), "r"(m), "r"(c)); return d; }
__device__ __forceinline__ uint32_t imad(uint32_t a, uint32_t one, uint32_t b) { uint32_t d; asm("mad.lo.u32 %0,%1,%2,%3;" : "=r"(d) : "r"(a), "r"(one), "r"(b)); return d; }

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t *sink, uint32_t seed, uint32_t one, uint32_t m7, uint32_t m29) {
    uint32_t x[CHAINS], y[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { x[c] = seed + threadIdx.x * 2654435761u + c; y[c] = seed ^ (blockIdx.x + 0x9e3779b9u * c); }
#pragma unroll 4
    for (int it = 0; it < STEPS; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            if (MODE == 0) { // 3 SHF + 1 LOP3 + 1 IADD3  (5 ALU)
                uint32_t r = rotr(x[c], 6) ^ rotr(x[c], 11) ^ rotr(x[c], 25);
                x[c] = y[c]; y[c] = y[c] + r + 0x428a2f98u;
            } else if (MODE == 1) { // 3 SHF + 1 LOP3 (4 ALU) + 2 IMAD
                uint32_t r = rotr(x[c], 6) ^ rotr(x[c], 11) ^ rotr(x[c], 25);
                x[c] = y[c]; y[c] = imad(imad(y[c], one, r), one, 0x428a2f98u);
            } else if (MODE == 2) { // 5 IMAD only
                uint32_t r = imad(x[c], one, y[c]); r = imad(r, one, x[c]); r = imad(r, one, 7u); r = imad(r, one, y[c]);
                x[c] = y[c]; y[c] = imad(r, one, 0x428a2f98u);
            } else if (MODE == 3) { // 2 SHF + 1 LOP3 (3 ALU) + 3 IMAD
                uint32_t r = rotr(x[c], 6) ^ rotr(x[c], 11) ^ y[c];
                x[c] = y[c]; y[c] = imad(imad(imad(y[c], one, r), one, 0x428a2f98u), one, x[c]);
            } else if (MODE == 5) { // 5 IMAD.HI only
                uint32_t r = mulhi(x[c], m29); r = madhi(r, m7, x[c]); r = madhi(r, m29, y[c]); r = madhi(r, m7, x[c]);
                x[c] = y[c]; y[c] = madhi(r, m29, y[c]);
            } else if (MODE == 6) { // rotr25 = x*2^7 + hi(x*2^7): 2 SHF + 1 LOP3 (3 ALU) + IMAD.HI + 2 IMAD
                uint32_t r25 = imad(x[c], m7, mulhi(x[c], m7));
                uint32_t r = rotr(x[c], 6) ^ rotr(x[c], 11) ^ r25;
                x[c] = y[c]; y[c] = imad(y[c], one, r);
            } else if (MODE == 7) { // sig0 with the plain shift as IMAD.HI: 2 SHF + LOP3 (3 ALU) + IMAD.HI + IMAD
                uint32_t r = rotr(x[c], 7) ^ rotr(x[c], 18) ^ mulhi(x[c], m29);
                x[c] = y[c]; y[c] = imad(y[c], one, r);
            } else if (MODE == 4) { // 4 ALU + 1 IMAD
                uint32_t r = rotr(x[c], 6) ^ rotr(x[c], 11) ^ rotr(x[c], 25);
                x[c] = y[c]; y[c] = imad(y[c], one, r);
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc ^= x[c] + y[c];
    if (acc == 0x12345678u) sink[0] = acc;
}
template <int MODE> void run(const char *name, double alu, double fma, uint32_t *sink) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e9;
    for (int r = 0; r < 5; r++) { cudaEventRecord(a); k<MODE><<<148 * 8, 256>>>(sink, 1, 1, 1u << 7, 1u << 29); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (r) best = ms < best ? ms : best; }
    double lanes = 148.0 * 8 * 256 * STEPS * CHAINS;
    printf("%-28s %.3f ms  ALU %.2f T/s  FMA %.2f T/s  total %.2f T instr-lanes/s\n", name, best, lanes * alu / best / 1e9, lanes * fma / best / 1e9, lanes * (alu + fma) / best / 1e9);
}
int main() {
    uint32_t *sink; cudaMalloc(&sink, 256);
    run<0>("5 ALU", 5, 0, sink);
    run<4>("4 ALU + 1 IMAD", 4, 1, sink);
    run<1>("4 ALU + 2 IMAD", 4, 2, sink);
    run<3>("3 ALU + 3 IMAD", 3, 3, sink);
    run<2>("5 IMAD", 0, 5, sink);
    run<5>("5 IMAD.HI", 0, 5, sink);
    run<6>("3 ALU + 1 IMAD.HI + 2 IMAD", 3, 3, sink);
    run<7>("3 ALU + 1 IMAD.HI + 1 IMAD", 3, 2, sink);
    return 0;
}
