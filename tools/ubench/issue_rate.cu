// Issue rate of a lone warp: N independent vs dependent ALU chains per loop trip; 1 vs 4 warps per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#define TRIPS 512
template <int CHAINS, int KIND> __global__ void k(float* out, long long* cyc, float seed) {
    float a[CHAINS]; unsigned u[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { a[i] = seed + i + threadIdx.x; u[i] = (unsigned)a[i]; }
    long long t0 = clock64();
#pragma unroll 1
    for (int t = 0; t < TRIPS; ++t) {
#pragma unroll
        for (int r = 0; r < 16 / CHAINS; ++r)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) {
                if (KIND == 0) a[i] = a[i] + 1.5f;                       // FADD
                if (KIND == 1) u[i] = (u[i] ^ 0x5u) + 3u;                // LOP3 + IADD
                if (KIND == 2) a[i] = fmaxf(a[i], 3.0f) + 1.0f;          // FMNMX + FADD
            }
    }
    long long t1 = clock64();
    float s = 0; unsigned us = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) { s += a[i]; us += u[i]; }
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + us;
}
template <int CHAINS, int KIND> void run(const char* name, int threads) {
    float* d; long long* c; cudaMalloc(&d, 4096 * 4); cudaMalloc(&c, 8);
    k<CHAINS, KIND><<<1, threads>>>(d, c, 1.0f); k<CHAINS, KIND><<<1, threads>>>(d, c, 2.0f);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    const int per = (KIND == 0 ? 16 : 32);
    printf("%-28s chains %2d threads %4d: %6.2f cycles per instruction (%d ALU instr + 3 loop per trip)\n", name, CHAINS, threads, (double)h / TRIPS / (per + 3), per);
    cudaFree(d); cudaFree(c);
}
int main() {
    run<1, 0>("FADD", 32); run<4, 0>("FADD", 32); run<16, 0>("FADD", 32);
    run<1, 1>("LOP3+IADD", 32); run<4, 1>("LOP3+IADD", 32); run<16, 1>("LOP3+IADD", 32);
    run<1, 2>("FMNMX+FADD", 32); run<16, 2>("FMNMX+FADD", 32);
    run<1, 0>("FADD", 128); run<16, 0>("FADD", 128); run<16, 0>("FADD", 1024);
    return 0;
}
