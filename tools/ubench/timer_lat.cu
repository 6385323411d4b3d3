// timer_lat.cu — latency of a %globaltimer / %clock64 read whose value is needed at once (development aid)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__global__ void k(unsigned long long* out, int n) {
    long long c0 = clock64();
    int cnt = 0;
    for (int i = 0; i < n;) { const unsigned long long x = gt(); i += (x != 0ull) ? 1 : 2; cnt += (int)(x & 1ull); }   // loop control needs x
    long long c1 = clock64();
    for (int i = 0; i < n;) { const unsigned long long x = (unsigned long long)clock64(); i += (x != 0ull) ? 1 : 2; cnt += (int)(x & 1ull); }
    long long c2 = clock64();
    unsigned long long g0 = gt(), g1 = g0; int spins = 0;
    while (g1 == g0) { g1 = gt(); ++spins; }            // tick size of the timer
    unsigned long long g2 = g1; while (g2 == g1) g2 = gt();
    if (threadIdx.x == 0) { out[0] = (c1 - c0) / n; out[1] = (c2 - c1) / n; out[2] = cnt; out[3] = g2 - g1; }
}
int main() {
    unsigned long long* d; cudaMalloc(&d, 64);
    for (int threads : {1, 32, 1024}) {
        k<<<1, threads>>>(d, 1000); cudaDeviceSynchronize();
        unsigned long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
        printf("threads %4d: dependent globaltimer read %llu cycles, dependent clock64 read %llu cycles, globaltimer tick %llu ns\n", threads, h[0], h[1], h[3]);
    }
    return 0;
}
