// Latency of dependent chains of warp-level primitives for a lone warp (one CTA of 32 threads).
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
template <int OP> __global__ void k(unsigned* out, long long* cyc, unsigned seed) {
    __shared__ unsigned sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = (i * 7 + 1) & 1023;
    __syncwarp();
    unsigned x = seed + threadIdx.x;
    float f = (float)x;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        if (OP == 0) x = __reduce_max_sync(0xffffffffu, x) + threadIdx.x;
        if (OP == 1) x = __shfl_sync(0xffffffffu, x, (i * 5) & 31) + 1;
        if (OP == 2) x = __ballot_sync(0xffffffffu, x & 1) + threadIdx.x;
        if (OP == 3) x = sm[x & 1023];
        if (OP == 4) { f = f * 1.0001f + 1.0f; }
        if (OP == 5) x = __reduce_min_sync(0xffffffffu, (int)x) + threadIdx.x;
        if (OP == 6) x = __reduce_or_sync(0xffffffffu, x) + threadIdx.x;
        if (OP == 7) { x = __match_any_sync(0xffffffffu, x & 3) + threadIdx.x; }
        if (OP == 8) { __syncwarp(); x += 1; }
        if (OP == 9) { sm[threadIdx.x] = x; __syncwarp(); x = sm[(threadIdx.x + 1) & 31]; }
        if (OP == 10) { x = __shfl_xor_sync(0xffffffffu, x, 16) + 1; }
        if (OP == 11) { x = atomicMax(&sm[x & 7], x) + 1; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { *cyc = t1 - t0; }
    out[threadIdx.x] = x + (unsigned)f;
}
int main() {
    unsigned* d; long long* c; cudaMalloc(&d, 256); cudaMalloc(&c, 8);
    const char* names[] = {"redux.max", "shfl.idx", "ballot", "lds chain", "fmul+fadd (no fma)", "redux.min.s32", "redux.or", "match.any", "syncwarp", "sts+syncwarp+lds", "shfl.xor", "atoms.max"};
    for (int op = 0; op < 12; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (op) {
#define C(o) case o: k<o><<<1, 32>>>(d, c, rep); break;
                C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11)
            }
            cudaDeviceSynchronize();
        }
        long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("%-22s %6.1f cycles per dependent op\n", names[op], (double)h / N);
    }
    return 0;
}
