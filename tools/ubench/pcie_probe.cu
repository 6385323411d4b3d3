// pcie_probe.cu — how fast can SMs read scattered sectors of page-locked host memory, and does a
// copy-engine transfer running beside them slow them down?  (development aid for the host-buffer path)
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o pcie_probe pcie_probe.cu && ./pcie_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

// each thread reads `per` sectors; WIDTH = bytes read per request by a group of WIDTH/4 lanes
template <int WIDTH>
__global__ void gather(const float* __restrict__ host, const unsigned* __restrict__ idx, int n, int per, float* out) {
    constexpr int LANES = WIDTH / 4;              // lanes that share one request
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = t / LANES, l = t % LANES;
    float acc = 0.f;
    for (int k = 0; k < per; ++k) {
        const int r = g * per + k;
        if (r < n) acc += __ldg(host + (size_t)idx[r] * 32 + l);   // idx counts 128-byte lines
    }
    if (acc == 123.456f) out[0] = acc;
}

__global__ void dense(const float4* __restrict__ host, size_t n4, float* out) {
    float acc = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = host[i]; acc += v.x + v.y + v.z + v.w;
    }
    if (acc == 123.456f) out[0] = acc;
}

int main() {
    const size_t bytes = 480ull << 20;           // 480 MB of page-locked host memory
    float* h; CK(cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
    for (size_t i = 0; i < bytes / 4; i += 1024) h[i] = (float)i;
    float* dout; CK(cudaMalloc(&dout, 64));
    float* ddst; CK(cudaMalloc(&ddst, 64 << 20));
    const int NMAX = 1 << 20;
    std::vector<unsigned> hi(NMAX);
    uint64_t s = 0x9E3779B97F4A7C15ull;
    const unsigned nlines = (unsigned)(bytes / 128);
    for (int i = 0; i < NMAX; ++i) { s = s * 6364136223846793005ull + 1442695040888963407ull; hi[i] = (unsigned)((s >> 33) % nlines); }
    unsigned* didx; CK(cudaMalloc(&didx, NMAX * 4)); CK(cudaMemcpy(didx, hi.data(), NMAX * 4, cudaMemcpyHostToDevice));
    cudaStream_t s1, s2; CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, f0, f1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&f0)); CK(cudaEventCreate(&f1));

    auto run_gather = [&](int width, int n, int per, int threads) {
        const int lanes = width / 4;
        const long long total = ((long long)(n + per - 1) / per) * lanes;
        const int blocks = (int)((total + threads - 1) / threads);
        if (width == 32) gather<32><<<blocks, threads, 0, s1>>>(h, didx, n, per, dout);
        else if (width == 64) gather<64><<<blocks, threads, 0, s1>>>(h, didx, n, per, dout);
        else if (width == 128) gather<128><<<blocks, threads, 0, s1>>>(h, didx, n, per, dout);
        else gather<4><<<blocks, threads, 0, s1>>>(h, didx, n, per, dout);
    };
    auto time_gather = [&](int width, int n, int per, int threads) {
        run_gather(width, n, per, threads); CK(cudaStreamSynchronize(s1));
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
            CK(cudaEventRecord(e0, s1)); run_gather(width, n, per, threads); CK(cudaEventRecord(e1, s1)); CK(cudaStreamSynchronize(s1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
        }
        return best * 1e3f;
    };
    printf("scattered reads of page-locked host memory (n requests, one per 128-byte line chosen at random)\n");
    for (int width : {4, 32, 64, 128}) for (int n : {32768, 131072, 524288}) for (int per : {1, 8}) {
        const float us = time_gather(width, n, per, 256);
        printf("  width %3d B  n %7d  per-thread %d : %8.1f us  %7.1f M req/s  %6.2f GB/s\n", width, n, per, us, n / us, (double)n * width / us / 1e3);
    }
    // dense SM read and copy-engine transfer
    for (size_t mb : {2, 8, 32}) {
        const size_t b = mb << 20;
        dense<<<296, 256, 0, s1>>>((const float4*)h, b / 16, dout); CK(cudaStreamSynchronize(s1));
        CK(cudaEventRecord(e0, s1)); dense<<<296, 256, 0, s1>>>((const float4*)h + (64 << 20) / 16, b / 16, dout); CK(cudaEventRecord(e1, s1)); CK(cudaStreamSynchronize(s1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpyAsync(ddst, h, b, cudaMemcpyHostToDevice, s2)); CK(cudaStreamSynchronize(s2));
        CK(cudaEventRecord(f0, s2)); CK(cudaMemcpyAsync(ddst, h + (128 << 20) / 4, b, cudaMemcpyHostToDevice, s2)); CK(cudaEventRecord(f1, s2)); CK(cudaStreamSynchronize(s2));
        float ms2; CK(cudaEventElapsedTime(&ms2, f0, f1));
        printf("dense %2zu MB: SM 128-bit loads %7.1f us (%5.1f GB/s) | copy engine %7.1f us (%5.1f GB/s)\n", mb, ms * 1e3, b / ms / 1e6, ms2 * 1e3, b / ms2 / 1e6);
    }
    // 2-D copy of the 64 confidence rows (33.6 KB each, pitch 56 rows)
    {
        const size_t row = 8400 * 4, pitch = 56 * row;
        CK(cudaMemcpy2DAsync(ddst, row, h, pitch, row, 64, cudaMemcpyHostToDevice, s2)); CK(cudaStreamSynchronize(s2));
        CK(cudaEventRecord(f0, s2)); CK(cudaMemcpy2DAsync(ddst, row, h + (200 << 20) / 4, pitch, row, 64, cudaMemcpyHostToDevice, s2)); CK(cudaEventRecord(f1, s2)); CK(cudaStreamSynchronize(s2));
        float ms2; CK(cudaEventElapsedTime(&ms2, f0, f1));
        printf("2-D copy 64 x 33.6 KB rows (pitch 1.88 MB): %7.1f us (%5.1f GB/s)\n", ms2 * 1e3, 64 * row / ms2 / 1e6);
        const size_t row5 = 5 * row;
        CK(cudaEventRecord(f0, s2)); CK(cudaMemcpy2DAsync(ddst, row5, h + (300 << 20) / 4, pitch, row5, 64, cudaMemcpyHostToDevice, s2)); CK(cudaEventRecord(f1, s2)); CK(cudaStreamSynchronize(s2));
        CK(cudaEventElapsedTime(&ms2, f0, f1));
        printf("2-D copy 64 x 168 KB (5 rows each): %7.1f us (%5.1f GB/s)\n", ms2 * 1e3, 64 * row5 / ms2 / 1e6);
    }
    // scattered SM reads with a copy-engine transfer running beside them
    for (size_t mb : {8, 32}) {
        const int n = 524288;
        const size_t b = mb << 20;
        CK(cudaEventRecord(f0, s2));
        for (int r = 0; r < 4; ++r) CK(cudaMemcpyAsync(ddst, h + (128 << 20) / 4, b, cudaMemcpyHostToDevice, s2));
        CK(cudaEventRecord(f1, s2));
        CK(cudaEventRecord(e0, s1)); run_gather(32, n, 1, 256); CK(cudaEventRecord(e1, s1));
        CK(cudaStreamSynchronize(s1)); CK(cudaStreamSynchronize(s2));
        float ms, ms2; CK(cudaEventElapsedTime(&ms, e0, e1)); CK(cudaEventElapsedTime(&ms2, f0, f1));
        printf("concurrent: gather 32 B x %d: %7.1f us (%6.1f M req/s) | 4 copies of %zu MB: %7.1f us (%5.1f GB/s)\n", n, ms * 1e3, n / (ms * 1e3), mb, ms2 * 1e3, 4.0 * b / ms2 / 1e6);
    }
    return 0;
}
