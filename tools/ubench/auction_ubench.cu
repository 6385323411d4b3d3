// Cycles per iteration of the single-warp auction solves on "stuck" problems (more active rows than columns):
// 20 regular rows + E extra rows that keep evicting each other until the iteration limit.
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../../yolo-pose-cpp_b200/csrc/auction.cuh"
using namespace pb;
__global__ void k(const float* cost, int R, int C, int na, int variant, int* row_out, int* col_out, long long* cyc, unsigned long long* tele_out) {
    extern __shared__ __align__(16) unsigned char sm[];
    float* cost_s = (float*)sm;                 // [R*C]
    float* cc = cost_s + R * C;                 // [32*C]
    float* price = cc + 32 * C;
    int* owner = (int*)(price + C);
    unsigned* colbid = (unsigned*)(owner + C);
    int* colrow = (int*)(colbid + C);
    int* row = colrow + C;
    int* col = row + R;
    int* act_list = col + C;
    unsigned long long* tele = (unsigned long long*)(act_list + R + (R & 1));
    for (int i = threadIdx.x; i < R * C; i += 32) cost_s[i] = cost[i];
    for (int i = threadIdx.x; i < na; i += 32) act_list[i] = i * (R / na);        // spread over the slots
    for (int i = threadIdx.x; i < 20; i += 32) tele[i] = 0;
    __syncwarp();
    for (int i = threadIdx.x; i < na * C; i += 32) cc[i] = cost_s[act_list[i / C] * C + i % C];
    __syncwarp();
    long long t0 = clock64();
    if (variant == 0) auction_solve_hybrid32(cost_s, R, C, act_list, na, row, col, price, owner, colbid, colrow);
    else if (C <= 32) auction_solve_lean32<1>(cc, R, C, act_list, na, row, col, price, owner, colbid, colrow);
    else auction_solve_lean32<2>(cc, R, C, act_list, na, row, col, price, owner, colbid, colrow);
    long long t1 = clock64();
    __syncwarp();
    if (threadIdx.x == 0) *cyc = t1 - t0;
    for (int i = threadIdx.x; i < R; i += 32) row_out[i] = row[i];
    for (int i = threadIdx.x; i < C; i += 32) col_out[i] = col[i];
    for (int i = threadIdx.x; i < 20; i += 32) tele_out[i] = tele[i];
}
int main() {
    const int R = 128;
    for (int C : {20}) for (int extra : {1, 2, 3}) {
        const int na = C + extra;
        std::vector<float> cost((size_t)R * C, 1e9f);
        const int step = R / na;
        for (int i = 0; i < C; ++i) {
            float* r = &cost[(size_t)(i * step) * C];
            r[i] = 0.1f + 0.01f * i; if (i > 0) r[i - 1] = 0.6f + 0.003f * i; if (i + 1 < C) r[i + 1] = 0.62f + 0.002f * i;
        }
        for (int e = 0; e < extra; ++e) { float* r = &cost[(size_t)((C + e) * step) * C]; r[3 + 5 * e] = 0.5f; r[4 + 5 * e] = 0.55f; }
        float* d; int *dr, *dc; long long* dcy; unsigned long long* dt;
        cudaMalloc(&d, cost.size() * 4); cudaMalloc(&dr, R * 4); cudaMalloc(&dc, C * 4); cudaMalloc(&dcy, 8); cudaMalloc(&dt, 160);
        cudaMemcpy(d, cost.data(), cost.size() * 4, cudaMemcpyHostToDevice);
        const size_t smem = (size_t)(R * C + 32 * C + 6 * C + 2 * R + 2) * 4 + 160;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        std::vector<int> ref;
        for (int variant = 0; variant < 2; ++variant) {   // 0 hybrid (generic C), 1 lean (C <= 64)
            long long h = 0;
            for (int rep = 0; rep < 3; ++rep) { k<<<1, 32, smem>>>(d, R, C, na, variant, dr, dc, dcy, dt); cudaDeviceSynchronize(); }
            cudaMemcpy(&h, dcy, 8, cudaMemcpyDeviceToHost);
            std::vector<int> col(C); cudaMemcpy(col.data(), dc, C * 4, cudaMemcpyDeviceToHost);
            unsigned long long t[20]; cudaMemcpy(t, dt, 160, cudaMemcpyDeviceToHost);
            if (variant == 0) ref = col;
            printf("C %2d extra %d variant %d: %7lld cycles per solve (%5.0f per iteration of 50)  %s", C, extra, variant, h, h / 50.0, col == ref ? "same" : "DIFFERENT");
            if (variant == 3) printf("  iters nb1 %llu nb2 %llu nb>2 %llu loop %llu | eval %llu apply %llu redux-or %llu", t[15] / 1000, t[16] / 1000, t[17] / 1000, t[18], t[0], t[1], t[2]);
#ifdef LEAN_PROFILE
            if (variant == 1) { unsigned long long g[8]; cudaMemcpyFromSymbol(g, g_prof, 64); unsigned long long z[8] = {0}; cudaMemcpyToSymbol(g_prof, z, 64);
                if (g[3]) printf("  | per multi-bidder iteration: values %llu bids %llu logic %llu, whole loop trip %llu (%llu its)", g[0] / g[3], g[1] / g[3], g[2] / g[3], g[4] / g[3], g[3]); }
#endif
            printf("\n");
        }
        cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("CUDA error %s\n", cudaGetErrorString(e));
    }
    return 0;
}
