// auction.cuh — forward auction for one assignment problem, executed by one CTA.
//
// Semantics of LinearAssignmentCUDA::solveDeviceAsyncWithActive (reference
// src/cuda/hungarian.cu:358-405) with kernelAuctionBidding (:27-75) and
// kernelAuctionAssignment (:78-123): assignments reset to -1, prices to 0,
// eps0 = 1/(rows+1), min(3*rows, 50) iterations, eps *= 0.9 per iteration; every
// unassigned active row bids best-second+eps on its best column (first/lowest column on
// equal value), every column takes its highest bid (lowest row on equal bid), evicts its
// owner and raises its price by the bid.  The `threshold` argument is ignored upstream.
//
// Mapping: one warp scans one bidder row (lanes stride the columns, shuffle top-2
// reduction); bids meet in a packed 64-bit shared-memory atomicMax
// (bid bits << 32 | ~row: bids are > 0, so their bit patterns order like the floats);
// two block barriers per iteration.  The loop stops at the first iteration in which no
// row bids: prices and assignments are then a fixed point of all remaining iterations.
#pragma once
#include "pb_common.cuh"

namespace pb {

// cost: flat [rows, cols] (shared or global).  active may be nullptr (all rows active).
// row/col/price/colbid/flags(2 ints) must be shared memory.  Ends with a block barrier.
__device__ __forceinline__ void auction_solve_cta(const float* cost, int R, int C, const int* active,
                                                  int* row, int* col, float* price,
                                                  unsigned long long* colbid, int* flags,
                                                  int tid, int nthreads) {
    const unsigned FULL = 0xffffffffu;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    for (int t = tid; t < R; t += nthreads) row[t] = -1;
    for (int d = tid; d < C; d += nthreads) { col[d] = -1; price[d] = 0.0f; colbid[d] = 0ull; }
    if (tid == 0) { flags[0] = 0; flags[1] = 0; }
    float eps = 1.0f / (float)(R + 1);                                     // :378
    const int iters = (R * 3 < 50) ? R * 3 : 50;                           // :379
    __syncthreads();
    if (R == 0 || C == 0) return;                                          // :368
    for (int it = 0; it < iters; ++it) {
        for (int base = warp * 32; base < R; base += nwarps * 32) {
            const int r = base + lane;
            const bool bidder = (r < R) && (active == nullptr || active[r] != 0) && (row[r] < 0);
            unsigned bm = __ballot_sync(FULL, bidder);
            while (bm) {
                const int rb = base + __ffs(bm) - 1;
                bm &= bm - 1;
                const float* cr = cost + (size_t)rb * C;
                float bv = -1e9f, sv = -1e9f;
                int bc = -1;
                for (int d = lane; d < C; d += 32) {
                    const float v = -cr[d] - price[d];                     // :61
                    if (v > bv) { sv = bv; bv = v; bc = d; }
                    else if (v > sv) { sv = v; }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const float ob = __shfl_xor_sync(FULL, bv, off);
                    const float os = __shfl_xor_sync(FULL, sv, off);
                    const int oc = __shfl_xor_sync(FULL, bc, off);
                    // lowest column wins ties (strict '>' in ascending column order, :63);
                    // a lane without candidate (bc = -1, bv = -1e9) never wins.
                    const bool other = (oc >= 0) && ((bc < 0) || (ob > bv) || (ob == bv && oc < bc));
                    if (other) { sv = pb_max(os, bv); bv = ob; bc = oc; }
                    else { sv = pb_max(sv, ob); }
                }
                if (lane == 0 && bc >= 0) {
                    const float bid = bv - sv + eps;                       // :99
                    const unsigned long long key = ((unsigned long long)__float_as_uint(bid) << 32) |
                                                   (unsigned long long)(0xffffffffu - (unsigned)rb);
                    atomicMax(&colbid[bc], key);                           // highest bid, lowest row (:100)
                    flags[it & 1] = 1;
                }
            }
        }
        __syncthreads();
        for (int d = tid; d < C; d += nthreads) {                          // :93-122
            const unsigned long long key = colbid[d];
            if (key != 0ull) {
                const int winner = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
                const float bid = __uint_as_float((unsigned)(key >> 32));
                const int prev = col[d];
                if (prev >= 0) row[prev] = -1;
                col[d] = winner;
                row[winner] = d;
                price[d] += bid;
                colbid[d] = 0ull;
            }
        }
        if (tid == 0) flags[(it + 1) & 1] = 0;
        __syncthreads();
        if (!flags[it & 1]) break;
        eps *= 0.9f;                                                       // :402
    }
    __syncthreads();
}

}  // namespace pb
