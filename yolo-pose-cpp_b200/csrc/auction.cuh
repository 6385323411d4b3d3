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
        // Every warp derives the same bidder set (unassigned active rows) with ballots over all
        // rows and takes the bidders whose running index is congruent to its warp id, so the
        // bidder rows are spread evenly over the warps without an extra barrier.
        int bidx = 0;
        for (int base = 0; base < R; base += 32) {
            const int r = base + lane;
            const bool bidder = (r < R) && (active == nullptr || active[r] != 0) && (row[r] < 0);
            unsigned bm = __ballot_sync(FULL, bidder);
            while (bm) {
                const int rb = base + __ffs(bm) - 1;
                bm &= bm - 1;
                const bool mine = (bidx % nwarps) == warp;
                ++bidx;
                if (!mine) continue;
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


// ---------------------------------------------------------------------------------------
// Single-warp variant for small problems whose cost matrix sits in shared memory
// (the tracker's 128 x 64 case).  Same semantics and tie-breaks as above, different mapping:
// warp 0 runs the whole solve, one LANE per bidder row (rows r = lane + 32*w), so the bidders
// of an iteration are evaluated in parallel and an iteration needs three warp-level barriers
// instead of two block barriers.  Each lane scans its row's columns starting at a rotated
// offset (bank-conflict free for any column count); ties are resolved by explicit
// (value, column) / (bid, row) comparison, so the scan order does not matter:
//   best column  = lowest column among the maximum values      (hungarian.cu:63)
//   second value = multiset second maximum, floor -1e9          (hungarian.cu:67-69)
//   column owner = highest bid, lowest row among equal bids     (hungarian.cu:100)
// Column bids are positive floats, so a 32-bit shared atomicMax on their bit patterns finds
// the highest bid; a second atomicMin over the rows that reached it finds the lowest row.
// All threads of the CTA call this; it ends with a block barrier.
//   rowbc [R] int, rowbid [R] unsigned, colbid [C] unsigned, colrow [C] int: shared scratch.
// Requires R <= 1024.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void auction_solve_warp(const float* cost, int R, int C, const int* active,
                                                   int* row, int* col, float* price, unsigned* colbid,
                                                   int* colrow, int* rowbc, unsigned* rowbid, int tid) {
    const unsigned FULL = 0xffffffffu;
    if (tid < 32 && R > 0 && C > 0) {
        const int lane = tid;
        const int W = (R + 31) >> 5;
        for (int t = lane; t < R; t += 32) row[t] = -1;
        for (int d = lane; d < C; d += 32) { col[d] = -1; price[d] = 0.0f; colbid[d] = 0u; colrow[d] = 0x7fffffff; }
        __syncwarp();
        float eps = 1.0f / (float)(R + 1);
        const int iters = (R * 3 < 50) ? R * 3 : 50;
        const int rot = (C & 1) ? 0 : 1;                 // lane stride C + rot is odd: no bank conflicts
        for (int it = 0; it < iters; ++it) {
            unsigned mybid = 0u;                         // bit w: my row lane + 32*w placed a bid
            for (int w = 0; w < W; ++w) {
                const int r = lane + 32 * w;
                if (r < R && (active == nullptr || active[r] != 0) && row[r] < 0) {
                    const float* cr = cost + (size_t)r * C;
                    float bv = -1e9f, sv = -1e9f;
                    int bc = -1;
                    int d = (lane * rot) % C;
#pragma unroll 4
                    for (int i = 0; i < C; ++i) {
                        const float v = -cr[d] - price[d];
                        if (v > bv || (v == bv && d < bc)) { sv = bv; bv = v; bc = d; }
                        else if (v > sv) { sv = v; }
                        ++d; if (d == C) d = 0;
                    }
                    if (bc >= 0) {
                        const unsigned bits = __float_as_uint(bv - sv + eps);      // :99
                        rowbc[r] = bc; rowbid[r] = bits;
                        atomicMax(&colbid[bc], bits);
                        mybid |= 1u << w;
                    }
                }
            }
            if (!__any_sync(FULL, mybid != 0u)) break;   // fixed point
            __syncwarp();
            for (int w = 0; w < W; ++w)
                if (mybid & (1u << w)) {
                    const int r = lane + 32 * w;
                    if (rowbid[r] == colbid[rowbc[r]]) atomicMin(&colrow[rowbc[r]], r);
                }
            __syncwarp();
            unsigned won = 0u;
            for (int w = 0; w < W; ++w)
                if (mybid & (1u << w)) {
                    const int r = lane + 32 * w;
                    if (rowbid[r] == colbid[rowbc[r]] && colrow[rowbc[r]] == r) won |= 1u << w;
                }
            __syncwarp();
            for (int w = 0; w < W; ++w)
                if (won & (1u << w)) {                   // :107-121, one winner per column
                    const int r = lane + 32 * w;
                    const int bc = rowbc[r];
                    const int prev = col[bc];
                    if (prev >= 0) row[prev] = -1;
                    col[bc] = r;
                    row[r] = bc;
                    price[bc] += __uint_as_float(rowbid[r]);
                    colbid[bc] = 0u; colrow[bc] = 0x7fffffff;
                }
            __syncwarp();
            eps *= 0.9f;                                 // :402
        }
    }
    __syncthreads();
}


// ---------------------------------------------------------------------------------------
// Register-resident single-warp variant: columns <= 32*CPL (CPL 1 or 2), rows <= 128.
// Lane d owns column d (and d+32): its price, its owner and this iteration's best bid live in
// registers; the set of unassigned active rows is four uniform 32-bit masks.  One bidder row
// at a time: every lane loads its column's cost, the warp finds (best value, lowest column)
// and the second value with three redux.sync operations on order-preserving integer keys, and
// the owning lane records the bid.  No shared-memory atomics, no barriers inside the loop.
// Same results as auction_solve_cta (explicit tie-breaks: lowest column among equal values,
// lowest row among equal bids because rows are visited in ascending order with strict '>').
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned auc_ord(float v) {          // order-preserving float -> uint
    unsigned u = __float_as_uint(v);
    if (u == 0x80000000u) u = 0u;                               // -0 and +0 compare equal
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float auc_dec(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

template <int CPL>
__device__ __forceinline__ void auction_solve_regs(const float* cost, int R, int C, const int* active,
                                                   int* row, int* col, int tid) {
    const unsigned FULL = 0xffffffffu;
    if (tid < 32 && R > 0 && C > 0) {
        const int lane = tid;
        const unsigned kFloor = auc_ord(-1e9f);
        float price[CPL];
        int owner[CPL];
#pragma unroll
        for (int q = 0; q < CPL; ++q) { price[q] = 0.0f; owner[q] = -1; }
        unsigned um[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int r = 32 * w + lane;
            um[w] = __ballot_sync(FULL, r < R && (active == nullptr || active[r] != 0));
        }
        float eps = 1.0f / (float)(R + 1);
        const int iters = (R * 3 < 50) ? R * 3 : 50;
        for (int it = 0; it < iters; ++it) {
            unsigned bidbits[CPL];
            int bidrow[CPL];
#pragma unroll
            for (int q = 0; q < CPL; ++q) { bidbits[q] = 0u; bidrow[q] = -1; }
            bool any = false;
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                unsigned bm = um[w];
                while (bm) {
                    const int rb = 32 * w + __ffs(bm) - 1;
                    bm &= bm - 1;
                    const float* cr = cost + (size_t)rb * C;
                    unsigned kb = 0u, ks = 0u;
                    int qb = 0;
#pragma unroll
                    for (int q = 0; q < CPL; ++q) {
                        const int d = lane + 32 * q;
                        const unsigned k = (d < C) ? auc_ord(-cr[d] - price[q]) : 0u;     // :61
                        if (k > kb) { ks = kb; kb = k; qb = q; }
                        else if (k > ks) { ks = k; }
                    }
                    const unsigned m1 = __reduce_max_sync(FULL, kb);
                    if (m1 <= kFloor) continue;                  // no column with value > -1e9 (:55-63)
                    const unsigned bc = __reduce_min_sync(FULL, (kb == m1) ? (unsigned)(lane + 32 * qb) : 0x7fffffffu);
                    unsigned m2 = __reduce_max_sync(FULL, ((unsigned)(lane + 32 * qb) == bc) ? ks : kb);
                    if (m2 < kFloor) m2 = kFloor;
                    const float bid = auc_dec(m1) - auc_dec(m2) + eps;                     // :99
                    if ((int)(bc & 31u) == lane) {
#pragma unroll
                        for (int q = 0; q < CPL; ++q)
                            if ((int)(bc >> 5) == q && (bidrow[q] < 0 || bid > __uint_as_float(bidbits[q]))) {
                                bidbits[q] = __float_as_uint(bid); bidrow[q] = rb;         // :100
                            }
                    }
                    any = true;
                }
            }
            if (!any) break;                                     // fixed point
            unsigned clr[4] = {0u, 0u, 0u, 0u}, set[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                if (bidrow[q] >= 0) {                            // :107-121
                    const int prev = owner[q];
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        if ((bidrow[q] >> 5) == w) clr[w] |= 1u << (bidrow[q] & 31);
                        if (prev >= 0 && (prev >> 5) == w) set[w] |= 1u << (prev & 31);
                    }
                    owner[q] = bidrow[q];
                    price[q] += __uint_as_float(bidbits[q]);
                }
            }
#pragma unroll
            for (int w = 0; w < 4; ++w)
                um[w] = (um[w] & ~__reduce_or_sync(FULL, clr[w])) | __reduce_or_sync(FULL, set[w]);
            eps *= 0.9f;                                         // :402
        }
        for (int t = lane; t < R; t += 32) row[t] = -1;
        __syncwarp();
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            const int d = lane + 32 * q;
            if (d < C) { col[d] = owner[q]; if (owner[q] >= 0) row[owner[q]] = d; }
        }
    }
    __syncthreads();
}

}  // namespace pb
