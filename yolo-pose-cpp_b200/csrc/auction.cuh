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
// Row-parallel single-warp solve for the tracker's common case: at most 32 ACTIVE rows, cost
// matrix in shared memory.  Same results as auction_solve_cta (tests compare them).
//
// Lane i owns the i-th active row (act_list, ascending slots) and one "unassigned" flag;
// prices and owners live in shared memory.  All bidders of an iteration scan their rows in
// parallel (SIMT over rows; a rolled, 4-way software-pipelined loop).  A single bidder applies
// its bid directly; several bidders are grouped by column with match.any, the group's highest
// bid found with one redux and the lowest row among equal bids with one ballot
// (hungarian.cu:100) — no shared-memory atomics (a 64-bit shared atomicMax compiles to a CAS
// spin loop on sm_100a and serialises when many rows want one column).
// The cost of an iteration is ~constant instead of proportional to the number of bidders,
// which bounds the slowest stream of a batch — the tail that sets the kernel time.  The loop
// body is kept small on purpose: it runs up to 150 times per stream-frame and has to stay
// inside the instruction cache (an unrolled multi-path version of this loop ran 3x slower).
// ---------------------------------------------------------------------------------------
struct AucScratch {          // shared memory
    float* price;            // [C]
    int* owner;              // [C] row INDEX (position in act_list), -1 = free
};

static __device__ __noinline__ void auction_solve_rows32(const float* cost_s, int R, int C, const int* act_list, int na,
                                                  int* row, int* col, float* price, int* owner,
                                                  unsigned* colbid, int* colrow) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    for (int t = lane; t < R; t += 32) row[t] = -1;
    for (int d = lane; d < C; d += 32) { col[d] = -1; price[d] = 0.0f; owner[d] = -1; colbid[d] = 0u; colrow[d] = 0x7fffffff; }
    if (na <= 0 || C <= 0) return;
    const bool mine = lane < na;
    const float* cr = cost_s + (size_t)(mine ? act_list[lane] : 0) * C;
    bool unas = mine;
    __syncwarp();
    float eps = 1.0f / (float)(R + 1);                                                     // :378
    const int iters = (R * 3 < 50) ? R * 3 : 50;                                           // :379
    const int C4 = C & ~3;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const unsigned ub = __ballot_sync(FULL, unas);
        if (ub == 0u) break;                                                               // fixed point
        int bc = -1;
        unsigned bid = 0u;
        if (unas) {
            float bv = -1e9f, sv = -1e9f;
#pragma unroll 1
            for (int d = 0; d < C4; d += 4) {                                              // loads first, then the compare chain
                const float v0 = -cr[d] - price[d], v1 = -cr[d + 1] - price[d + 1];        // :61
                const float v2 = -cr[d + 2] - price[d + 2], v3 = -cr[d + 3] - price[d + 3];
                if (v0 > bv) { sv = bv; bv = v0; bc = d; } else if (v0 > sv) sv = v0;      // ascending d: lowest column on ties (:63)
                if (v1 > bv) { sv = bv; bv = v1; bc = d + 1; } else if (v1 > sv) sv = v1;
                if (v2 > bv) { sv = bv; bv = v2; bc = d + 2; } else if (v2 > sv) sv = v2;
                if (v3 > bv) { sv = bv; bv = v3; bc = d + 3; } else if (v3 > sv) sv = v3;
            }
#pragma unroll 1
            for (int d = C4; d < C; ++d) {
                const float v = -cr[d] - price[d];
                if (v > bv) { sv = bv; bv = v; bc = d; } else if (v > sv) sv = v;
            }
            if (bc >= 0) bid = __float_as_uint(bv - sv + eps);                             // :99 (positive: bits order like the floats)
        }
        const unsigned pm = __ballot_sync(FULL, bc >= 0);
        if (pm == 0u) break;                                                               // no bid: fixed point
        bool win = bc >= 0;
        const int nbid = __popc(pm);
        if (nbid > 1 && nbid <= 8) {
            // few bidders: every bidder compares itself with the others (highest bid, lowest row, :100)
            unsigned rem = pm;
            while (rem) {
                const int j = __ffs(rem) - 1;
                rem &= rem - 1;
                const int obc = __shfl_sync(FULL, bc, j);
                const unsigned obid = __shfl_sync(FULL, bid, j);
                if (obc == bc && (obid > bid || (obid == bid && j < lane))) win = false;
            }
        } else if (nbid > 8) {
            if (bc >= 0) atomicMax(&colbid[bc], bid);
            __syncwarp();
            if (bc >= 0 && colbid[bc] == bid) atomicMin(&colrow[bc], lane);
            __syncwarp();
            win = bc >= 0 && colbid[bc] == bid && colrow[bc] == lane;
            __syncwarp();
            if (win) { colbid[bc] = 0u; colrow[bc] = 0x7fffffff; }
        }
        int prev = -1;
        if (win) {                                                                         // :107-121
            prev = owner[bc];
            owner[bc] = lane;
            price[bc] += __uint_as_float(bid);
            unas = false;
        }
        const unsigned em = __reduce_or_sync(FULL, prev >= 0 ? (1u << prev) : 0u);         // evicted owners bid again
        if ((em >> lane) & 1u) unas = true;
        __syncwarp();
        eps *= 0.9f;                                                                       // :402
    }
    __syncwarp();
    for (int d = lane; d < C; d += 32) {
        const int o = owner[d];
        if (o >= 0) { const int slot = act_list[o]; col[d] = slot; row[slot] = d; }
    }
}

}  // namespace pb
