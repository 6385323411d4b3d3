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
// Mapping: one warp scans one bidder row (lanes stride the columns, then three CREDUX
// reductions on order-preserving keys); bids meet in a packed 64-bit shared-memory atomicMax
// (bid bits << 32 | ~row: bids are > 0, so their bit patterns order like the floats);
// two block barriers per iteration.  The loop stops at the first iteration in which no
// row bids: prices and assignments are then a fixed point of all remaining iterations.
#pragma once
#include "pb_common.cuh"

namespace pb {

// float order -> unsigned order and back (warp reductions on values run on these keys)
__device__ __forceinline__ unsigned hy_ord(float f) {            // float order -> unsigned order (-0.0 == +0.0)
    unsigned b = __float_as_uint(f);
    b = (b == 0x80000000u) ? 0u : b;
    return b ^ ((unsigned)((int)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float hy_unord(unsigned u) {
    return __uint_as_float(u ^ (((u >> 31) - 1u) | 0x80000000u));
}

// cost: flat [rows, cols] (shared or global).  active may be nullptr (all rows active).
// row/col/price/colbid/flags(2 ints) must be shared memory.  Ends with a block barrier.
__device__ __forceinline__ void auction_solve_cta(const float* cost, int R, int C, const int* active,
                                                  int* row, int* col, float* price,
                                                  unsigned long long* colbid, int* flags,
                                                  int tid, int nthreads, int max_iters = -1, const int* locked = nullptr) {
    const unsigned FULL = 0xffffffffu;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    for (int t = tid; t < R; t += nthreads) row[t] = -1;
    for (int d = tid; d < C; d += nthreads) { col[d] = -1; price[d] = 0.0f; colbid[d] = 0ull; }
    if (tid == 0) { flags[0] = 0; flags[1] = 0; }
    float eps = 1.0f / (float)(R + 1);                                     // :378
    const int iters = max_iters >= 0 ? max_iters : ((R * 3 < 50) ? R * 3 : 50);   // :379 (legacy solve: 3R, :283)
    __syncthreads();
    if (R == 0 || C == 0) return;                                          // :368
    for (int it = 0; it < iters; ++it) {
        // Every warp derives the same bidder set (unassigned active rows) with ballots over all
        // rows and takes the bidders whose running index is congruent to its warp id, so the
        // bidder rows are spread evenly over the warps without an extra barrier.
        int bidx = 0;
        for (int base = 0; base < R; base += 32) {
            const int r = base + lane;
            // locked (may be null): rows with locked[r] >= 0 were matched in an earlier tier; all their cells are 1e9 (lock_pairs),
            // so they would scan their row, find nothing above the -1e9 floor and never bid (hungarian.cu:55-69): left out
            const bool bidder = (r < R) && (active == nullptr || active[r] != 0) && (row[r] < 0) && (locked == nullptr || locked[r] < 0);
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
                // eight independent loads in flight per lane (the cost matrix of large tables lives in global memory / L2:
                // one round trip per group instead of one per element)
                for (int d0 = lane; d0 < C; d0 += 32 * 8) {
                    float cv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) { const int d = d0 + 32 * u; cv[u] = (d < C) ? cr[d] : 0.0f; }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int d = d0 + 32 * u;
                        if (d < C) {
                            const float v = -cv[u] - price[d];             // :61
                            if (v > bv) { sv = bv; bv = v; bc = d; }
                            else if (v > sv) { sv = v; }
                        }
                    }
                }
                // warp-level top-2 with three CREDUX reductions on order-preserving integer keys (see the lean
                // solve below): best value, lowest column holding it (:63), best of the other columns (:67-69)
                const unsigned kb = hy_ord(bv);
                const unsigned m = __reduce_max_sync(FULL, kb);
                if (m != hy_ord(-1e9f)) {                                  // some column is above the -1e9 floor
                    const int key = (kb == m) ? bc : 0x7fffffff;
                    const int bcw = __reduce_min_sync(FULL, key);
                    const unsigned m2 = __reduce_max_sync(FULL, (key == bcw) ? hy_ord(sv) : kb);
                    if (lane == 0) {
                        const float bid = hy_unord(m) - hy_unord(m2) + eps;    // :99
                        const unsigned long long kk = ((unsigned long long)__float_as_uint(bid) << 32) |
                                                      (unsigned long long)(0xffffffffu - (unsigned)rb);
                        atomicMax(&colbid[bcw], kk);                       // highest bid, lowest row (:100)
                        flags[it & 1] = 1;
                    }
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
// Hybrid single-warp solve: at most 32 active rows, any number of columns (the lean solve below takes the
// tracker's own case of at most 64 columns).
//
// The bidder set is a warp-UNIFORM bitmask `ub` (bit i = i-th active row is unassigned), so no
// ballot is needed to find it.  Per iteration one of two mappings is chosen by the number of
// bidders:
//   * many bidders (> HY_COLPAR_MAX, the first iterations of a solve): lane = row, every bidder
//     scans its own cost row;
//   * few bidders (the long tail: an eviction chain of one to four rows that runs to the
//     iteration limit whenever a stream has one active row more than it has detections):
//     lane = column.  For each bidder the 32 lanes evaluate value = -cost - price of their columns
//     in one step and three warp reductions on order-preserving integer keys give the best value,
//     the lowest column holding it (hungarian.cu:63) and the second value (multiset second
//     maximum, floor -1e9; :67-69).  All lanes then hold the same (column, bid) and update the
//     per-column highest bid (ascending rows + strict '>' = lowest row among equal bids, :100),
//     owner, price and `ub` redundantly — no broadcast, no atomics.
// An iteration of the tail costs ~1/4 of a row scan.  Results are those of auction_solve_cta
// bit for bit (tests compare the three implementations on random and degenerate problems).
// ---------------------------------------------------------------------------------------
constexpr int HY_COLPAR_MAX = 4;

static __device__ __noinline__ void auction_solve_hybrid32(const float* cost_s, int R, int C, const int* act_list, int na,
                                                           int* row, int* col, float* price, int* owner,
                                                           unsigned* colbid, int* colrow) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    for (int t = lane; t < R; t += 32) row[t] = -1;
    for (int d = lane; d < C; d += 32) { col[d] = -1; price[d] = 0.0f; owner[d] = -1; colbid[d] = 0u; colrow[d] = 0x7fffffff; }
    if (na <= 0 || C <= 0) return;
    const bool mine = lane < na;
    const float* cr = cost_s + (size_t)(mine ? act_list[lane] : 0) * C;
    unsigned ub = (na >= 32) ? FULL : ((1u << na) - 1u);
    __syncwarp();
    float eps = 1.0f / (float)(R + 1);                                                     // :378
    const int iters = (R * 3 < 50) ? R * 3 : 50;                                           // :379
    const int C4 = C & ~3;
    const unsigned ORD_FLOOR = hy_ord(-1e9f);
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        if (ub == 0u) break;                                                               // fixed point
        if (__popc(ub) <= HY_COLPAR_MAX) {
            // ---- lane = column ----
            unsigned long long bcs = 0ull;                                                 // bid columns, 16 bits each
            int nbids = 0;
            unsigned rem = ub;
#pragma unroll 1
            while (rem) {
                const int j = __ffs(rem) - 1;
                rem &= rem - 1;
                const float* crj = cost_s + (size_t)act_list[j] * C;
                float bv = -1e9f, sv = -1e9f;
                int bc = 0x7fffffff;
#pragma unroll 1
                for (int d = lane; d < C; d += 32) {
                    const float v = -crj[d] - price[d];                                    // :61
                    if (v > bv) { sv = bv; bv = v; bc = d; } else if (v > sv) sv = v;
                }
                const unsigned kb = hy_ord(bv);
                const unsigned m = __reduce_max_sync(FULL, kb);
                if (m == ORD_FLOOR) continue;                                              // no column above -1e9: no bid
                const int bcj = __reduce_min_sync(FULL, (kb == m) ? bc : 0x7fffffff);      // lowest column among the best
                const unsigned m2 = __reduce_max_sync(FULL, hy_ord(bc == bcj ? sv : bv));  // best of the other columns
                const unsigned bits = __float_as_uint(hy_unord(m) - hy_unord(m2) + eps);   // :99 (positive)
                const unsigned cur = colbid[bcj];
                if (bits > cur) { colbid[bcj] = bits; colrow[bcj] = j; }                   // ascending j: lowest row on ties
                bcs |= (unsigned long long)bcj << (16 * nbids);
                ++nbids;
            }
            if (nbids == 0) break;                                                         // no bid: fixed point
            __syncwarp();
#pragma unroll 1
            for (int k = 0; k < nbids; ++k) {                                              // :107-121
                const int bc = (int)((bcs >> (16 * k)) & 0xffffull);
                const unsigned wb = colbid[bc];
                const int w = colrow[bc], prev = owner[bc];
                const float p = price[bc];
                __syncwarp();
                if (wb != 0u) {
                    owner[bc] = w; price[bc] = p + __uint_as_float(wb);
                    colbid[bc] = 0u; colrow[bc] = 0x7fffffff;
                    ub &= ~(1u << w);
                    if (prev >= 0) ub |= 1u << prev;
                }
                __syncwarp();
            }
        } else {
            // ---- lane = row ----
            const bool unas = (ub >> lane) & 1u;
            int bc = -1;
            unsigned bid = 0u;
            if (unas) {
                float bv = -1e9f, sv = -1e9f;
#pragma unroll 1
                for (int d = 0; d < C4; d += 4) {
                    const float v0 = -cr[d] - price[d], v1 = -cr[d + 1] - price[d + 1];    // :61
                    const float v2 = -cr[d + 2] - price[d + 2], v3 = -cr[d + 3] - price[d + 3];
                    if (v0 > bv) { sv = bv; bv = v0; bc = d; } else if (v0 > sv) sv = v0;  // ascending d: lowest column on ties (:63)
                    if (v1 > bv) { sv = bv; bv = v1; bc = d + 1; } else if (v1 > sv) sv = v1;
                    if (v2 > bv) { sv = bv; bv = v2; bc = d + 2; } else if (v2 > sv) sv = v2;
                    if (v3 > bv) { sv = bv; bv = v3; bc = d + 3; } else if (v3 > sv) sv = v3;
                }
#pragma unroll 1
                for (int d = C4; d < C; ++d) {
                    const float v = -cr[d] - price[d];
                    if (v > bv) { sv = bv; bv = v; bc = d; } else if (v > sv) sv = v;
                }
                if (bc >= 0) bid = __float_as_uint(bv - sv + eps);                         // :99
            }
            const unsigned pm = __ballot_sync(FULL, bc >= 0);
            if (pm == 0u) break;                                                           // no bid: fixed point
            if (bc >= 0) atomicMax(&colbid[bc], bid);
            __syncwarp();
            if (bc >= 0 && colbid[bc] == bid) atomicMin(&colrow[bc], lane);
            __syncwarp();
            const bool win = bc >= 0 && colbid[bc] == bid && colrow[bc] == lane;
            __syncwarp();
            int prev = -1;
            if (win) {                                                                     // :107-121
                colbid[bc] = 0u; colrow[bc] = 0x7fffffff;
                prev = owner[bc];
                owner[bc] = lane;
                price[bc] += __uint_as_float(bid);
            }
            const unsigned wm = __ballot_sync(FULL, win);
            const unsigned em = __reduce_or_sync(FULL, prev >= 0 ? (1u << prev) : 0u);     // evicted owners bid again
            ub = (ub & ~wm) | em;
            __syncwarp();
        }
        eps *= 0.9f;                                                                       // :402
    }
    __syncwarp();
    for (int d = lane; d < C; d += 32) {
        const int o = owner[d];
        if (o >= 0) { const int slot = act_list[o]; col[d] = slot; row[slot] = d; }
    }
}


// ---------------------------------------------------------------------------------------
// Lean single-warp solve: at most 32 active rows, at most 64 columns (the tracker's own case:
// max_detections = 64).  Same decisions as auction_solve_cta bit for bit.
//
// lane = column for every iteration with up to eight bidders (all but the first iteration of a
// solve in practice).  Each lane keeps the price and the owner of its NC columns in REGISTERS;
// the cost rows of the active rows were compacted to cc[i*C + d] (i = position in act_list) by
// the whole CTA, so a bidder's row is one conflict-free load.  A bidder is evaluated by all
// lanes at once: value = -cost - price of the lane's columns, then three CREDUX reductions on
// order-preserving integer keys: the best value, the lowest column holding it packed with that
// column's owner (hungarian.cu:63), the best of the other columns (multiset second maximum,
// floor -1e9; :67-69).  Every lane then holds the same (column, bid, evicted owner) for every
// bidder, so who wins a column (highest bid, lowest row among equal bids, :100), the price and
// owner update (by the lane owning the column) and the next bidder set are plain warp-uniform
// arithmetic: no ballot, no shuffle, no atomics, no shared-memory traffic inside the loop.
//   * 1 bidder: the eviction chain.  The next bidder is the evicted owner, known right after
//     the second reduction, so consecutive iterations overlap.
//   * 2-4 bidders: independent instruction chains, evaluated together.
//   * 5-8 bidders: one after the other, per-lane best bid, one REDUX.OR of toggle bits.
// Rows that find no column above -1e9 leave the bidder set for good (prices only rise, so they
// can never bid again; they stay unassigned exactly as upstream).
// With more than eight bidders (iteration 0 of a solve) lane = row scans the compacted rows.
// ---------------------------------------------------------------------------------------
// sm_100a warp reductions on fp32 (CREDUX.MAX.F32): no order-preserving integer keys, no conversions on the chain
__device__ __forceinline__ float warp_max_f32(float v) {
    float m;
    asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
    return m;
}

// value of the lane's columns for bidder row j: fmaxf drops NaN and everything at or below the -1e9 floor
// (never best, never second: :55-69).  -0.0 and +0.0 compare equal below, as in the reference's comparisons.
template <int NC>
__device__ __forceinline__ void lean_values(const float* cc, int C, int j, const float (&p)[NC], int lane,
                                            float& bv, float& sv, int& bsel) {
    float v[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int d = lane + 32 * c;
        v[c] = (d < C) ? fmaxf(-cc[j * C + d] - p[c], -1e9f) : -1e9f;                      // :61
    }
    bv = v[0]; sv = -1e9f; bsel = 0;
    if (NC == 2) {
        if (v[NC - 1] > bv) { sv = bv; bv = v[NC - 1]; bsel = 1; } else sv = v[NC - 1];
    }
}

#ifdef LEAN_PROFILE
__device__ unsigned long long g_prof[8];
#define LEAN_T(x) const long long x = clock64();
#else
#define LEAN_T(x)
#endif
struct LeanBid { float m; int rk; unsigned bits; };   // warp-uniform: best value, (column << 8 | owner + 1), bid bits

// After the first reduction (best value m) the other three are independent of each other: the lowest column holding m
// packed with its owner (:63), the best of the values that are not a lane's own m, and the lanes holding m.  With two or
// more such lanes the second value is m itself (multiset second maximum, :67-69), otherwise the best of the rest.
template <int NC>
__device__ __forceinline__ LeanBid lean_bid(float bv, float sv, int bsel, const int (&own)[NC], int lane, float eps) {
    const unsigned FULL = 0xffffffffu;
    LeanBid r;
    r.m = warp_max_f32(bv);                                                                // best value
    const bool top = (bv == r.m);
    const int osel = (NC == 2 && bsel) ? own[NC - 1] : own[0];
    const int key = top ? (((lane + 32 * bsel) << 8) | (osel + 1)) : 0x7fffffff;
    r.rk = __reduce_min_sync(FULL, key);                                                   // lowest column holding it (:63) + its owner
    const float m2p = warp_max_f32(top ? sv : bv);
    const unsigned ties = __ballot_sync(FULL, top);
    const float m2 = (ties & (ties - 1u)) ? r.m : m2p;                                     // best of the other columns
    r.bits = __float_as_uint(r.m - m2 + eps);                                              // :99 (positive)
    return r;
}

// one iteration with NB (2..4) bidders given as isolated bits of ub (ascending rows); returns false when nobody bid.
// Branch-free warp-uniform resolution: a bidder without a column gets a unique negative column and bid 0.
template <int NB, int NC>
__device__ __forceinline__ bool lean_iter(const float* cc, int C, unsigned& ub, const unsigned (&bit)[NB], float eps,
                                          float (&p)[NC], int (&own)[NC], int lane) {
    const float ORD_FLOOR = -1e9f;
    LEAN_T(t0)
    int j[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) j[b] = 31 - __clz(bit[b]);
    float bv[NB], sv[NB];
    int bsel[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) lean_values<NC>(cc, C, j[b], p, lane, bv[b], sv[b], bsel[b]);
    LEAN_T(t1)
    LeanBid q[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) q[b] = lean_bid<NC>(bv[b], sv[b], bsel[b], own, lane, eps);
    LEAN_T(t2)
    int bc[NB];
    unsigned vm[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        vm[b] = (q[b].m != ORD_FLOOR) ? 0xffffffffu : 0u;                                  // bids at all
        bc[b] = (q[b].m != ORD_FLOOR) ? (q[b].rk >> 8) : (-1 - b);
    }
    unsigned nub = ub, anym = 0u;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        unsigned lose = 0u;                                                                // highest bid, lowest row (:100)
#pragma unroll
        for (int o = 0; o < NB; ++o) {
            if (o < b) lose |= (bc[o] == bc[b] && q[o].bits >= q[b].bits) ? 0xffffffffu : 0u;
            if (o > b) lose |= (bc[o] == bc[b] && q[o].bits > q[b].bits) ? 0xffffffffu : 0u;
        }
        const unsigned wm = vm[b] & ~lose;
        const int ow = q[b].rk & 0xff;                                                     // owner + 1
        const unsigned evict = ow ? (1u << (ow - 1)) : 0u;
        nub ^= bit[b] & (wm | ~vm[b]);                                                     // winners and rows without a column leave
        nub |= evict & wm;                                                                 // evicted owners enter (:109-111)
        anym |= wm;
#pragma unroll
        for (int c = 0; c < NC; ++c)
            if (wm != 0u && bc[b] == lane + 32 * c) { p[c] += __uint_as_float(q[b].bits); own[c] = j[b]; }   // :113-118
    }
    ub = nub;
#ifdef LEAN_PROFILE
    const long long t3 = clock64();
    if (lane == 0) { g_prof[0] += t1 - t0; g_prof[1] += t2 - t1; g_prof[2] += t3 - t2; g_prof[3] += 1; }
#endif
    return anym != 0u;
}

// Runs iterations while exactly NB rows bid (an eviction chase usually keeps its number of bidders for many
// iterations): the next bidders are peeled off the new mask with 2*NB operations, no dispatch in between.
// Returns false at a fixed point (nobody bid); otherwise `it`/`eps` are advanced past the iterations done.
template <int NB, int NC>
__device__ __forceinline__ bool lean_regime(const float* cc, int C, unsigned& ub, float& eps, int& it, int iters,
                                            float (&p)[NC], int (&own)[NC], int lane) {
    unsigned bit[NB];
    unsigned rest = ub;
#pragma unroll
    for (int b = 0; b < NB; ++b) { bit[b] = rest & (0u - rest); rest ^= bit[b]; }
#pragma unroll 1
    for (;;) {
        if (!lean_iter<NB, NC>(cc, C, ub, bit, eps, p, own, lane)) return false;
        eps *= 0.9f;                                                                       // :402
        ++it;
        rest = ub;
#pragma unroll
        for (int b = 0; b < NB; ++b) { bit[b] = rest & (0u - rest); rest ^= bit[b]; }
        if (it >= iters || rest != 0u || bit[NB - 1] == 0u) return true;                   // limit reached or another bidder count
    }
}

// The bids of a solve's FIRST iteration (prices 0, nobody assigned) depend on a row's own costs only: one warp per row,
// lane = column, the arithmetic of lean_bid.  The tracker computes them with the whole CTA while it compacts the cost
// rows, so the single-warp solve below does not start with a 32-lane row scan (measured: ~2 250 of the ~4 300 cycles
// of a typical tier-1 solve).  cr: the row's costs [C]; bc (-1: no column above the floor) and bid bits as lean32 wants them.
template <int NC>
__device__ __forceinline__ void lean_first_bid(const float* cr, int C, int R, int lane, int& bc, unsigned& bits) {
    float p0[NC];
    int own0[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { p0[c] = 0.0f; own0[c] = -1; }
    float bv, sv;
    int bsel;
    lean_values<NC>(cr, C, 0, p0, lane, bv, sv, bsel);
    const LeanBid q = lean_bid<NC>(bv, sv, bsel, own0, lane, 1.0f / (float)(R + 1));
    bc = (q.m != -1e9f) ? (q.rk >> 8) : -1;
    bits = q.bits;
}

// pre_bc / pre_bits (may be null): the first iteration's bids of the active rows, by position in act_list (lean_first_bid)
// COPY: a second instantiation for a second call site (tracker_body.cuh: auction_solve_unlocked), so that the compiler
// keeps specialising the hot one for its single caller's shared-memory pointers (measured: +17 % on the eviction chase when
// both callers shared one copy with generic loads)
template <int NC, int COPY = 0>
static __device__ __noinline__ void auction_solve_lean32(const float* cc, int R, int C, const int* act_list, int na,
                                                         int* row, int* col, float* price, int* owner,
                                                         unsigned* colbid, int* colrow, unsigned ub0 = 0xffffffffu,
                                                         unsigned long long* tele = nullptr,
                                                         const int* pre_bc = nullptr, const unsigned* pre_bits = nullptr) {
#ifdef PB_AUCTION_TELE
#define LEAN_COUNT(slot) if (tele && lane == 0) tele[slot] += 1000ull;
#else
#define LEAN_COUNT(slot)
#endif
    const unsigned FULL = 0xffffffffu;
    const float ORD_FLOOR = -1e9f;
    const int lane = threadIdx.x & 31;
    for (int t = lane; t < R; t += 32) row[t] = -1;
    for (int d = lane; d < C; d += 32) { col[d] = -1; colbid[d] = 0u; colrow[d] = 0x7fffffff; }
    if (na <= 0 || C <= 0) return;
    float p[NC];
    int own[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { p[c] = 0.0f; own[c] = -1; }
    // unassigned rows that may still bid; ub0 lets the caller leave out rows it knows to hold no cell below 1e9
    // (they would scan their row, find nothing and never bid: hungarian.cu:55-69)
    unsigned ub = ((na >= 32) ? FULL : ((1u << na) - 1u)) & ub0;
    __syncwarp();
    float eps = 1.0f / (float)(R + 1);                                                     // :378
    const int iters = (R * 3 < 50) ? R * 3 : 50;                                           // :379
    const int C4 = C & ~3;
    int it = 0;
    if (pre_bc != nullptr && ub != 0u && iters > 0) {
        // ---- iteration 0 from the caller's first bids (lean_first_bid): every lane takes, for its own columns, the highest
        // bid (ascending rows + '>' = lowest row among equal bids, :100) out of broadcast loads; nobody is evicted yet ----
        unsigned best[NC];
        int w[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) { best[c] = 0u; w[c] = -1; }
        unsigned drop = 0u;
#pragma unroll 1
        for (int j0 = 0; j0 < na; j0 += 8) {                                               // eight rows per round: independent broadcast loads
            int bcj[8];
            unsigned bj[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int jb = j0 + u;
                const bool in = (jb < na) && ((ub >> jb) & 1u);
                bcj[u] = in ? pre_bc[jb] : -2;
                bj[u] = in ? pre_bits[jb] : 0u;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (bcj[u] == -1) drop |= 1u << (j0 + u);                                  // no column above -1e9: leaves for good
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if (bcj[u] == lane + 32 * c && bj[u] > best[c]) { best[c] = bj[u]; w[c] = j0 + u; }
            }
        }
        unsigned tog = 0u;
#pragma unroll
        for (int c = 0; c < NC; ++c)
            if (w[c] >= 0) { tog |= 1u << w[c]; own[c] = w[c]; p[c] += __uint_as_float(best[c]); }   // :107-121
        tog = __reduce_or_sync(FULL, tog);
        ub = (ub & ~drop) ^ tog;                                                           // winners leave
        if (tog == 0u) ub = 0u;                                                            // no bid: fixed point
        eps *= 0.9f;                                                                       // :402
        it = 1;
    }
#pragma unroll 1
    while (it < iters && ub != 0u) {
#ifdef LEAN_PROFILE
        { static __device__ long long last; const long long now = clock64(); if (lane == 0) { if (it > 1) g_prof[4] += now - last; last = now; } }
#endif
        // how many bidders?  (peeling the mask: no bidder count, no jump table on the way to the common cases)
        const unsigned r1 = ub & (ub - 1u), r2 = r1 & (r1 - 1u), r3 = r2 & (r2 - 1u), r4 = r3 & (r3 - 1u);
        bool any = true;
        if (r1 == 0u) {
            // ---- the eviction chain: one bidder per iteration until the limit or the first free column ----
            int j = 31 - __clz(ub);
#pragma unroll 1
            for (; it < iters; ++it) {
                float bv, sv;
                int bsel;
                LEAN_COUNT(15)
                lean_values<NC>(cc, C, j, p, lane, bv, sv, bsel);
                const LeanBid q = lean_bid<NC>(bv, sv, bsel, own, lane, eps);
                if (q.m == ORD_FLOOR) break;                                               // cannot bid: fixed point
                const int bc = q.rk >> 8;
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if (bc == lane + 32 * c) { p[c] += __uint_as_float(q.bits); own[c] = j; }
                j = (q.rk & 0xff) - 1;                                                     // the evicted owner bids next
                eps *= 0.9f;                                                               // :402
                if (j < 0) break;                                                          // everybody is assigned
            }
            break;
        } else if (r4 == 0u) {
            if (r2 == 0u) { if (!lean_regime<2, NC>(cc, C, ub, eps, it, iters, p, own, lane)) break; }
            else if (r3 == 0u) { if (!lean_regime<3, NC>(cc, C, ub, eps, it, iters, p, own, lane)) break; }
            else { if (!lean_regime<4, NC>(cc, C, ub, eps, it, iters, p, own, lane)) break; }
            continue;                                                                      // it and eps are already advanced
        } else if (__popc(ub) <= 8) {
            // ---- one bidder after the other: per-lane highest bid, toggle bits through one REDUX.OR ----
            LEAN_COUNT(18)
            unsigned best[NC];
            int w[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) { best[c] = 0u; w[c] = -1; }
            unsigned rem = ub, drop = 0u;
#pragma unroll 1
            while (rem) {
                const int jb = __ffs(rem) - 1;
                rem &= rem - 1;
                float bv, sv;
                int bsel;
                lean_values<NC>(cc, C, jb, p, lane, bv, sv, bsel);
                const LeanBid q = lean_bid<NC>(bv, sv, bsel, own, lane, eps);
                if (q.m == ORD_FLOOR) { drop |= 1u << jb; continue; }
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if ((q.rk >> 8) == lane + 32 * c && q.bits > best[c]) { best[c] = q.bits; w[c] = jb; }   // ascending rows + '>'
            }
            unsigned tog = 0u;
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (w[c] >= 0) {                                                           // :107-121
                    tog |= 1u << w[c];
                    if (own[c] >= 0) tog |= 1u << own[c];
                    own[c] = w[c];
                    p[c] += __uint_as_float(best[c]);
                }
            tog = __reduce_or_sync(FULL, tog);
            any = tog != 0u;
            ub = (ub & ~drop) ^ tog;                                                       // winners leave, evicted owners enter
        } else {
            // ---- lane = row ----
            LEAN_COUNT(19)
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (lane + 32 * c < C) { price[lane + 32 * c] = p[c]; owner[lane + 32 * c] = own[c]; }
            __syncwarp();
            const bool unas = (ub >> lane) & 1u;
            const float* cr = cc + lane * C;
            int bc = -1;
            unsigned bid = 0u;
            if (unas) {
                float bv = -1e9f, sv = -1e9f;
#pragma unroll 1
                for (int d = 0; d < C4; d += 4) {
                    const float v0 = -cr[d] - price[d], v1 = -cr[d + 1] - price[d + 1];    // :61
                    const float v2 = -cr[d + 2] - price[d + 2], v3 = -cr[d + 3] - price[d + 3];
                    if (v0 > bv) { sv = bv; bv = v0; bc = d; } else if (v0 > sv) sv = v0;  // ascending d: lowest column on ties (:63)
                    if (v1 > bv) { sv = bv; bv = v1; bc = d + 1; } else if (v1 > sv) sv = v1;
                    if (v2 > bv) { sv = bv; bv = v2; bc = d + 2; } else if (v2 > sv) sv = v2;
                    if (v3 > bv) { sv = bv; bv = v3; bc = d + 3; } else if (v3 > sv) sv = v3;
                }
#pragma unroll 1
                for (int d = C4; d < C; ++d) {
                    const float v = -cr[d] - price[d];
                    if (v > bv) { sv = bv; bv = v; bc = d; } else if (v > sv) sv = v;
                }
                if (bc >= 0) bid = __float_as_uint(bv - sv + eps);                         // :99
            }
            const unsigned pm = __ballot_sync(FULL, bc >= 0);
            any = pm != 0u;
            if (any) {
                if (bc >= 0) atomicMax(&colbid[bc], bid);
                __syncwarp();
                if (bc >= 0 && colbid[bc] == bid) atomicMin(&colrow[bc], lane);
                __syncwarp();
                const bool win = bc >= 0 && colbid[bc] == bid && colrow[bc] == lane;
                __syncwarp();
                int prev = -1;
                if (win) {                                                                 // :107-121
                    colbid[bc] = 0u; colrow[bc] = 0x7fffffff;
                    prev = owner[bc];
                    owner[bc] = lane;
                    price[bc] += __uint_as_float(bid);
                }
                const unsigned wm = __ballot_sync(FULL, win);
                const unsigned em = __reduce_or_sync(FULL, prev >= 0 ? (1u << prev) : 0u); // evicted owners bid again
                ub = (pm & ~wm) | em;                                                      // rows without a column leave for good
                __syncwarp();
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if (lane + 32 * c < C) { p[c] = price[lane + 32 * c]; own[c] = owner[lane + 32 * c]; }
            }
        }
        if (!any) break;                                                                   // no bid: fixed point
        eps *= 0.9f;                                                                       // :402
        ++it;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c)
        if (own[c] >= 0) { const int slot = act_list[own[c]]; col[lane + 32 * c] = slot; row[slot] = lane + 32 * c; }
}

// ---------------------------------------------------------------------------------------
// Wide single-warp solve: up to 128 active rows and 128 columns (the crowd tables: 80 tracks x 85 detections), where the
// CTA-wide solve above pays two block barriers per iteration for the one to four rows that still bid after the first few
// iterations (measured: 85 us per solve).  Same decisions as auction_solve_cta bit for bit; the mapping of the lean solve:
// lane = column (NC columns per lane, prices and owners in registers), the bidder set a warp-uniform mask of NW words,
// bidders evaluated one after the other (two at a time: independent chains) with the arithmetic of lean_bid, per-lane
// highest bid (ascending rows + '>' = lowest row among equal bids), winners / evicted owners toggled through one
// REDUX.OR per mask word.  The first iteration takes the caller's first bids (lean_first_bid, one warp per row).
// cc: compacted cost rows [na, C] (position in act_list); pre_bc / pre_bits: first bids by position.
// ---------------------------------------------------------------------------------------
template <int NC>
__device__ __forceinline__ void lean_values_n(const float* cc, int C, int j, const float (&p)[NC], int lane, float& bv, float& sv, int& bsel) {
    bv = -1e9f; sv = -1e9f; bsel = 0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int d = lane + 32 * c;
        const float v = (d < C) ? fmaxf(-cc[j * C + d] - p[c], -1e9f) : -1e9f;              // :61
        if (c == 0) bv = v;
        else if (v > bv) { sv = bv; bv = v; bsel = c; }                                    // ascending columns: lowest column on ties (:63)
        else if (v > sv) sv = v;
    }
}
template <int NC>
__device__ __forceinline__ LeanBid lean_bid_n(float bv, float sv, int bsel, const int (&own)[NC], int lane, float eps) {
    const unsigned FULL = 0xffffffffu;
    LeanBid r;
    r.m = warp_max_f32(bv);
    const bool top = (bv == r.m);
    int osel = own[0];
#pragma unroll
    for (int c = 1; c < NC; ++c) osel = (bsel == c) ? own[c] : osel;
    const int key = top ? (((lane + 32 * bsel) << 8) | (osel + 1)) : 0x7fffffff;            // owner + 1 <= 128
    r.rk = __reduce_min_sync(FULL, key);
    const float m2p = warp_max_f32(top ? sv : bv);
    const unsigned ties = __ballot_sync(FULL, top);
    const float m2 = (ties & (ties - 1u)) ? r.m : m2p;
    r.bits = __float_as_uint(r.m - m2 + eps);                                              // :99
    return r;
}
template <int NC>
__device__ __forceinline__ void lean_first_bid_n(const float* cr, int C, int R, int lane, int& bc, unsigned& bits) {
    float p0[NC];
    int own0[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { p0[c] = 0.0f; own0[c] = -1; }
    float bv, sv;
    int bsel;
    lean_values_n<NC>(cr, C, 0, p0, lane, bv, sv, bsel);
    const LeanBid q = lean_bid_n<NC>(bv, sv, bsel, own0, lane, 1.0f / (float)(R + 1));
    bc = (q.m != -1e9f) ? (q.rk >> 8) : -1;
    bits = q.bits;
}

constexpr int WIDE_NW = 4, WIDE_NC = 4;            // 128 rows, 128 columns

__device__ __forceinline__ void wide_mask_xor(unsigned (&m)[WIDE_NW], int r, bool on) {   // r warp-uniform
    const unsigned bit = on ? (1u << (r & 31)) : 0u;
    const int w = r >> 5;
#pragma unroll
    for (int k = 0; k < WIDE_NW; ++k) m[k] ^= (w == k) ? bit : 0u;
}
__device__ __forceinline__ int wide_next(unsigned (&rem)[WIDE_NW]) {                       // lowest set bit, cleared; -1: none
#pragma unroll
    for (int k = 0; k < WIDE_NW; ++k)
        if (rem[k]) { const int j = 32 * k + __ffs(rem[k]) - 1; rem[k] &= rem[k] - 1u; return j; }
    return -1;
}

// One iteration with NB (1..4) bidders j[0] < ... < j[NB-1]: the structure of lean_iter — every lane ends up with the same
// (column, bid, evicted owner) of every bidder, so winners (highest bid, lowest row, :100), the bidder mask and the price /
// owner updates are warp-uniform arithmetic: no reduction, no shared-memory traffic besides the cost rows.
template <int NB>
__device__ __forceinline__ bool wide_iter(const float* cc, int C, unsigned (&ub)[WIDE_NW], const int (&j)[NB], float eps,
                                          float (&p)[WIDE_NC], int (&own)[WIDE_NC], int lane) {
    constexpr int NC = WIDE_NC;
    float bv[NB], sv[NB];
    int bs[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) lean_values_n<NC>(cc, C, j[b], p, lane, bv[b], sv[b], bs[b]);
    LeanBid q[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) q[b] = lean_bid_n<NC>(bv[b], sv[b], bs[b], own, lane, eps);
    int bc[NB];
    bool vm[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) { vm[b] = q[b].m != -1e9f; bc[b] = vm[b] ? (q[b].rk >> 8) : (-1 - b); }
    bool anyw = false;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        bool lose = false;
#pragma unroll
        for (int o = 0; o < NB; ++o) {
            if (o < b) lose |= (bc[o] == bc[b] && q[o].bits >= q[b].bits);
            if (o > b) lose |= (bc[o] == bc[b] && q[o].bits > q[b].bits);
        }
        const bool win = vm[b] && !lose;
        const int ow = (q[b].rk & 0xff) - 1;                                               // the column's owner so far (-1: none)
        wide_mask_xor(ub, j[b], win || !vm[b]);                                            // winners and rows without a column leave
        wide_mask_xor(ub, ow < 0 ? 0 : ow, win && ow >= 0);                                // evicted owners enter (:109-111)
        anyw |= win;
#pragma unroll
        for (int c = 0; c < NC; ++c)
            if (win && bc[b] == lane + 32 * c) { p[c] += __uint_as_float(q[b].bits); own[c] = j[b]; }   // :113-118
    }
    return anyw;
}

// colkey: [C] the first iteration's highest bid per column, (bid bits << 32) | ~row position (0: no bid), gathered by the
// caller with atomicMax while it computed the first bids (lean_first_bid_n); pre_bc: [na] first-bid columns (-1: none).
static __device__ __noinline__ void auction_solve_wide(const float* cc, int R, int C, const int* act_list, int na,
                                                       int* row, int* col, const unsigned (&ub0)[WIDE_NW],
                                                       const int* pre_bc, const unsigned long long* colkey) {
    constexpr int NC = WIDE_NC, NW = WIDE_NW;
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    for (int t = lane; t < R; t += 32) row[t] = -1;
    for (int d = lane; d < C; d += 32) col[d] = -1;
    if (na <= 0 || C <= 0) return;
    float p[NC];
    int own[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) { p[c] = 0.0f; own[c] = -1; }
    unsigned ub[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const int lo = 32 * w;
        const unsigned all = (na >= lo + 32) ? FULL : (na > lo ? ((1u << (na - lo)) - 1u) : 0u);
        ub[w] = all & ub0[w];
    }
    __syncwarp();
    float eps = 1.0f / (float)(R + 1);                                                     // :378
    const int iters = (R * 3 < 50) ? R * 3 : 50;                                           // :379
    int it = 0;
    if ((ub[0] | ub[1] | ub[2] | ub[3]) != 0u && iters > 0) {
        // ---- iteration 0: the caller's first bids, already reduced per column; nobody is evicted yet ----
        unsigned any = 0u;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int d = lane + 32 * c;
            const unsigned long long key = (d < C) ? colkey[d] : 0ull;
            if (key != 0ull) { own[c] = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull)); p[c] += __uint_as_float((unsigned)(key >> 32)); }
        }
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int i = 32 * w + lane;
            const bool in = (i < na) && ((ub[w] >> lane) & 1u);
            const int bci = in ? pre_bc[i] : -2;
            unsigned won = 0u;                                                             // rows that won a column: position bits of this word
#pragma unroll
            for (int c = 0; c < NC; ++c) won |= (own[c] >= 0 && (own[c] >> 5) == w) ? (1u << (own[c] & 31)) : 0u;
            won = __reduce_or_sync(FULL, won);
            const unsigned nobid = __ballot_sync(FULL, bci == -1);                         // no column above -1e9: leave for good
            ub[w] &= ~(won | nobid);
            any |= won;
        }
        if (any == 0u) { ub[0] = 0u; ub[1] = 0u; ub[2] = 0u; ub[3] = 0u; }                 // no bid: fixed point
        eps *= 0.9f;                                                                       // :402
        it = 1;
    }
#pragma unroll 1
    for (; it < iters; ++it) {
        const int nb = __popc(ub[0]) + __popc(ub[1]) + __popc(ub[2]) + __popc(ub[3]);
        if (nb == 0) break;                                                                // everybody assigned or out: fixed point
        bool any;
        if (nb <= 4) {
            unsigned rem[NW] = {ub[0], ub[1], ub[2], ub[3]};
            if (nb == 1) { const int j[1] = {wide_next(rem)}; any = wide_iter<1>(cc, C, ub, j, eps, p, own, lane); }
            else if (nb == 2) { int j[2]; j[0] = wide_next(rem); j[1] = wide_next(rem); any = wide_iter<2>(cc, C, ub, j, eps, p, own, lane); }
            else if (nb == 3) { int j[3]; j[0] = wide_next(rem); j[1] = wide_next(rem); j[2] = wide_next(rem); any = wide_iter<3>(cc, C, ub, j, eps, p, own, lane); }
            else { int j[4]; j[0] = wide_next(rem); j[1] = wide_next(rem); j[2] = wide_next(rem); j[3] = wide_next(rem); any = wide_iter<4>(cc, C, ub, j, eps, p, own, lane); }
        } else {
            // many bidders: one after the other, per-lane highest bid of the lane's columns (ascending rows + '>' = lowest
            // row among equal bids), winners / evicted owners toggled through one REDUX.OR per mask word
            unsigned best[NC];
            int wr[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) { best[c] = 0u; wr[c] = -1; }
            unsigned drop[NW] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                unsigned rem = ub[w];
#pragma unroll 1
                while (rem) {
                    const int jb = 32 * w + __ffs(rem) - 1;
                    rem &= rem - 1;
                    float bv, sv;
                    int bs;
                    lean_values_n<NC>(cc, C, jb, p, lane, bv, sv, bs);
                    const LeanBid q = lean_bid_n<NC>(bv, sv, bs, own, lane, eps);
                    if (q.m == -1e9f) { drop[w] |= 1u << (jb & 31); continue; }            // no column above -1e9: leaves for good
#pragma unroll
                    for (int c = 0; c < NC; ++c)
                        if ((q.rk >> 8) == lane + 32 * c && q.bits > best[c]) { best[c] = q.bits; wr[c] = jb; }
                }
            }
            unsigned tog[NW] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (wr[c] >= 0) {                                                          // :107-121
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        if ((wr[c] >> 5) == w) tog[w] |= 1u << (wr[c] & 31);
                        if (own[c] >= 0 && (own[c] >> 5) == w) tog[w] |= 1u << (own[c] & 31);
                    }
                    own[c] = wr[c];
                    p[c] += __uint_as_float(best[c]);
                }
            unsigned anyb = 0u;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                tog[w] = __reduce_or_sync(FULL, tog[w]);
                anyb |= tog[w];
                ub[w] = (ub[w] & ~drop[w]) ^ tog[w];                                       // winners leave, evicted owners enter
            }
            any = anyb != 0u;
        }
        if (!any) break;                                                                   // no bid: fixed point
        eps *= 0.9f;                                                                       // :402
    }
#pragma unroll
    for (int c = 0; c < NC; ++c)
        if (own[c] >= 0) { const int slot = act_list[own[c]]; col[lane + 32 * c] = slot; row[slot] = lane + 32 * c; }
}

}  // namespace pb
