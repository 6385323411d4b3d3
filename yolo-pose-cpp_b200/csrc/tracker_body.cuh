// tracker_body.cuh — the PoseBYTE tracker update of one stream, executed by one CTA (device code shared by the
// stand-alone tracker kernel, tracker.cu, and the fused per-stream kernel, fused.cu).
//
//
// Replaces GPUTracker::update + getActiveTracks (reference src/cuda/gpu_tracker.cu:
// 1057-1158, 1559-1639) — ~490 stream operations and 2 host synchronisations per frame
// and stream upstream — with ONE launch for all streams.  The stage order, every
// formula and every persistent buffer follow the reference literally (SURVEY.md §8a rows
// A5-A16, quirks Q1-Q8); the stages run back to back inside the CTA with the working set
// (cost matrix, detections, predicted poses, gate bitmasks, assignments, prices) in
// shared memory:
//   predict (:102-138) -> keypoint-box centres (:196-237) -> velocity-adaptive spatial
//   gate (:241-317), bit-packed -> tier 1: visibility-masked OKS (:333-425) + auction +
//   lock (:540-567) -> tier 2: torso OKS (:429-490) + auction + merge (:575-588) + lock
//   -> tier 3: lost-track gate x1.3, OKS, auction, merge -> constant-gain update
//   (:141-189, :612-648) -> ageing (:651-688) -> new tracks (:695-780, rules R3/R4) ->
//   IoU de-duplication (:788-895, rule R5) -> TrackOutput assembly (:1594-1636).
// The auction (hungarian.cu:27-123, 358-405) keeps the reference's bid arithmetic and
// tie-breaks (R6); one warp runs the whole solve with lane = column, prices and owners in
// registers (auction.cuh), and the loop stops at the first iteration without a bid — a
// fixed point of the reference's 50 fixed iterations.
#pragma once
#include "pb_common.cuh"
#include "auction.cuh"

namespace pb {


constexpr unsigned FULLM = 0xffffffffu;
constexpr int DUP_CAP = 256;

struct TkSmem {
    int *active, *states, *hits, *ids, *ages, *row, *rowb, *act_list, *elig_list, *rowbc, *rowbid;   // [T]
    int *col, *colb, *slot_for_det, *out_list;                                        // [Dm]
    float *price, *dscore, *darea;                                                    // [Dm]
    unsigned long long* colbid;                                                       // [Dm]
    float *tcent, *tarea, *tav;                                                       // [T*4],[T],[T]
    float* dcent;                                                                     // [Dm*4]
    unsigned *gate, *lgate;                                                           // [T*Dw]
    unsigned* colmask;                                                                // [Dw]
    int* dup;                                                                         // [DUP_CAP]
    int* misc;                                                                        // [32]
    unsigned long long* acc;                                                          // [20] telemetry
    float* terms;                                                                     // [term_floats]
    float* sig;                                                                       // [17]
    int* cell_list;                                                                   // [cell_cap] gated cells compacted per chunk of active rows
    int* aowner;                                                                      // [Dm] auction scratch
    float *cost, *det, *pred;                                                         // optional
    float *poses, *vel; int* dirty;                                                   // optional (resident tracker): [T*51], [T*34], [T]
};

__host__ __device__ inline size_t tk_align(size_t x) { return (x + 15) & ~(size_t)15; }

// The arrays that receive this frame's detections (scores, poses) come first: in the fused per-stream kernel they are
// written by the NMS stage while the rest of the layout is still the NMS stage's own (`prefix`: their size).
__host__ __device__ inline size_t tk_carve(unsigned char* base, int T, int Dm, int cost_s, int det_s,
                                           int pred_s, int term_floats, int cell_cap, TkSmem* s, size_t* prefix = nullptr, int res_s = 0) {
    const int Dw = (Dm + 31) / 32;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = tk_align(off + bytes); return o; };
    size_t o_dscore = take((size_t)Dm * 4);
    size_t o_det = det_s ? take((size_t)Dm * POSE_F * 4) : 0;
    if (prefix) *prefix = off;
    size_t o_colbid = take((size_t)Dm * 8), o_acc = take(20 * 8);
    size_t o_i[11]; for (int i = 0; i < 11; ++i) o_i[i] = take((size_t)T * 4);
    size_t o_d[4]; for (int i = 0; i < 4; ++i) o_d[i] = take((size_t)Dm * 4);
    size_t o_f[2]; for (int i = 0; i < 2; ++i) o_f[i] = take((size_t)Dm * 4);
    size_t o_tcent = take((size_t)T * 16), o_tarea = take((size_t)T * 4), o_tav = take((size_t)T * 4);
    size_t o_dcent = take((size_t)Dm * 16);
    size_t o_gate = take((size_t)T * Dw * 4), o_lgate = take((size_t)T * Dw * 4), o_colmask = take((size_t)Dw * 4);
    size_t o_dup = take(DUP_CAP * 4), o_misc = take(32 * 4);
    size_t o_terms = take((size_t)term_floats * 4), o_sig = take(KP * 4), o_cl = take((size_t)cell_cap * 4);
    size_t o_aown = take((size_t)Dm * 4);
    size_t o_cost = cost_s ? take((size_t)T * Dm * 4) : 0;
    size_t o_pred = pred_s ? take((size_t)T * POSE_F * 4) : 0;
    // resident tracker: the stream's poses, velocities and dirty flags stay in shared memory over the frames of a launch
    size_t o_poses = res_s ? take((size_t)T * POSE_F * 4) : 0, o_vel = res_s ? take((size_t)T * 34 * 4) : 0, o_dirty = res_s ? take((size_t)T * 4) : 0;
    if (s) {
        s->colbid = (unsigned long long*)(base + o_colbid);
        s->acc = (unsigned long long*)(base + o_acc);
        int** ip[11] = {&s->active, &s->states, &s->hits, &s->ids, &s->ages, &s->row, &s->rowb, &s->act_list, &s->elig_list,
                        &s->rowbc, &s->rowbid};
        for (int i = 0; i < 11; ++i) *ip[i] = (int*)(base + o_i[i]);
        int** dp[4] = {&s->col, &s->colb, &s->slot_for_det, &s->out_list};
        for (int i = 0; i < 4; ++i) *dp[i] = (int*)(base + o_d[i]);
        float** fp[2] = {&s->price, &s->darea};
        for (int i = 0; i < 2; ++i) *fp[i] = (float*)(base + o_f[i]);
        s->dscore = (float*)(base + o_dscore);
        s->tcent = (float*)(base + o_tcent); s->tarea = (float*)(base + o_tarea); s->tav = (float*)(base + o_tav);
        s->dcent = (float*)(base + o_dcent);
        s->gate = (unsigned*)(base + o_gate); s->lgate = (unsigned*)(base + o_lgate);
        s->colmask = (unsigned*)(base + o_colmask);
        s->dup = (int*)(base + o_dup); s->misc = (int*)(base + o_misc);
        s->terms = (float*)(base + o_terms); s->sig = (float*)(base + o_sig); s->cell_list = (int*)(base + o_cl);
        s->aowner = (int*)(base + o_aown);
        s->cost = cost_s ? (float*)(base + o_cost) : nullptr;
        s->det = det_s ? (float*)(base + o_det) : nullptr;
        s->pred = pred_s ? (float*)(base + o_pred) : nullptr;
        s->poses = res_s ? (float*)(base + o_poses) : nullptr;
        s->vel = res_s ? (float*)(base + o_vel) : nullptr;
        s->dirty = res_s ? (int*)(base + o_dirty) : nullptr;
    }
    return off;
}


// Shared-memory layout as byte offsets, computed once on the host (tracker_plan) and handed to the kernel as a
// launch parameter: the 36 pointers of TkSmem are then one constant-bank add each wherever the compiler
// rematerialises them, instead of the offset arithmetic of tk_carve (measured: ~1000 SASS instructions).
__device__ __forceinline__ void tk_from_offsets(unsigned char* base, const SmemOffsets& o, TkSmem& s) {
    s.active = reinterpret_cast<int*>(base + o.off[0]);
    s.states = reinterpret_cast<int*>(base + o.off[1]);
    s.hits = reinterpret_cast<int*>(base + o.off[2]);
    s.ids = reinterpret_cast<int*>(base + o.off[3]);
    s.ages = reinterpret_cast<int*>(base + o.off[4]);
    s.row = reinterpret_cast<int*>(base + o.off[5]);
    s.rowb = reinterpret_cast<int*>(base + o.off[6]);
    s.act_list = reinterpret_cast<int*>(base + o.off[7]);
    s.elig_list = reinterpret_cast<int*>(base + o.off[8]);
    s.rowbc = reinterpret_cast<int*>(base + o.off[9]);
    s.rowbid = reinterpret_cast<int*>(base + o.off[10]);
    s.col = reinterpret_cast<int*>(base + o.off[11]);
    s.colb = reinterpret_cast<int*>(base + o.off[12]);
    s.slot_for_det = reinterpret_cast<int*>(base + o.off[13]);
    s.out_list = reinterpret_cast<int*>(base + o.off[14]);
    s.price = reinterpret_cast<float*>(base + o.off[15]);
    s.dscore = reinterpret_cast<float*>(base + o.off[16]);
    s.darea = reinterpret_cast<float*>(base + o.off[17]);
    s.colbid = reinterpret_cast<unsigned long long*>(base + o.off[18]);
    s.tcent = reinterpret_cast<float*>(base + o.off[19]);
    s.tarea = reinterpret_cast<float*>(base + o.off[20]);
    s.tav = reinterpret_cast<float*>(base + o.off[21]);
    s.dcent = reinterpret_cast<float*>(base + o.off[22]);
    s.gate = reinterpret_cast<unsigned*>(base + o.off[23]);
    s.lgate = reinterpret_cast<unsigned*>(base + o.off[24]);
    s.colmask = reinterpret_cast<unsigned*>(base + o.off[25]);
    s.dup = reinterpret_cast<int*>(base + o.off[26]);
    s.misc = reinterpret_cast<int*>(base + o.off[27]);
    s.acc = reinterpret_cast<unsigned long long*>(base + o.off[28]);
    s.terms = reinterpret_cast<float*>(base + o.off[29]);
    s.sig = reinterpret_cast<float*>(base + o.off[30]);
    s.cell_list = reinterpret_cast<int*>(base + o.off[31]);
    s.aowner = reinterpret_cast<int*>(base + o.off[32]);
    s.cost = reinterpret_cast<float*>(base + o.off[33]);
    s.det = reinterpret_cast<float*>(base + o.off[34]);
    s.pred = reinterpret_cast<float*>(base + o.off[35]);
    s.poses = reinterpret_cast<float*>(base + o.off[36]);
    s.vel = reinterpret_cast<float*>(base + o.off[37]);
    s.dirty = reinterpret_cast<int*>(base + o.off[38]);
}
// ---------------------------------------------------------------------------------------
// keypoint-box statistics of one pose (17 lanes of a warp would be overkill: T+D poses)
// centres as kernelComputeBboxCenters (:196-237); area as the scale term of
// kernelOKSWithGating (:364-392).  Both use keypoints with conf > 0.1.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void pose_box(const float* p, float* cent4, float* area) {
    float lx = 1e9f, ly = 1e9f, hx = -1e9f, hy = -1e9f;
    int valid = 0;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        if (p[k * 3 + 2] > 0.1f) {
            const float x = p[k * 3], y = p[k * 3 + 1];
            lx = pb_min(lx, x); ly = pb_min(ly, y); hx = pb_max(hx, x); hy = pb_max(hy, y);
            ++valid;
        }
    }
    *area = (hx - lx) * (hy - ly);
    if (valid < 2) { cent4[0] = 0.f; cent4[1] = 0.f; cent4[2] = 0.f; cent4[3] = 0.f; return; }
    const float w = hx - lx, h = hy - ly;
    cent4[0] = (lx + hx) * 0.5f; cent4[1] = (ly + hy) * 0.5f; cent4[2] = w; cent4[3] = h;
}

// kernelSpatialGate (:241-317) for one active row / one detection.
__device__ __forceinline__ bool gate_cell(const float* tc, const float* dc, float av, bool lost,
                                          float base, bool gating) {
    const float tw = tc[2], th = tc[3], dw = dc[2], dh = dc[3];
    if (tw < 1.0f || th < 1.0f || dw < 1.0f || dh < 1.0f) return true;
    if (!gating) return true;
    const float dx = tc[0] - dc[0], dy = tc[1] - dc[1];
    const float size = (tw + th + dw + dh) * 0.25f;
    const float den = size + 1e-6f;
    // Cheap exact rejection (most pairs of a crowded frame): the threshold below is at most base * 3 (* 2 for a lost track)
    // and the ratio at least max(|dx|, |dy|) / den up to three roundings, so a centre offset beyond that bound by a margin
    // of 1e-5 cannot pass the test.  NaNs fail both comparisons and take the complete test.
    const float lim = (base * 3.0f) * (lost ? 2.0f : 1.0f) * den * 1.00001f;
    if (fabsf(dx) > lim || fabsf(dy) > lim) return false;
    const float dist = sqrtf(dx * dx + dy * dy);
    const float ratio = dist / den;
    const float vf = 1.0f + pb_min(av / den, 2.0f);
    float thr = base * vf;
    if (lost) thr *= 2.0f;
    return ratio < thr;
}

// kernelTorsoOKS cell (:455-489).
__device__ __forceinline__ float torso_cost(const float* tp, const float* dp) {
    const int torso[4] = {5, 6, 11, 12};
    const float scale_sq = 10000.0f;
    float sum = 0.0f;
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = torso[i];
        if (dp[k * 3 + 2] > 0.1f && tp[k * 3 + 2] > 0.1f) {
            const float dx = dp[k * 3] - tp[k * 3], dy = dp[k * 3 + 1] - tp[k * 3 + 1];
            const float d2 = dx * dx + dy * dy;
            const float sg = kSigmas[k] * 3.0f;
            sum += pb_expf(-d2 / (2.0f * scale_sq * sg * sg));
            ++cnt;
        }
    }
    const float oks = (cnt >= 2) ? (sum / (float)cnt) : 0.0f;
    return 1.0f - oks;
}

struct Ctx {
    TkSmem s;
    int T, D, Dw, tid, nthreads, lane, warp, nwarps;
    unsigned magicD, magicW;   // ceil(2^32 / D), ceil(2^32 / words): i / D == __umulhi(i, magicD) for i * D < 2^32 (0: divisor 1)
    bool warp_auction;  // small problem with the cost matrix in shared memory: single-warp solve
    const int* od_flag; int od_want;   // (may be null) the predecessor's second release flag and the value to see: polled by an idle warp during the tier-1 solve
    bool sub_solve;     // tiers 2 and 3 of larger tables: solve the unmatched rows x unmatched detections only (auction_solve_unlocked)
    int term_floats, cell_cap;
    float* cost;        // shared or global, flat [t*D + d]
    const float* det;   // shared or global scratch [d*51]
    float* pred;        // shared or global (persistent) [t*51]
};

// i / d for the small non-negative indices of the stage loops, d fixed per frame: one multiply instead of the
// ~25-instruction division sequence (magic = ceil(2^32 / d), exact while i * d < 2^32; d == 1 gives magic 0).
__device__ __forceinline__ int fast_div(int i, unsigned magic) { return magic ? (int)__umulhi((unsigned)i, magic) : i; }
__device__ __forceinline__ unsigned div_magic(int d) { return d > 1 ? (0xffffffffu / (unsigned)d + 1u) : 0u; }

// Tiers 2 and 3 on tables beyond the lean solve's own case (crowd tables, the 512 x 512 stress tables).  Rows matched in
// an earlier tier are locked — all their cells are 1e9 (lock_pairs), they never bid — and so are the matched columns: their
// cells are 1e9 for every active row, so for a bidder they sit at the -1e9 floor, never best and never second
// (hungarian.cu:55-69).  What is left is the auction of the unmatched active rows over the unmatched detections, usually a
// handful of each: same eps0 = 1/(T+1), same iteration limit, same tie-breaks (both index maps are ascending).  With at most
// 32 such rows and 64 such columns it is compacted into the idle term buffer and solved by the lean single-warp solve
// instead of the CTA-wide one (two block barriers and a pass over all columns per iteration; an eviction chase among the
// left-over rows runs to the 50-iteration limit: 68 + 30 us per 512 x 512 frame).  false: too many left, nothing was done.
struct SubSolveArgs {        // by value: a reference to the CTA's context would force it into local memory
    const float* cost; int T, D, na, tid, nthreads;
    const int *act_list, *rowb, *colb;
    int *row, *col, *cell_list, *rowbc, *aowner;
    float *terms, *price;
    unsigned long long* colbid;
};
static __device__ __noinline__ bool auction_solve_unlocked(const SubSolveArgs a) {
    int* sub_act = a.cell_list;                 // [32] unmatched active rows (slots, ascending); the cell list is idle between cost passes
    int* sub_col = a.cell_list + 32;            // [64] unmatched detections (ascending)
    int* sub_own = a.cell_list + 96;            // [64] the solve's column -> slot
    int* cnt = a.cell_list + 160;               // [2]
    const int D = a.D, lane = a.tid & 31;
    if (a.tid < 32) {
        int nr = 0, nc = 0;
#pragma unroll 1
        for (int a0 = 0; a0 < a.na; a0 += 32) {
            const int ai = a0 + lane;
            const int t = ai < a.na ? a.act_list[ai] : 0;
            const bool un = ai < a.na && a.rowb[t] < 0;
            const unsigned bm = __ballot_sync(FULLM, un);
            const int pos = nr + __popc(bm & ((1u << lane) - 1u));
            if (un && pos < 32) sub_act[pos] = t;
            nr += __popc(bm);
        }
#pragma unroll 1
        for (int d0 = 0; d0 < D; d0 += 32) {
            const int d = d0 + lane;
            const bool un = d < D && a.colb[d] < 0;
            const unsigned bm = __ballot_sync(FULLM, un);
            const int pos = nc + __popc(bm & ((1u << lane) - 1u));
            if (un && pos < 64) sub_col[pos] = d;
            nc += __popc(bm);
        }
        if (lane == 0) { cnt[0] = nr; cnt[1] = nc; }
    }
    __syncthreads();
    const int nr = cnt[0], nc = cnt[1];
    if (nr > 32 || nc > 64) return false;       // (uniform over the CTA)
#pragma unroll 1
    for (int t = a.tid; t < a.T; t += a.nthreads) a.row[t] = -1;                            // :372-376
#pragma unroll 1
    for (int d = a.tid; d < D; d += a.nthreads) a.col[d] = -1;
    if (nr == 0 || nc == 0) { __syncthreads(); return true; }
    float* cc = a.terms;
#pragma unroll 1
    for (int i = a.tid; i < nr * nc; i += a.nthreads) {
        const int r = i / nc, j = i - r * nc;
        cc[i] = a.cost[(size_t)sub_act[r] * D + sub_col[j]];
    }
    __syncthreads();
    if (a.tid < 32) {
        unsigned* cb = reinterpret_cast<unsigned*>(a.colbid);
        int* cr = reinterpret_cast<int*>(a.colbid) + nc;
        // (row scratch: rowbc, idle since the centre stage; the solve clears and fills it by slot)
        if (nc <= 32) auction_solve_lean32<1, 1>(cc, a.T, nc, sub_act, nr, a.rowbc, sub_own, a.price, a.aowner, cb, cr);
        else auction_solve_lean32<2, 1>(cc, a.T, nc, sub_act, nr, a.rowbc, sub_own, a.price, a.aowner, cb, cr);
        __syncwarp();
        for (int j = lane; j < nc; j += 32) {
            const int slot = sub_own[j];
            if (slot >= 0) { a.col[sub_col[j]] = slot; a.row[slot] = sub_col[j]; }
        }
    }
    __syncthreads();
    return true;
}

// Auction solve: see auction.cuh.  Leaves row/col in s.row / s.col.
__device__ __forceinline__ void auction_solve(Ctx& c, int na, bool after_lock) {
    TkSmem& s = c.s;
    const bool lean_case = c.warp_auction && na <= 32 && c.D <= 64;
    if (!lean_case && after_lock && c.sub_solve) {
        const SubSolveArgs a{c.cost, c.T, c.D, na, c.tid, c.nthreads, s.act_list, s.rowb, s.colb, s.row, s.col, s.cell_list, s.rowbc, s.aowner,
                             s.terms, s.price, s.colbid};
        if (auction_solve_unlocked(a)) return;
    }
    if (lean_case) {
        if (after_lock) {
            // Rows matched in an earlier tier are locked (all their cells are 1e9, lock_pairs) and can never bid.  When
            // that is every active row — the usual frame: everybody found its detection in tier 1 — the solve would
            // clear the assignments and stop before its first iteration: do just that, without compacting the rows.
            const bool may = c.lane < na && s.rowb[s.act_list[c.lane]] < 0;
            if (__ballot_sync(FULLM, may) == 0u) {                 // the same in every warp
#pragma unroll 1
                for (int t = c.tid; t < c.T; t += c.nthreads) s.row[t] = -1;
#pragma unroll 1
                for (int d = c.tid; d < c.D; d += c.nthreads) s.col[d] = -1;
                __syncthreads();
                return;
            }
        }
        // compact the active rows (cc[i*D + d], i = position in act_list) into the term buffer, which is idle
        // between cost passes: the single-warp solve then reads a bidder's row with one conflict-free load
        float* cc = s.terms;
        const int D = c.D;
#pragma unroll 1
        for (int i = c.tid; i < na * D; i += c.nthreads) { const int ai = fast_div(i, c.magicD); cc[i] = s.cost[s.act_list[ai] * D + (i - ai * D)]; }
        // ... and the first iteration's bids (prices 0, nobody assigned: a row's own costs decide), one warp per active row
        int* pre_bc = s.rowbc;                                       // [T], idle since the centre stage
        unsigned* pre_bits = reinterpret_cast<unsigned*>(s.rowbid);
#pragma unroll 1
        for (int ai = c.warp; ai < na; ai += c.nwarps) {
            const float* cr0 = s.cost + (size_t)s.act_list[ai] * D;
            int bc;
            unsigned bits;
            if (D <= 32) lean_first_bid<1>(cr0, D, c.T, c.lane, bc, bits); else lean_first_bid<2>(cr0, D, c.T, c.lane, bc, bits);
            if (c.lane == 0) { pre_bc[ai] = bc; pre_bits[ai] = bits; }
        }
        if (!after_lock && c.od_flag != nullptr && c.tid == c.nthreads - 32) {
            // The last warp has no row here: it looks at the predecessor's SECOND release (records assembled, g_poses no longer
            // read), which the update stage needs — an L2 round trip that would otherwise sit on the chain behind the tiers.
            int v;
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(c.od_flag) : "memory");
            if (v - c.od_want >= 0) s.misc[15] = 1;
        }
        __syncthreads();
        if (c.tid < 32) {
            unsigned* cb = reinterpret_cast<unsigned*>(s.colbid);
            int* cr = reinterpret_cast<int*>(s.colbid) + c.D;
            // rows matched in an earlier tier are locked: all their cells are 1e9 (lock_pairs), they can never bid
            const bool may_bid = c.tid < na && !(after_lock && s.rowb[s.act_list[c.tid]] >= 0);
            const unsigned ub0 = __ballot_sync(FULLM, may_bid);
            if (D <= 32) auction_solve_lean32<1>(cc, c.T, D, s.act_list, na, s.row, s.col, s.price, s.aowner, cb, cr, ub0, s.acc, pre_bc, pre_bits);
            else auction_solve_lean32<2>(cc, c.T, D, s.act_list, na, s.row, s.col, s.price, s.aowner, cb, cr, ub0, s.acc, pre_bc, pre_bits);
        }
        __syncthreads();
    } else if (na <= 32 * WIDE_NW && c.D <= 32 * WIDE_NC && na * c.D <= c.term_floats && !(c.warp_auction && na <= 32)) {
        // Crowd tables (up to 128 active rows x 128 detections): the wide single-warp solve.  The whole CTA compacts the
        // active rows into the idle term buffer and computes the first iteration's bids, one warp per row, reducing them
        // per column with a 64-bit atomicMax (highest bid, lowest row).
        const int D = c.D;
        int* pre_bc = s.rowbc;                                       // [T], idle since the centre stage
        float* cc = s.terms;
#pragma unroll 1
        for (int d = c.tid; d < D; d += c.nthreads) s.colbid[d] = 0ull;
#pragma unroll 1
        for (int i = c.tid; i < na * D; i += c.nthreads) { const int ai = fast_div(i, c.magicD); cc[i] = c.cost[(size_t)s.act_list[ai] * D + (i - ai * D)]; }
        __syncthreads();
#pragma unroll 1
        for (int ai = c.warp; ai < na; ai += c.nwarps) {
            if (after_lock && s.rowb[s.act_list[ai]] >= 0) { if (c.lane == 0) pre_bc[ai] = -1; continue; }   // locked: all cells 1e9, no bid
            int bc;
            unsigned bits;
            lean_first_bid_n<WIDE_NC>(cc + (size_t)ai * D, D, c.T, c.lane, bc, bits);
            if (c.lane == 0) {
                pre_bc[ai] = bc;
                if (bc >= 0) atomicMax(&s.colbid[bc], ((unsigned long long)bits << 32) | (unsigned long long)(0xffffffffu - (unsigned)ai));
            }
        }
        __syncthreads();
        if (c.tid < 32) {
            unsigned ub0[WIDE_NW];
#pragma unroll
            for (int w = 0; w < WIDE_NW; ++w) {
                const int i = 32 * w + c.lane;
                // rows matched in an earlier tier are locked: all their cells are 1e9 (lock_pairs), they can never bid
                ub0[w] = __ballot_sync(FULLM, i < na && !(after_lock && s.rowb[s.act_list[i]] >= 0));
            }
            auction_solve_wide(cc, c.T, D, s.act_list, na, s.row, s.col, ub0, pre_bc, s.colbid);
        }
        __syncthreads();
    } else if (c.warp_auction && na <= 32) {
        if (c.tid < 32)
            auction_solve_hybrid32(s.cost, c.T, c.D, s.act_list, na, s.row, s.col, s.price, s.aowner,
                                   reinterpret_cast<unsigned*>(s.colbid), reinterpret_cast<int*>(s.colbid) + c.D);
        __syncthreads();
    } else if (c.warp_auction) {
        // colbid (8 B per column) doubles as the 32-bit bid array + the lowest-row array
        unsigned* colbid32 = reinterpret_cast<unsigned*>(s.colbid);
        int* colrow = reinterpret_cast<int*>(s.colbid) + c.D;
        auction_solve_warp(c.cost, c.T, c.D, s.active, s.row, s.col, s.price, colbid32, colrow, s.rowbc,
                           reinterpret_cast<unsigned*>(s.rowbid), c.tid);
    } else {
        auction_solve_cta(c.cost, c.T, c.D, s.active, s.row, s.col, s.price, s.colbid, &s.misc[8],
                          c.tid, c.nthreads, -1, after_lock ? s.rowb : nullptr);
    }
}

// One gated cell of kernelOKSWithGating (:360-424), all 17 keypoints by one thread, keypoint-ordered sum: the expressions of
// cost_pass_oks below (which spreads the 17 terms of a cell over threads and adds them in the same order; an invisible
// keypoint adds -0.0f there, i.e. nothing).  Used by the row-sliced pre-kernel of large tables.
__device__ __forceinline__ float oks_cell_cost(const float* tp, const float* dp, float tarea, float darea, const float* sig, float vis) {
    const float scale_sq = pb_max((darea + tarea) * 0.5f, 1000.0f);
    const float t2 = 2.0f * scale_sq;
    float sum = 0.0f;
    int cnt = 0;
#pragma unroll 1
    for (int k = 0; k < KP; ++k) {
        if (dp[k * 3 + 2] > vis && tp[k * 3 + 2] > vis) {
            const float dx = dp[k * 3] - tp[k * 3], dy = dp[k * 3 + 1] - tp[k * 3 + 1];
            const float d2 = dx * dx + dy * dy;
            const float sg = sig[k] * 2.0f;
            const float s2 = sg * sg;
            sum += pb_expf(-d2 / (t2 * s2));
            ++cnt;
        }
    }
    const float oks = (cnt >= 3) ? (sum / (float)cnt) : 0.0f;
    return 1.0f - oks;
}

// kernelLockMatchedPairs (:540-567) on cost + a bit-packed gate.  Only rows that were active at
// frame start are touched: the cells of inactive rows are never read by the auction (inactive
// rows do not bid) and are overwritten with 1.0 by the last cost pass of the frame anyway.
static __device__ __forceinline__ void lock_pairs(Ctx& c, unsigned* gate, int na, bool use_backup) {
    // use_backup: rowb / colb hold the assignments of the earlier tiers of this frame, whose rows and columns are locked in
    // the cost matrix already (nothing writes those cells in between: their gate bits were cleared with them) — only the
    // pairs this tier added need the 1e9 writes; the gate words are cleared for all matched rows and columns either way.
    TkSmem& s = c.s;
    const int D = c.D, Dw = c.Dw, words = (D + 31) >> 5;
    if (c.warp_auction) {
        // small tables (cost matrix in shared memory): one pass, one barrier — a warp per active row, the column mask by ballot
#pragma unroll 1
        for (int ai = c.warp; ai < na; ai += c.nwarps) {
            const int t = s.act_list[ai];
            const bool rowm = s.row[t] >= 0;
            for (int w = 0; w < words; ++w) {
                const int d = w * 32 + c.lane;
                const bool colm = (d < D) && (s.col[d] >= 0);
                const unsigned bm = __ballot_sync(FULLM, colm);
                if (d < D && (rowm || colm)) c.cost[(size_t)t * D + d] = 1e9f;
                if (c.lane == 0) gate[t * Dw + w] = rowm ? 0u : (gate[t * Dw + w] & ~bm);
            }
        }
        __syncthreads();
        return;
    }
    // column masks, one pass: matched columns (colmask) and the ones whose cells still need the write (price buffer words are
    // idle between solves: reused as the second mask)
    unsigned* colnew = reinterpret_cast<unsigned*>(s.price);
#pragma unroll 1
    for (int w = c.warp; w < words; w += c.nwarps) {
        const int d = w * 32 + c.lane;
        const bool colm = (d < D) && (s.col[d] >= 0);
        const bool cold = colm && use_backup && (s.colb[d] >= 0);
        const unsigned bm = __ballot_sync(FULLM, colm), bo = __ballot_sync(FULLM, cold);
        if (c.lane == 0) { s.colmask[w] = bm; colnew[w] = bm & ~bo; }
    }
    __syncthreads();
    // gate words: one thread per (active row, word)
#pragma unroll 1
    for (int i = c.tid; i < na * words; i += c.nthreads) {
        const int ai = fast_div(i, c.magicW), w = i - ai * words;
        const int t = s.act_list[ai];
        gate[t * Dw + w] = (s.row[t] >= 0) ? 0u : (gate[t * Dw + w] & ~s.colmask[w]);
    }
    // cost cells: one warp per (active row that was not locked before, word with something to write)
#pragma unroll 1
    for (int i = c.warp; i < na * words; i += c.nwarps) {
        const int ai = fast_div(i, c.magicW), w = i - ai * words;
        const int t = s.act_list[ai];
        if (use_backup && s.rowb[t] >= 0) continue;                   // whole row locked by an earlier tier
        const unsigned valid = (D - w * 32 >= 32) ? 0xffffffffu : ((1u << (D - w * 32)) - 1u);
        const unsigned old_cols = use_backup ? (s.colmask[w] & ~colnew[w]) : 0u;   // columns locked by an earlier tier
        const unsigned m = (s.row[t] >= 0) ? (valid & ~old_cols) : colnew[w];
        if ((m >> c.lane) & 1u) c.cost[(size_t)t * D + w * 32 + c.lane] = 1e9f;
    }
    __syncthreads();
}

// cost <- 1.0 on inactive rows (:351-354): warp per row.  Only the last writer of a frame matters
// for these rows (see lock_pairs), so this runs once per frame, in the lost-track tier.
static __device__ __forceinline__ void cost_inactive_rows(Ctx& c) {
    TkSmem& s = c.s;
#pragma unroll 1
    for (int t = c.warp; t < c.T; t += c.nwarps)
        if (s.active[t] == 0)
#pragma unroll 1
            for (int d = c.lane; d < c.D; d += 32) c.cost[(size_t)t * c.D + d] = 1.0f;
}

// Visibility-masked OKS cost on the gated cells of the active rows (kernelOKSWithGating :360-424).
// The gated cells are first compacted into a list (warp per row, ballot append); then the
// exponentials are spread over threads: one thread per (cell, keypoint) evaluates a term into
// shared memory and one thread per cell adds the terms of its visible keypoints in keypoint
// order (the reference's summation order) and finishes the cell.
static __device__ __forceinline__ void cost_pass_oks(Ctx& c, const unsigned* gate, int na, float vis) {
    TkSmem& s = c.s;
    const int D = c.D, Dw = c.Dw, words = (D + 31) >> 5;
    int rows_per_chunk = c.cell_cap / D;            // worst case: every cell of the chunk is gated
    if (rows_per_chunk < 1) rows_per_chunk = 1;
    const int cells_per_round = c.term_floats / KP;
    for (int a0 = 0; a0 < na; a0 += rows_per_chunk) {
        const int a1 = (a0 + rows_per_chunk < na) ? a0 + rows_per_chunk : na;
        // (the first chunk finds the counter at zero: the prologue cleared it and every pass clears it again at its end,
        // behind barriers — one block barrier less on the chain of frames)
        if (a0 > 0) {
            if (c.tid == 0) s.misc[5] = 0;
            __syncthreads();
        }
#pragma unroll 1
        for (int ai = a0 + c.warp; ai < a1; ai += c.nwarps) {
            const int t = s.act_list[ai];
            for (int w = 0; w < words; ++w) {
                const int d = w * 32 + c.lane;
                const unsigned gw = gate[t * Dw + w];
                if (gw == 0u) continue;
                const int n = __popc(gw);
                int base = 0;
                if (c.lane == 0) base = atomicAdd(&s.misc[5], n);
                base = __shfl_sync(FULLM, base, 0);
                if ((gw >> c.lane) & 1u) {
                    const int pos = base + __popc(gw & ((1u << c.lane) - 1u));
                    if (pos < c.cell_cap) s.cell_list[pos] = (t << 16) | d;
                }
            }
        }
        __syncthreads();
        const int ncell = s.misc[5] < c.cell_cap ? s.misc[5] : c.cell_cap;
        for (int cb = 0; cb < ncell; cb += cells_per_round) {
            const int ncur = (ncell - cb) < cells_per_round ? (ncell - cb) : cells_per_round;
#pragma unroll 1
            for (int idx = c.tid; idx < ncur * KP; idx += c.nthreads) {
                const int e = idx / KP, k = idx - e * KP;
                const int key = s.cell_list[cb + e];
                const int t = key >> 16, d = key & 0xffff;
                const float* tp = c.pred + (size_t)t * POSE_F + k * 3;
                const float* dp = c.det + (size_t)d * POSE_F + k * 3;
                float term = -0.0f;                  // marker of an invisible keypoint: adds nothing, is not counted (below)
                if (dp[2] > vis && tp[2] > vis) {
                    const float scale_sq = pb_max((s.darea[d] + s.tarea[t]) * 0.5f, 1000.0f);
                    const float t2 = 2.0f * scale_sq;
                    const float dx = dp[0] - tp[0], dy = dp[1] - tp[1];
                    const float d2 = dx * dx + dy * dy;
                    const float sg = s.sig[k] * 2.0f;
                    const float s2 = sg * sg;
                    term = pb_expf(-d2 / (t2 * s2));
                }
                s.terms[idx] = term;
            }
            __syncthreads();
#pragma unroll 1
            for (int e = c.tid; e < ncur; e += c.nthreads) {
                const int key = s.cell_list[cb + e];
                const int t = key >> 16, d = key & 0xffff;
                // keypoint-ordered sum over the visible keypoints (:408-419).  An invisible keypoint left -0.0f: x + -0.0f
                // is x for every x this sum can hold (it starts at +0 and its terms are >= +0 or NaN), and a visible
                // term is never -0.0f, so neither the confidences nor a branch are needed here.
                float sum = 0.0f;
                int cnt = 0;
#pragma unroll
                for (int k = 0; k < KP; ++k) {
                    const float tv = s.terms[e * KP + k];
                    sum += tv;
                    cnt += (__float_as_uint(tv) != 0x80000000u) ? 1 : 0;
                }
                const float oks = (cnt >= 3) ? (sum / (float)cnt) : 0.0f;
                c.cost[(size_t)t * D + d] = 1.0f - oks;
            }
            __syncthreads();
        }
    }
    if (c.tid == 0) s.misc[5] = 0;               // (every thread has read the count: the last round ended with a barrier, or the count was zero)
}

// Torso-only OKS (kernelTorsoOKS :455-489): four exponentials per cell, one thread per cell.
static __device__ __forceinline__ void cost_pass_torso(Ctx& c, const unsigned* gate, int na) {
    TkSmem& s = c.s;
    const int D = c.D, Dw = c.Dw, words = (D + 31) >> 5;
    // one warp per (active row, 32 detections): after the tier-1 lock most gate words are empty
#pragma unroll 1
    for (int i = c.warp; i < na * words; i += c.nwarps) {
        const int ai = fast_div(i, c.magicW), w = i - ai * words;
        const int t = s.act_list[ai];
        const unsigned gw = gate[t * Dw + w];
        if (gw == 0u) continue;
        const int d = w * 32 + c.lane;
        if ((gw >> c.lane) & 1u)
            c.cost[(size_t)t * D + d] = torso_cost(c.pred + (size_t)t * POSE_F, c.det + (size_t)d * POSE_F);
    }
    __syncthreads();
}

static __device__ __forceinline__ void backup_assign(Ctx& c) {      // no barrier: callers synchronise before the next solve
    TkSmem& s = c.s;
#pragma unroll 1
    for (int t = c.tid; t < c.T; t += c.nthreads) s.rowb[t] = s.row[t];
#pragma unroll 1
    for (int d = c.tid; d < c.D; d += c.nthreads) s.colb[d] = s.col[d];
}
static __device__ __forceinline__ void merge_assign(Ctx& c) {   // kernelMergeAssignments :575-588
    TkSmem& s = c.s;
#pragma unroll 1
    for (int t = c.tid; t < c.T; t += c.nthreads) if (s.rowb[t] >= 0) s.row[t] = s.rowb[t];
#pragma unroll 1
    for (int d = c.tid; d < c.D; d += c.nthreads) if (s.colb[d] >= 0) s.col[d] = s.colb[d];
    __syncthreads();
}

__device__ __forceinline__ float center_iou(const float* a, const float* b) {   // kernelTrackIoU :825-854
    const float cx1 = a[0], cy1 = a[1], w1 = a[2], h1 = a[3];
    const float cx2 = b[0], cy2 = b[1], w2 = b[2], h2 = b[3];
    const float x1a = cx1 - w1 * 0.5f, x1b = cx1 + w1 * 0.5f, y1a = cy1 - h1 * 0.5f, y1b = cy1 + h1 * 0.5f;
    const float x2a = cx2 - w2 * 0.5f, x2b = cx2 + w2 * 0.5f, y2a = cy2 - h2 * 0.5f, y2b = cy2 + h2 * 0.5f;
    const float ix1 = pb_max(x1a, x2a), iy1 = pb_max(y1a, y2a);
    const float ix2 = pb_min(x1b, x2b), iy2 = pb_min(y1b, y2b);
    const float iw = pb_max(0.0f, ix2 - ix1), ih = pb_max(0.0f, iy2 - iy1);
    const float inter = iw * ih;
    const float a1 = w1 * h1, a2 = w2 * h2;
    const float uni = a1 + a2 - inter;
    return (uni > 0) ? (inter / uni) : 0.0f;
}

// ALLSMEM: cost matrix, detections and predicted poses all live in shared memory (the plan's usual case); the
// flags are then compile-time constants and every access to them is an LDS instead of a generic load.
// FUSED: called from the per-stream kernel of fused.cu right after the NMS sweep; this frame's detections (poses,
// scores) already sit in the tracker's shared-memory arrays and D_fused is their number (<= max_detections).
// In that kernel the CTA also OWNS the stream's chain of frames (chain word, fused.cu): the previous frame of this video
// stream is complete, so the two waits below are compiled out.  seq / frame_id / outputs / num_outputs: this frame's
// sequence number, frame id and record buffers (a fused CTA may go on with later frames of its stream).
// Chain word (below): bit 0 busy, bits 1..31 next sequence number, bits 32..46 mask of published frames, bits 47..61 mask of
// waiting frames (bit k of either: frame next + k).
__host__ __device__ __forceinline__ unsigned long long chain_make(unsigned waiting, unsigned mask, int next, int busy) {
    return ((unsigned long long)(waiting & 0x7fffu) << 47) | ((unsigned long long)(mask & 0x7fffu) << 32) |
           ((unsigned long long)((unsigned)next & 0x7fffffffu) << 1) | (unsigned long long)(busy & 1);
}
// nothing published, nobody owns the chain, frame `next` is next
__host__ __device__ __forceinline__ unsigned long long chain_pack(unsigned mask, int next, int busy) { return chain_make(0u, mask, next, busy); }

// ---- the chain word: who runs a video stream's next tracker stage -------------------------------------------------
// One 64-bit word per stream (chain_make), changed only by compare-and-swap:
//   next      sequence number of the stream's next tracker stage that has not been started;
//   busy      a CTA is running the stage next - 1 (it "owns the chain");
//   mask      bit k: the NMS stage of frame next + k has PUBLISHED its kept detections (ring slot (next + k) % depth) and its CTA is gone;
//   waiting   bit k: the CTA of frame next + k is alive and WAITS for its turn, its detections in shared memory.
// A CTA that finishes the NMS stage of frame s
//   * takes the chain (nobody owns it and s == next) and runs the tracker stage, or
//   * WAITS for its turn if it is close to the head of the chain (s - next < wait_window): it runs the detection-only part
//     of the tracker stage, then spins until the frames before its own have been run and the owner yields, or
//   * publishes its frame in the mask and EXITS: the owner will run the frame from the ring slot.
// At most wait_window CTAs per video stream wait at any time, each for at most wait_window tracker stages, and the host layer
// keeps wait_window * num_streams below the SM count, so the CTAs they wait for always find an SM.
// The owner, when it has written the stream's state back, yields if the next frame's CTA waits (that CTA then runs the
// state-dependent stages while this one still assembles its TrackOutput records), else goes on with the next frame itself
// if it has been published, else gives the chain up.  Frames of one video stream are thus processed strictly in order by
// whichever CTA holds the chain, streams never wait for each other, and a slow frame (an auction that runs to its iteration
// limit) delays only its own stream.
// A frame's tracker stage runs in the CTA of its own step's kernel or of an EARLIER step's: every tracker stage of steps <= i
// has run once the kernels of all steps <= i have completed — that is what the host layer's events rely on.
struct ChainW { unsigned mask, waiting; int next, busy; };
__device__ __forceinline__ ChainW chain_decode(unsigned long long v) {
    ChainW d;
    d.busy = (int)(v & 1ull); d.next = (int)((v >> 1) & 0x7fffffffull);
    d.mask = (unsigned)((v >> 32) & 0x7fffull); d.waiting = (unsigned)((v >> 47) & 0x7fffull);
    return d;
}
__device__ __forceinline__ unsigned long long chain_load(const unsigned long long* w) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(w) : "memory");
    return v;
}
// NMS stage of frame `seq` done, detections published.  0: exit; 1: this CTA owns the chain (frame seq is next); 2: wait for the turn
__device__ __forceinline__ int chain_arrive(unsigned long long* w, int seq, int wait_window) {
    __threadfence();                                     // release: detections, counts, frame id
    for (;;) {
        const unsigned long long cur = chain_load(w);
        const ChainW d = chain_decode(cur);
        const int k = seq - d.next;                      // >= 0: a tracker stage never starts before its NMS stage arrived
        unsigned long long want;
        int r;
        if (!d.busy && k == 0) { want = chain_make(d.waiting >> 1, d.mask >> 1, d.next + 1, 1); r = 1; }
        else if (k < wait_window) { want = chain_make(d.waiting | (1u << k), d.mask, d.next, d.busy); r = 2; }
        else { want = chain_make(d.waiting, d.mask | (1u << k), d.next, d.busy); r = 0; }
        if (atomicCAS(w, cur, want) == cur) { __threadfence(); return r; }    // acquire: the previous owner's state
    }
}
// waiting CTA: spin until the frames before `seq` have been run and the chain is free, then take it.  false: time-out.
__device__ __forceinline__ bool chain_take_over(unsigned long long* w, int seq) {
    const unsigned long long t0 = globaltimer_ns();
    for (;;) {
        const unsigned long long cur = chain_load(w);
        const ChainW d = chain_decode(cur);
        if (d.busy || d.next != seq) {                   // an owner is running, or an earlier frame has not arrived yet (its CTA will take the chain)
            if (globaltimer_ns() - t0 > 500000000ull) return false;
            __nanosleep(32);
            continue;
        }
        const unsigned long long want = chain_make(d.waiting >> 1, d.mask >> 1, d.next + 1, 1);
        if (atomicCAS(w, cur, want) == cur) { __threadfence(); return true; }
    }
}
// state written back: true = keep the chain and go on with the next frame (published, its CTA gone); false: the chain was
// handed to the waiting CTA of the next frame, or given up
__device__ __forceinline__ bool chain_advance(unsigned long long* w) {
    __threadfence();                                     // release: the stream's state
    for (;;) {
        const unsigned long long cur = chain_load(w);
        const ChainW d = chain_decode(cur);
        const bool go = !(d.waiting & 1u) && (d.mask & 1u);
        const unsigned long long want = go ? chain_make(d.waiting >> 1, d.mask >> 1, d.next + 1, 1) : chain_make(d.waiting, d.mask, d.next, 0);
        if (atomicCAS(w, cur, want) == cur) { __threadfence(); return go; }   // acquire: the published detections
    }
}

// ---------------------------------------------------------------------------------------
// Bulk asynchronous staging of the state slabs (cp.async.bulk, the 1-D form of the TMA: SASS UBLKCP) for the stand-alone
// tracker kernel: one thread queues the copies of a stream's slabs right after it has acquired the predecessor's release,
// the copy engine of the SM moves them into shared memory while the CTA goes on, and an mbarrier (expect-tx byte count)
// tells everybody when they have landed — instead of one LDG -> STS round trip per element on the issue slots of a
// latency-bound kernel, with the ordered active list then derived from shared memory instead of from L2.
// Sizes and addresses must be multiples of 16 bytes (T % 4 == 0; the slabs are cudaMalloc'ed and indexed by b * T).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");     // the initialised barrier is visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// wait_mode (stand-alone): 2 = the previous frame of this video stream was run by this very CTA (resident tracker,
// tracker.cu: pb_tracker_seq_kernel): both waits for the predecessor are skipped.
// wait_mode (fused): 1 = this CTA waits for its turn and takes the chain at the wait point; first_of_owner: the first
// frame this CTA runs (the previous owner may still be assembling its records: second wait).  Returns (fused) 1 when the
// CTA keeps the chain and the stream's next frame has arrived, 0 when the chain was handed on or given up, -1 on a time-out.
// RES (resident tracker, tracker.cu: pb_tracker_seq_kernel; stand-alone mode only): the CTA runs consecutive frames of its
// stream and the stream's state stays in shared memory between them — poses, velocities and dirty flags in arrays of their
// own, the small slabs, centres and the cost matrix where the stages keep them anyway.  res_first: load the state from
// global memory (first frame of the launch); res_last: write it back and release the stream (last frame).  In between a
// frame reads nothing but its detections from global memory and stores only what nobody on the device waits for
// (predictions, centres, scores, records).  Same expressions on the same values: same results.
template <int NTHREADS, bool ALLSMEM, bool FUSED, bool RES = false, bool BULK = false>
__device__ __forceinline__ int tracker_body(const TrackBuffers& tb, const TrackParams& P, const DetSource& src, const int b,
                                             unsigned char* smem_raw, const int D_fused, const int seq, const int frame_id,
                                             void* outputs, int* num_outputs, const int wait_mode = 0, const bool first_of_owner = true,
                                             const bool res_first = true, const bool res_last = true) {
    const bool st_load = !RES || res_first, st_store = !RES || res_last;
    // (no local copy of the parameter block: it stays in constant memory)
    const bool cost_in_smem = ALLSMEM || P.cost_in_smem, det_in_smem = ALLSMEM || P.det_in_smem, pred_in_smem = ALLSMEM || P.pred_in_smem;
    Ctx c;
    tk_from_offsets(smem_raw, P.so, c.s);
    c.term_floats = P.term_floats;
    TkSmem& s = c.s;
    const int T = P.T, Dm = P.Dm;
    c.cell_cap = P.cell_cap;
    c.T = T; c.tid = threadIdx.x; c.nthreads = NTHREADS; c.lane = threadIdx.x & 31;
    c.warp = threadIdx.x >> 5; c.nwarps = NTHREADS >> 5;
    const int tid = c.tid, NT = c.nthreads;

    // per-stream slabs (resident tracker: poses, velocities and dirty flags are the shared-memory copies)
    float* const gm_poses = tb.poses + (size_t)b * T * POSE_F;
    float* const gm_vel = tb.vel + (size_t)b * T * 34;
    int* const gm_dirty = tb.pred_dirty + (size_t)b * T;
    float* g_poses = RES ? s.poses : gm_poses;
    float* g_vel = RES ? s.vel : gm_vel;
    float* g_scores = tb.scores + (size_t)b * T;
    float* g_pred = tb.predicted + (size_t)b * T * POSE_F;
    float* g_tcent = tb.tcent + (size_t)b * T * 4;
    float* g_cost = tb.cost + (size_t)b * T * Dm;
    float* g_dscore = tb.det_scores + (size_t)b * Dm;
    int* g_states = tb.states + (size_t)b * T; int* g_ids = tb.ids + (size_t)b * T;
    int* g_hits = tb.hits + (size_t)b * T; int* g_ages = tb.ages + (size_t)b * T;
    int* g_last = tb.last_frame + (size_t)b * T; int* g_active = tb.active + (size_t)b * T;
    int* g_dirty = RES ? s.dirty : gm_dirty;
    int* g_scal = tb.scalars + (size_t)b * 4;
    unsigned long long* g_ns = tb.stage_ns + (size_t)b * 20;

    unsigned long long t_stamp = 0;
    if (tid == 0) t_stamp = globaltimer_ns();
    const unsigned long long t_begin = t_stamp;
    // stage telemetry (TrackerTiming): thread 0 accumulates globaltimer deltas in shared memory
    // and flushes them once at the end of the kernel
    auto stamp = [&](int slot) {
#ifndef PB_NO_STAMPS
        if (tid == 0) { const unsigned long long now = globaltimer_ns(); s.acc[slot] += now - t_stamp; t_stamp = now; }
#endif
    };

    // ---------------- prologue (:1065-1088) ----------------
    const int n_in = FUSED ? D_fused : src.num[b];
    const int D = n_in < Dm ? (n_in < 0 ? 0 : n_in) : Dm;
    c.D = D; c.Dw = (Dm + 31) / 32;
    c.magicD = div_magic(D); c.magicW = div_magic((D + 31) / 32);   // before the wait: off the chain of dependent frames
    const int Dw = c.Dw;
    float* det_w = det_in_smem ? s.det : (tb.det_poses_scratch + (size_t)b * Dm * POSE_F);
    const float* src_pose = FUSED ? det_w : src.poses + (size_t)b * src.stride * POSE_F;
    const float* src_score = FUSED ? s.dscore : src.scores + (size_t)b * src.stride;
    // Detections in shared memory: copied now.  In the global scratch (large max_detections) they are copied after the
    // wait below: the predecessor of this video stream — possibly still running on another lane — reads the same
    // scratch until its first release.
    if (!FUSED && det_in_smem) {
#pragma unroll 1
        for (int i = tid; i < D * POSE_F; i += NT) det_w[i] = src_pose[i];
    }
#pragma unroll 1
    for (int d = tid; d < D; d += NT) { if (!FUSED) s.dscore[d] = src_score[d]; s.col[d] = -1; }
    if (tid < 32) s.misc[tid] = 0;
    // state slabs by bulk asynchronous copy (stand-alone kernel, everything in shared memory, 16-byte multiples)
    __shared__ __align__(8) unsigned long long state_bar;
    // (P.bulk_off: 0 everything by bulk copy, 1 nothing, 2 centres + cost matrix only, 3 the six per-slot slabs only)
    const bool bulk_ok = BULK && ALLSMEM && !FUSED && !RES && (T & 3) == 0 && P.bulk_off != 1;
    const bool bulk = bulk_ok && P.bulk_off != 2;        // the six per-slot integer slabs
    const bool bulk_big = bulk_ok && P.bulk_off != 3;    // centres and cost matrix
    if (bulk_ok && tid == 0) mbar_init(&state_bar, 1);
    if (tid < 20 && st_load) s.acc[tid] = 0ull;     // (resident tracker: accumulated over the frames of the launch)
    if (tid < KP) s.sig[tid] = kSigmas[tid];
    if (D > 0) {
        // detection centres / areas (kernelComputeBboxCenters on the detections, :1180-1186) and the cleared gate
        // words depend on this frame's detections only: done here, before the wait for the predecessor, by the
        // upper half of the CTA (the lower half is still copying the records)
        const int h = NT >> 1;
#pragma unroll 1
        for (int d = tid - h; d >= 0 && d < D; d += h) {
            float area;
            pose_box(src_pose + (size_t)d * POSE_F, &s.dcent[d * 4], &area);
            s.darea[d] = area;
        }
#pragma unroll 1
        for (int i = tid; i < T * Dw; i += NT) { s.gate[i] = 0u; s.lgate[i] = 0u; }
    }
    // ---- per-stream ordering across launches ----
    // Consecutive tracker launches run on different CUDA streams (up to three "lanes", pb_api.cu) and may overlap:
    // the CTA of stream b only needs the state ITS predecessor (the previous frame of the same video stream) left
    // behind, not the whole previous grid, so a video stream whose auction ran to the iteration limit delays
    // nobody but itself.  Everything above reads this frame's detections only; from here on the CTA touches the
    // stream's persistent state and waits for seq_done[b] == seq - 1 (first release of the predecessor, acquire
    // here).  Launches are issued in sequence order and the host keeps at most `lanes` of them in flight; the
    // oldest never waits and the younger ones hold fewer SMs than the device has, so the oldest always runs to
    // completion; the time-out only guards against misuse.
    if (FUSED && wait_mode == 1 && tid == 0) {
        // waiting CTA: the frames before this one have been run and their owner yields at its state release
        if (!chain_take_over(tb.chain + b, seq)) { atomicExch(tb.error_flag, 1); s.misc[10] = 1; }
    }
    if (!FUSED && wait_mode != 2 && tid == 0) {
        const int want = seq - 1;
        const int* flag = tb.seq_done + b;
        const unsigned long long w0 = globaltimer_ns();
        bool timed_out = false;
        for (;;) {
            int v;
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if (v - want >= 0) break;
            if (globaltimer_ns() - w0 > 500000000ull) { atomicExch(tb.error_flag, 1); s.misc[10] = 1; timed_out = true; break; }
        }
        if (bulk_ok && !timed_out) {
            // the predecessor's slabs were written through the generic proxy and released; this thread has acquired them:
            // order that before the reads of the async proxy, then queue the copies (they complete on state_bar)
            asm volatile("fence.proxy.async.global;" ::: "memory");
            const unsigned tb4 = (unsigned)T * 4u, cost_b = (unsigned)T * (unsigned)D * 4u;
            mbar_expect_tx(&state_bar, (bulk ? 6u * tb4 : 0u) + (bulk_big ? 4u * tb4 + cost_b : 0u));
            if (bulk_big) {
                if (cost_b) bulk_g2s(s.cost, g_cost, cost_b, &state_bar);
                bulk_g2s(s.tcent, g_tcent, 4u * tb4, &state_bar);
            }
            if (bulk) {
                bulk_g2s(s.active, g_active, tb4, &state_bar);
                bulk_g2s(s.states, g_states, tb4, &state_bar);
                bulk_g2s(s.hits, g_hits, tb4, &state_bar);
                bulk_g2s(s.ids, g_ids, tb4, &state_bar);
                bulk_g2s(s.ages, g_ages, tb4, &state_bar);
                bulk_g2s(s.rowbc, gm_dirty, tb4, &state_bar);      // predicted pose changed since its centre was derived
            }
        }
        const unsigned long long w1 = globaltimer_ns();
        if (tb.dbg) { unsigned long long* q = tb.dbg + ((size_t)(seq & 63) * P.B + b) * 6; q[0] = t_begin; q[1] = w1; }
        s.acc[15] += w1 - w0;                        // telemetry: time spent waiting for the predecessor
        s.acc[16] += w0 - t_begin;                   // telemetry: detection-only prologue
        if (tb.dbg) {                                   // (PB_TIMELINE only: a dependent L2 round trip on the chain of frames)
            const unsigned long long prev_end = g_ns[18];   // absolute time at which the predecessor released the stream
            if (prev_end != 0ull && w1 > prev_end) s.acc[17] += w1 - prev_end;              // predecessor's release -> this CTA goes on
            if (prev_end != 0ull && t_begin > prev_end) s.acc[19] += t_begin - prev_end;    // ... of which: this CTA had not started yet
        }
    }
    __syncthreads();
    if (s.misc[10]) {
        // Time-out (misuse, or a foreign kernel starving the predecessor for 0.5 s): the stream's state is not this
        // frame's predecessor state.  Leave it untouched, pass the sequence number on so that later frames do not
        // wait again, and let the sticky error flag invalidate the results (every synchronising entry point reports it).
        if (tid == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(tb.seq_done + b), "r"(seq) : "memory");
            asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(tb.out_done + b), "r"(seq) : "memory");
            if (!FUSED) tb.chain[b] = chain_pack(0u, seq + 1, 0);
        }
        return -1;
    }
    if (!FUSED && !det_in_smem) {
#pragma unroll 1
        for (int i = tid; i < D * POSE_F; i += NT) det_w[i] = src_pose[i];
    }
#pragma unroll 1
    for (int d = tid; d < D; d += NT) g_dscore[d] = s.dscore[d];
    if (bulk) mbar_wait(&state_bar, 0u);            // (one use per CTA: phase 0; without the slabs the wait comes after the list, below)
    // State slabs -> shared memory, and in the same pass the ordered active list (ascending t).  The list position of
    // a row is its rank among the active rows — the auction breaks ties between equal bids by it (lowest row,
    // hungarian.cu:100) — so it must not depend on which warp gets here first: every warp derives the number of
    // active rows in front of its block of 32 from ballots over the preceding blocks (read from the state slab
    // itself, so no barrier is needed between the copy and the list) instead of from an atomic counter.
#pragma unroll 1
    for (int t0 = c.warp * 32; t0 < T; t0 += NT) {
        const int t = t0 + c.lane;
        int a = 0, st = 0;
        if (t < T) {
            if (bulk) {                     // the slabs have landed in shared memory (mbarrier above)
                a = s.active[t]; st = s.states[t];
            } else if (st_load) {
                a = g_active[t]; st = g_states[t];
                s.active[t] = a; s.states[t] = st; s.hits[t] = g_hits[t]; s.ids[t] = g_ids[t]; s.ages[t] = g_ages[t];
                const int dy = gm_dirty[t];
                if (RES) s.dirty[t] = dy;
                s.rowbc[t] = dy;            // predicted pose changed since its centre was derived (idle auction scratch)
            } else {                        // resident tracker: the previous frame left the slabs in shared memory
                a = s.active[t]; st = s.states[t];
                s.rowbc[t] = s.dirty[t];
            }
            s.row[t] = -1;
        }
        const int* act_src = (st_load && !bulk) ? g_active : s.active;
        int start = 0;
        if (T <= 256) {
            // all preceding blocks' flags in flight at once (from L2 when the slabs are not in shared memory: one round trip
            // instead of one per block); blocks at or beyond this warp's own contribute nothing
            int fl[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) fl[k] = (k * 32 < t0) ? act_src[k * 32 + c.lane] : 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) start += __popc(__ballot_sync(FULLM, fl[k] == 1));
        } else {
#pragma unroll 1
            for (int pb = 0; pb < t0; pb += 32) start += __popc(__ballot_sync(FULLM, act_src[pb + c.lane] == 1));
        }
        const bool act = (a == 1);
        const unsigned bm = __ballot_sync(FULLM, act);
        const unsigned lm = __ballot_sync(FULLM, act && st == ST_LOST);
        if (c.lane == 0 && lm) atomicAdd(&s.misc[6], __popc(lm));              // LOST rows at frame start (a sum: order-free)
        if (act) s.act_list[start + __popc(bm & ((1u << c.lane) - 1u))] = t;
        if (c.lane == 0 && t0 + 32 >= T) s.misc[0] = start + __popc(bm);
    }
    if (st_load && !bulk_big) {
#pragma unroll 1
        for (int i = tid; i < T * 4; i += NT) s.tcent[i] = g_tcent[i];
        if (RES) {
#pragma unroll 1
            for (int i = tid; i < T * POSE_F; i += NT) s.poses[i] = gm_poses[i];
#pragma unroll 1
            for (int i = tid; i < T * 34; i += NT) s.vel[i] = gm_vel[i];
        }
    }
    c.det = det_w;
    c.cost = cost_in_smem ? s.cost : g_cost;
    c.warp_auction = cost_in_smem && T <= 1024 && (long)T * Dm <= 16384;
    c.sub_solve = P.term_floats >= 2048 && P.cell_cap >= 192 && P.sub_solve_off == 0;
    c.od_flag = (!FUSED && wait_mode != 2) ? (tb.out_done + b) : nullptr;
    c.od_want = seq - 1;
    c.pred = pred_in_smem ? s.pred : g_pred;
    // (resident tracker: the whole persistent matrix is loaded once — later frames index it with their own D, quirk Q1)
    if (cost_in_smem && st_load && !bulk_big) for (int i = tid; i < (RES ? T * Dm : T * D); i += NT) s.cost[i] = g_cost[i];
    if (bulk_big && !bulk) mbar_wait(&state_bar, 0u);
    __syncthreads();
    const int na = s.misc[0];       // num_active_tracks_ at frame start (:1083-1088)
    stamp(0);

    // ---------------- predict (:1160-1175, kernel :102-138) ----------------
    const bool pre = P.precomputed != 0;            // large tables: done by the row-sliced pre-kernel (tracker.cu) just before this launch
    if (pre) {
        if (pred_in_smem) {
#pragma unroll 1
            for (int i = tid; i < na * POSE_F; i += NT) { const int ai = i / POSE_F, e = i - ai * POSE_F; const int t = s.act_list[ai]; s.pred[t * POSE_F + e] = g_pred[t * POSE_F + e]; }
        }
    } else if (na > 0) {
#pragma unroll 1
        for (int i = tid; i < na * KP; i += NT) {
            const int ai = i / KP, k = i - ai * KP;
            const int t = s.act_list[ai];
            const int po = t * POSE_F + k * 3, vo = t * 34 + k * 2;
            const float x = g_poses[po], y = g_poses[po + 1], cf = g_poses[po + 2];
            const float vx = g_vel[vo], vy = g_vel[vo + 1];
            const float dt = 1.0f;
            const float px = x + vx * dt, py = y + vy * dt;
            g_pred[po] = px; g_pred[po + 1] = py; g_pred[po + 2] = cf;
            if (pred_in_smem) { s.pred[po] = px; s.pred[po + 1] = py; s.pred[po + 2] = cf; }
            if (s.states[t] == ST_LOST) { g_vel[vo] = vx * 0.95f; g_vel[vo + 1] = vy * 0.95f; }
            if (k == 0) g_dirty[t] = 1;
        }
    }
    __syncthreads();
    stamp(1);

    const bool assoc12 = (na > 0) && (D > 0);
    // ---------------- centres + gates (:1177-1208, kernels :196-317) ----------------
    if (assoc12 && pre) {
        // centres came with the state slab (prologue); areas and gates from the pre-kernel's buffers
        const float* ga = tb.tarea_g + (size_t)b * T;
        const unsigned* gg = tb.gate_g + (size_t)b * T * Dw;
        const unsigned* gl = tb.lgate_g + (size_t)b * T * Dw;
#pragma unroll 1
        for (int t = tid; t < T; t += NT) s.tarea[t] = ga[t];
#pragma unroll 1
        for (int i = tid; i < T * Dw; i += NT) { s.gate[i] = gg[i]; s.lgate[i] = gl[i]; }
    } else if (assoc12) {
        // track centres: every slot whose predicted pose changed since its centre was derived
        // (== the reference recomputing all T slots: unchanged slots give unchanged centres)
#pragma unroll 1
        for (int t = tid; t < T; t += NT) {
            if (s.rowbc[t] || (s.active[t] == 1)) {     // predict marks every active slot dirty (below)
                const float* pp = (pred_in_smem && s.active[t] == 1) ? (s.pred + (size_t)t * POSE_F) : (g_pred + (size_t)t * POSE_F);
                float area;
                pose_box(pp, &s.tcent[t * 4], &area);
                s.tarea[t] = area;
                g_tcent[t * 4] = s.tcent[t * 4]; g_tcent[t * 4 + 1] = s.tcent[t * 4 + 1];
                g_tcent[t * 4 + 2] = s.tcent[t * 4 + 2]; g_tcent[t * 4 + 3] = s.tcent[t * 4 + 3];
                g_dirty[t] = 0;
            }
        }
#pragma unroll 1
        for (int ai = tid - (NT >> 1); ai >= 0 && ai < na; ai += (NT >> 1)) {   // mean torso speed per active row (:287-298), upper half of the CTA
            const int t = s.act_list[ai];
            const int torso[4] = {5, 6, 11, 12};
            float av = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float vx = g_vel[t * 34 + torso[i] * 2], vy = g_vel[t * 34 + torso[i] * 2 + 1];
                av += sqrtf(vx * vx + vy * vy);
            }
            s.tav[t] = av * 0.25f;
        }
    }
    __syncthreads();
    if (assoc12 && !pre) {
        // gate (base 3.0) minus LOST rows (tier 1 mask, :1231) and lost-tier gate (base 3.0*1.3,
        // only LOST rows survive the two state masks, :1359-1387).  One warp per (row, 32 dets).
        const int words = (D + 31) / 32;
#pragma unroll 1
        for (int i = c.warp; i < na * words; i += c.nwarps) {
            const int ai = fast_div(i, c.magicW), w = i - ai * words;
            const int t = s.act_list[ai];
            const int d = w * 32 + c.lane;
            const bool lost = (s.states[t] == ST_LOST);
            const float base = lost ? (3.0f * 1.3f) : 3.0f;
            bool g = false;
            if (d < D) g = gate_cell(&s.tcent[t * 4], &s.dcent[d * 4], s.tav[t], lost, base, P.gating_enabled != 0);
            const unsigned bm = __ballot_sync(FULLM, g);
            if (c.lane == 0) { if (lost) s.lgate[t * Dw + w] = bm; else s.gate[t * Dw + w] = bm; }
        }
    }
    __syncthreads();
    stamp(2);

    // ---------------- tiers 1-3 (:1210-1274, :1276-1335, :1337-1436) ----------------
    // One rolled loop: a single copy of the cost passes, the auction and the lock in the instruction
    // stream (the kernel runs each of them once or a few times per launch, from a cold instruction cache).
#pragma unroll 1
    for (int tier = 0; tier < 3; ++tier) {
        const bool run = (tier < 2) ? assoc12 : (D > 0);
        if (run) {
            if (tier > 0) backup_assign(c);
            if (tier == 0) {
                if (!pre) cost_pass_oks(c, s.gate, na, 0.2f);
                stamp(12);
            } else if (tier == 1) {
                cost_pass_torso(c, s.gate, na);
            } else {
                cost_inactive_rows(c);
                lock_pairs(c, s.lgate, na, true);
                // the lost-tier gate holds bits of LOST rows only (gate stage): without such a row it is empty
                if (s.misc[6] > 0) cost_pass_oks(c, s.lgate, na, 0.2f);
            }
            auction_solve(c, na, tier > 0);
            if (tier == 0) stamp(13);
            if (tier > 0) merge_assign(c);
            if (tier < 2) lock_pairs(c, s.gate, na, tier > 0);
            if (tier == 0) stamp(14);
            if (tier == 0 && c.warp_auction && na <= 32) {
                // Every active row matched in tier 1 (the usual frame: everybody found its detection), or every detection
                // taken (a stream with a track more than it has detections: the frames whose auction ran longest)?  Tiers 2
                // and 3 then change nothing but the cells of the inactive rows (cost <- 1.0, :351-354).  lock_pairs has
                // cleared the gate words of the matched rows and the gate bits of the matched columns and set the cells of
                // both to 1e9 in every active row, so neither cost pass writes a cell, no row finds a column above the -1e9
                // floor, both solves clear the assignments and stop before their first iteration, the merges restore them
                // and the locks rewrite the 1e9 cells of tier 1.  Do just that write and leave the loop: six block
                // barriers (2.5 + 3 us where a row is left over) less on the chain of dependent frames.
                const bool row_left = c.lane < na && s.row[s.act_list[c.lane]] < 0;
                bool col_left = false;
                for (int d = c.lane; d < c.D; d += 32) col_left |= s.col[d] < 0;
                // (the same in every warp: row and col are final, lock_pairs ended with a barrier)
                if (__ballot_sync(FULLM, row_left) == 0u || __ballot_sync(FULLM, col_left) == 0u) {
                    cost_inactive_rows(c);                             // (ordered by the barrier behind the loop)
                    stamp(3);
                    break;
                }
            }
        }
        stamp(3 + tier);
    }

    // The predecessor's record assembly reads g_poses after it released the state (second release at its end):
    // nothing before this point writes g_poses, tb.outputs or the telemetry slots; wait for it here (it finished
    // long ago unless the launches ran far apart from the usual order).
    if ((FUSED ? first_of_owner : wait_mode != 2) && tid == 0 && s.misc[15] == 0) {      // (misc[15]: seen during the tier-1 solve)
        const int want = seq - 1;
        const unsigned long long w0 = globaltimer_ns();
        for (;;) {
            int v;
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(tb.out_done + b) : "memory");
            if (v - want >= 0) break;
            if (globaltimer_ns() - w0 > 500000000ull) { atomicExch(tb.error_flag, 1); s.misc[10] = 1; break; }
        }
    }
    __syncthreads();
    if (s.misc[10]) {       // see the first wait: no update, no records; only the prediction scratch has been touched
        if (tid == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(tb.seq_done + b), "r"(seq) : "memory");
            asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(tb.out_done + b), "r"(seq) : "memory");
            if (!FUSED) tb.chain[b] = chain_pack(0u, seq + 1, 0);
        }
        return -1;
    }
    // ---------------- update matched (:1438-1472, kernels :141-189, :612-648) -------------
    if (D > 0) {
        const float process_noise = 0.1f, measurement_noise = 0.3f;
        const float K = measurement_noise / (measurement_noise + process_noise);
        const float alpha = 0.3f;
#pragma unroll 1
        for (int i = tid; i < na * KP; i += NT) {
            const int ai = i / KP, k = i - ai * KP;
            const int t = s.act_list[ai];
            const int d = s.row[t];
            if (d < 0) continue;
            const int to = t * POSE_F + k * 3, vo = t * 34 + k * 2;
            const float* dp = c.det + (size_t)d * POSE_F + k * 3;
            const float ox = g_poses[to], oy = g_poses[to + 1];
            const float zx = dp[0], zy = dp[1], zc = dp[2];
            const float nx = ox + K * (zx - ox);
            const float ny = oy + K * (zy - oy);
            const float dx = zx - ox, dy = zy - oy;
            g_vel[vo] = alpha * dx + (1 - alpha) * g_vel[vo];
            g_vel[vo + 1] = alpha * dy + (1 - alpha) * g_vel[vo + 1];
            g_poses[to] = nx; g_poses[to + 1] = ny; g_poses[to + 2] = zc;
        }
#pragma unroll 1
        for (int ai = tid; ai < na; ai += NT) {
            const int t = s.act_list[ai];
            const int d = s.row[t];
            if (d < 0) continue;
            g_scores[t] = s.dscore[d];
            const int h = s.hits[t] + 1;
            s.hits[t] = h;
            s.ages[t] = 0;
            g_last[t] = frame_id;
            const int st = s.states[t];
            if (st == ST_TENTATIVE && h >= P.min_hits) s.states[t] = ST_CONFIRMED;
            else if (st == ST_LOST) s.states[t] = ST_CONFIRMED;
        }
    }
    stamp(6);       // (no barrier: the ageing pass below touches the unmatched rows only, the loop above the matched ones)

    // ---------------- age unmatched (:1474-1487, kernel :651-688) ----------------
#pragma unroll 1
    for (int ai = tid; ai < na; ai += NT) {
        const int t = s.act_list[ai];
        if (s.row[t] >= 0) continue;
        const int age = s.ages[t] + 1;
        s.ages[t] = age;
        const int st = s.states[t];
        if (st == ST_TENTATIVE) { if (age > 2) s.active[t] = 0; }
        else if (st == ST_CONFIRMED) { if (age > P.max_age) s.states[t] = ST_LOST; }
        else if (st == ST_LOST) { if (age > P.max_age + 10) s.active[t] = 0; }
    }
    __syncthreads();
    stamp(7);

    // ---------------- new tracks (:1489-1526; R3/R4: ascending detection order) ------------
    if (D > 0) {
        // which detections start a track?  (unmatched and score >= new_track_thresh.)  Usually none: warp 0 finds that out
        // with ballots and the serial pass below, the pose copies and a barrier are skipped.
        if (tid < 32) {
            unsigned any = 0u;
#pragma unroll 1
            for (int base = 0; base < D; base += 32) {
                const int d = base + c.lane;
                bool q = false;
                if (d < D) { s.slot_for_det[d] = -1; q = s.col[d] < 0 && !(s.dscore[d] < P.new_track_thresh); }
                any |= __ballot_sync(FULLM, q);
            }
            if (c.lane == 0) s.misc[7] = any != 0u;
        }
        __syncthreads();
        const bool newborn = s.misc[7] != 0;
        if (newborn) {
            // R3/R4 in parallel.  The reference's rule, made sequential (ascending detection index; the k-th qualifying
            // detection starts at slot (hint + k) mod T and takes the first free slot from there, circularly), has a
            // closed form: starts advance by one per detection and free slots are distinct, so the k-th free slot in
            // circular order from s0 = hint mod T lies at least k slots beyond s0 and is never overtaken — the k-th
            // qualifying detection gets exactly the k-th free slot from s0, until the free slots run out (later
            // detections get none; the hint still advances for each of them).  Ids are issued in the same order.
            const int hint0 = g_scal[1], id0 = g_scal[0];
            const int s0 = hint0 % T;
            int* free_by_rank = s.elig_list;         // [T], idle until the de-duplication stage
            int* det_rank = s.out_list;              // [Dm], idle until the output stage
            if (tid < 32) {                           // free slots in all, and in [s0, T)
                int all = 0, up = 0;
#pragma unroll 1
                for (int pb = 0; pb < T; pb += 32) {
                    const int u = pb + c.lane;
                    const bool f = (u < T) && s.active[u] == 0;
                    all += __popc(__ballot_sync(FULLM, f));
                    up += __popc(__ballot_sync(FULLM, f && u >= s0));
                }
                if (c.lane == 0) { s.misc[13] = all; s.misc[14] = up; }
            }
            __syncthreads();
            const int nfree = s.misc[13], upper = s.misc[14];
            // rank of every free slot in circular order from s0: slots >= s0 first (ascending), then the slots below s0
#pragma unroll 1
            for (int base = c.warp * 32; base < T; base += c.nwarps * 32) {
                int lo = 0, ge = 0;
#pragma unroll 1
                for (int pb = 0; pb < base; pb += 32) {
                    const int u = pb + c.lane;
                    const bool f = s.active[u] == 0;
                    lo += __popc(__ballot_sync(FULLM, f && u < s0));
                    ge += __popc(__ballot_sync(FULLM, f && u >= s0));
                }
                const int t = base + c.lane;
                const bool fr = (t < T) && s.active[t] == 0;
                const unsigned mlo = __ballot_sync(FULLM, fr && t < s0), mge = __ballot_sync(FULLM, fr && t >= s0);
                const unsigned below = (1u << c.lane) - 1u;
                if (fr) free_by_rank[(t >= s0) ? (ge + __popc(mge & below)) : (upper + lo + __popc(mlo & below))] = t;
            }
            // rank of every qualifying detection (ascending detection index)
#pragma unroll 1
            for (int base = c.warp * 32; base < D; base += c.nwarps * 32) {
                int start = 0;
#pragma unroll 1
                for (int pb = 0; pb < base; pb += 32) {
                    const int u = pb + c.lane;
                    start += __popc(__ballot_sync(FULLM, s.col[u] < 0 && !(s.dscore[u] < P.new_track_thresh)));
                }
                const int d = base + c.lane;
                const bool q = (d < D) && s.col[d] < 0 && !(s.dscore[d] < P.new_track_thresh);
                const unsigned bm = __ballot_sync(FULLM, q);
                if (d < D) det_rank[d] = q ? (start + __popc(bm & ((1u << c.lane) - 1u))) : -1;
                if (base + 32 >= D && c.lane == 0) s.misc[12] = start + __popc(bm);            // qualifying detections in all
            }
            __syncthreads();
#pragma unroll 1
            for (int d = tid; d < D; d += NT) {
                const int k = det_rank[d];
                if (k < 0 || k >= nfree) continue;                   // not qualifying, or no free slot left for it
                const int sl = free_by_rank[k];
                s.slot_for_det[d] = sl;
                s.active[sl] = 1;
                s.ids[sl] = id0 + k;
                s.hits[sl] = 1; s.ages[sl] = 0; s.states[sl] = ST_TENTATIVE;
                g_scores[sl] = s.dscore[d];
                g_last[sl] = frame_id;
                s.col[d] = sl;
            }
            if (tid == 0) {
                const int nq = s.misc[12];
                g_scal[1] = hint0 + nq;
                g_scal[0] = id0 + (nq < nfree ? nq : nfree);
            }
        }
        if (newborn) {
        __syncthreads();
#pragma unroll 1
        for (int i = tid; i < D * POSE_F; i += NT) {
            const int d = i / POSE_F, e = i - d * POSE_F;
            const int sl = s.slot_for_det[d];
            if (sl >= 0) g_poses[sl * POSE_F + e] = c.det[(size_t)d * POSE_F + e];
        }
#pragma unroll 1
        for (int i = tid; i < D * 34; i += NT) {
            const int d = i / 34, e = i - d * 34;
            const int sl = s.slot_for_det[d];
            if (sl >= 0) g_vel[sl * 34 + e] = 0.0f;
        }
        }
    }
    __syncthreads();
    stamp(8);

    // ---------------- de-duplication (:1528-1557; R5: sequential semantics) ----------------
    {
        // misc[1] (eligible count) and misc[2] (duplicate pairs) are still zero from the prologue: nothing else uses them
        for (int base = c.warp * 32; base < T; base += c.nwarps * 32) {   // eligible list, ascending (ordered like act_list above)
            auto elig = [&](int t) { return (t < T) && s.active[t] == 1 && s.states[t] != ST_LOST && s.hits[t] >= P.min_hits; };
            int start = 0;
            for (int pb = 0; pb < base; pb += 32) start += __popc(__ballot_sync(FULLM, elig(pb + c.lane)));
            const int t = base + c.lane;
            const bool e = elig(t);
            const unsigned bm = __ballot_sync(FULLM, e);
            if (e) s.elig_list[start + __popc(bm & ((1u << c.lane) - 1u))] = t;
            if (c.lane == 0 && base + 32 >= T) s.misc[1] = start + __popc(bm);
        }
        __syncthreads();
        const int ne = s.misc[1];
        if (ne <= 96) {
            // short lists (the usual frame): one thread per ordered pair
            const unsigned magic_ne = div_magic(ne);
#pragma unroll 1
            for (int i = tid; i < ne * ne; i += NT) {
                const int ia = fast_div(i, magic_ne), ib = i - ia * ne;
                const int t1 = s.elig_list[ia], t2 = s.elig_list[ib];
                if (t1 < t2 && center_iou(&s.tcent[t1 * 4], &s.tcent[t2 * 4]) > 0.7f) {
                    const int pos = atomicAdd(&s.misc[2], 1);
                    if (pos < DUP_CAP) s.dup[pos] = (t1 << 16) | t2;
                }
            }
        } else {
        // pairs (a, b > a) of the eligible list: one thread per (a, chunk of 64 b's)
            const int chunks = (ne + 63) >> 6;
            const unsigned magic_ch = div_magic(chunks);
#pragma unroll 1
            for (int i = tid; i < ne * chunks; i += NT) {
                const int ia = fast_div(i, magic_ch), ch = i - ia * chunks;
                int ib = ch << 6;
                const int ie = (ib + 64 < ne) ? ib + 64 : ne;
                if (ib <= ia) ib = ia + 1;
                if (ib >= ie) continue;
                const int t1 = s.elig_list[ia];
                const float ax = s.tcent[t1 * 4], ay = s.tcent[t1 * 4 + 1], aw = s.tcent[t1 * 4 + 2], ah = s.tcent[t1 * 4 + 3];
#pragma unroll 1
                for (; ib < ie; ++ib) {
                    const int t2 = s.elig_list[ib];                       // ascending list: t1 < t2
                    const float* b4 = &s.tcent[t2 * 4];
                    // boxes whose centres are further apart than the sum of their half extents (plus a margin far above any
                    // rounding of the corner arithmetic) do not intersect: IoU 0, never above 0.7.  NaNs take the complete test.
                    const float hw = (fabsf(aw) + fabsf(b4[2])) * 0.5f, hh = (fabsf(ah) + fabsf(b4[3])) * 0.5f;
                    if (fabsf(ax - b4[0]) > hw * 1.001f + 1.0f || fabsf(ay - b4[1]) > hh * 1.001f + 1.0f) continue;
                    if (center_iou(&s.tcent[t1 * 4], b4) > 0.7f) {
                        const int pos = atomicAdd(&s.misc[2], 1);
                        if (pos < DUP_CAP) s.dup[pos] = (t1 << 16) | t2;
                    }
                }
            }
        }
        __syncthreads();
        const int ndup = s.misc[2];
        if (tid == 0 && ndup > 0) {
            auto resolve = [&](int t1, int t2) {
                if (s.active[t1] == 0 || s.active[t2] == 0) return;      // LOST excluded by eligibility
                if (s.hits[t1] < s.hits[t2] || (s.hits[t1] == s.hits[t2] && s.ids[t1] > s.ids[t2])) s.active[t1] = 0;
                else s.active[t2] = 0;
            };
            if (ndup <= DUP_CAP) {
                for (int i = 1; i < ndup; ++i) {                         // lexicographic (t1, t2)
                    const int key = s.dup[i];
                    int j = i - 1;
                    while (j >= 0 && s.dup[j] > key) { s.dup[j + 1] = s.dup[j]; --j; }
                    s.dup[j + 1] = key;
                }
                for (int i = 0; i < ndup; ++i) resolve(s.dup[i] >> 16, s.dup[i] & 0xffff);
            } else {                                                     // overflow: recompute in order
                for (int ia = 0; ia < ne; ++ia)
                    for (int ib = 0; ib < ne; ++ib) {
                        const int t1 = s.elig_list[ia], t2 = s.elig_list[ib];
                        if (t1 < t2 && center_iou(&s.tcent[t1 * 4], &s.tcent[t2 * 4]) > 0.7f) resolve(t1, t2);
                    }
            }
        }
        __syncthreads();
    }
    stamp(9);

    // ---------------- write back state ----------------
    // (resident tracker: after the last frame of the launch only — nobody reads the stream's state in between)
    // (the column assignments are stored every frame: a frame writes its own D entries and the tail keeps older frames' values)
#pragma unroll 1
    for (int d = tid; d < D; d += NT) tb.col_assign[(size_t)b * Dm + d] = s.col[d];
    if (st_store) {
    int cnt_local = 0;
#pragma unroll 1
    for (int t = tid; t < T; t += NT) {
        g_active[t] = s.active[t]; g_states[t] = s.states[t]; g_hits[t] = s.hits[t]; g_ids[t] = s.ids[t];
        g_ages[t] = s.ages[t];
        if (RES) gm_dirty[t] = s.dirty[t];
        tb.row_assign[(size_t)b * T + t] = s.row[t];
        cnt_local += (s.active[t] == 1);
    }
    if (cost_in_smem) for (int i = tid; i < (RES ? T * Dm : T * D); i += NT) g_cost[i] = s.cost[i];
    if (RES) {
#pragma unroll 1
        for (int i = tid; i < T * POSE_F; i += NT) gm_poses[i] = s.poses[i];
#pragma unroll 1
        for (int i = tid; i < T * 34; i += NT) gm_vel[i] = s.vel[i];
    }
    // block-wide sum of cnt_local (update()'s return value, :1130-1136)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cnt_local += __shfl_xor_sync(FULLM, cnt_local, off);
    if (c.lane == 0 && cnt_local) atomicAdd(&s.misc[4], cnt_local);
    __syncthreads();
    if (tid == 0) {
        // ---- first release: the stream's STATE is final; its next frame may go on (see the wait in the prologue).
        // The TrackOutput records below read g_poses and write tb.outputs; the successor touches neither before its
        // update stage, where it waits for the second flag (out_done).  The ~2 us of record assembly thus leave
        // the chain of dependent frames.
        g_scal[2] = D; g_scal[3] = s.misc[4];
        g_ns[18] = globaltimer_ns();
        if (tb.dbg) tb.dbg[((size_t)(seq & 63) * P.B + b) * 6 + 2] = g_ns[18];
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(tb.seq_done + b), "r"(seq) : "memory");
        // fused path: hand the chain on here — a waiting successor starts its state-dependent stages while this CTA assembles
        // the records; or keep it (the next frame has arrived: this CTA goes on with it afterwards); or give it up
        if (FUSED) s.misc[11] = chain_advance(tb.chain + b) ? 1 : 0;
    }
    }   // st_store

    // ---------------- outputs: getActiveTracks (:1594-1636) on the device ------------------
    if (tid < 32) {
        int count = 0;
        for (int base = 0; base < D; base += 32) {
            const int d = base + c.lane;
            bool emit = false;
            if (d < D) {
                const int sl = s.col[d];
                if (sl >= 0) {
                    const int st = s.states[sl];
                    emit = !(st == ST_TENTATIVE && s.hits[sl] < P.min_hits) && (st != ST_LOST);
                }
            }
            const unsigned bm = __ballot_sync(FULLM, emit);
            if (emit) s.out_list[count + __popc(bm & ((1u << c.lane) - 1u))] = d;
            count += __popc(bm);
        }
        if (c.lane == 0) s.misc[3] = count;
    }
    __syncthreads();
    const int n_out = s.misc[3];
    {
        // optional un-letterboxing of the records, scaleTrackOutputs (reference src/main.cpp:48-68):
        // value = (value - pad) * scale, applied to the box and the keypoints after the box was built
        const bool xf = tb.out_xform != nullptr;
        const float sx = xf ? tb.out_xform[b * 4 + 0] : 1.0f, sy = xf ? tb.out_xform[b * 4 + 1] : 1.0f;
        const float px_ = xf ? tb.out_xform[b * 4 + 2] : 0.0f, py_ = xf ? tb.out_xform[b * 4 + 3] : 0.0f;
        float* outw = reinterpret_cast<float*>(outputs) + (size_t)b * Dm * 57;
#pragma unroll 1
        for (int o = c.warp; o < n_out; o += c.nwarps) {
            const int d = s.out_list[o];
            const int sl = s.col[d];
            float* rec = outw + (size_t)o * 57;
            float x = 0.f, y = 0.f, cf = 0.f;
            float lx = 1e9f, ly = 1e9f, hx = -1e9f, hy = -1e9f;
            if (c.lane < KP) {
                x = g_poses[sl * POSE_F + c.lane * 3]; y = g_poses[sl * POSE_F + c.lane * 3 + 1]; cf = g_poses[sl * POSE_F + c.lane * 3 + 2];
                rec[6 + c.lane * 3] = xf ? (x - px_) * sx : x; rec[7 + c.lane * 3] = xf ? (y - py_) * sy : y; rec[8 + c.lane * 3] = cf;
                if (cf > 0.2f) { lx = x; ly = y; hx = x; hy = y; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                lx = pb_min(lx, __shfl_xor_sync(FULLM, lx, off)); ly = pb_min(ly, __shfl_xor_sync(FULLM, ly, off));
                hx = pb_max(hx, __shfl_xor_sync(FULLM, hx, off)); hy = pb_max(hy, __shfl_xor_sync(FULLM, hy, off));
            }
            if (c.lane == 0) {
                const float px = (hx - lx) * 0.1f, py = (hy - ly) * 0.1f;
                reinterpret_cast<int*>(rec)[0] = s.ids[sl];
                rec[1] = s.dscore[d];
                const float b0 = lx - px, b1 = ly - py, b2 = hx + px, b3 = hy + py;
                rec[2] = xf ? (b0 - px_) * sx : b0; rec[3] = xf ? (b1 - py_) * sy : b1;
                rec[4] = xf ? (b2 - px_) * sx : b2; rec[5] = xf ? (b3 - py_) * sy : b3;
            }
        }
    }

    if (tid == 0) {
        num_outputs[b] = n_out;
        const unsigned long long now = globaltimer_ns();
        s.acc[10] += now - t_begin;
        s.acc[11] += 1ull;
    }
    __syncthreads();
    if (!st_store) return 0;            // resident tracker: the next frame follows in this CTA
    if (tid < 18 && s.acc[tid] != 0ull) g_ns[tid] += s.acc[tid];
    if (tid == 19 && s.acc[19] != 0ull) g_ns[19] += s.acc[19];
    // second release: records and telemetry written, g_poses no longer read
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(tb.out_done + b), "r"(seq) : "memory");
        // stand-alone launches run with nothing of the fused path in flight (pb_api.cu joins first): a plain store keeps the
        // chain word of the fused path current (next frame = seq + 1, nothing published, nobody owns the chain)
        if (!FUSED) tb.chain[b] = chain_pack(0u, seq + 1, 0);
    }
    return FUSED ? s.misc[11] : 0;      // written before the last barrier above
}


}  // namespace pb
