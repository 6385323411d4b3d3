// nms_body.cuh — rank + OKS/IoU pose-NMS of one stream, executed by one CTA (device code shared by the stand-alone NMS
// kernel, decode_nms.cu, and the fused per-stream kernel, fused.cu).  See decode_nms.cu for the description of the sweep.
#pragma once
#include "pb_common.cuh"

namespace pb {

constexpr unsigned FULL = 0xffffffffu;

// =======================================================================================
// K2: rank + NMS
// =======================================================================================
constexpr int NM_LIST = 2048;                // undecided pairs per round (>= 64*63/2 = 2016)
constexpr int NM_TERM_PAIRS = 128;           // pairs whose 17 terms are evaluated at once

struct NmSmem {
    unsigned long long* tmask;   // [64] overlap rows of the current tile
    unsigned long long* acc;     // [16] stage telemetry
    float* score;      // [Ccap] by slot
    int* recidx;       // [Ccap] slot -> record index inside the stream's scratch
    int* order;        // [Ccap] rank -> slot
    float* kx;         // [17][CS] by rank (CS = Ccap + 1: bank-conflict-free transposing stores)
    float* ky;         // [17][CS]
    unsigned* vis;     // [Ccap] bit k: conf_k > 0.2
    float* box;        // [4][CS] cx,cy,w,h then x1,y1,x2,y2
    float* area;       // [Ccap]
    float* ext;        // [4][CS] keypoint extents lx, hx, ly, hy over all 17 keypoints
    unsigned* sup;     // [Ccap/32 + 2]
    int* keep;         // [Kcap] kept ranks
    int* tk;           // [64] ranks kept in the current tile
    int* misc;         // [16]: 0 C, 1 nkeep, 2 ntk, 3 list1 length, 4 list2 length
    unsigned* l1_key;  // [NM_LIST] pairs that passed stage A undecided (rank_i << 16 | rank_j)
    unsigned* l2_key;  // [NM_LIST] pairs that passed stage B undecided
    float* terms;      // [NM_TERM_PAIRS * 17]
    float* sig;        // [17]
    unsigned short* tri;  // [2016] (a << 8 | b) for a < b < 64
    unsigned* have;    // [Ccap/32 + 2] lazy sweep: keypoints of this rank are in shared memory
    int* fl;           // [64] lazy sweep: ranks to fetch / to test
    unsigned long long* spec;   // [2] lazy sweep: ranks of the current tile with keypoints fetched / added in this attempt
};

__host__ __device__ inline size_t nm_align(size_t x) { return (x + 15) & ~(size_t)15; }

// Fixed-size arrays first, per-candidate arrays after them (`cand_start`).  A CTA whose stream has more candidates than
// its shared-memory tier holds keeps the per-candidate arrays in a global scratch instead (spill mode, nms_body<NT, true>):
// the offsets of those arrays are then relative to the scratch base.
constexpr int NM_FIXED_ARRAYS = 12;          // off[0..11] fixed, off[12..22] per candidate
__host__ __device__ inline size_t nm_carve(unsigned char* base, int Ccap, int Kcap, NmSmem* s, size_t* cand_start = nullptr) {
    const size_t CS = (size_t)Ccap + 1;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = nm_align(off + bytes); return o; };
    const size_t o_tmask = take(64 * 8), o_acc = take(16 * 8), o_spec = take(16);
    const size_t o_keep = take((size_t)Kcap * 4), o_tk = take(64 * 4), o_misc = take(16 * 4);
    const size_t o_l1k = take(NM_LIST * 4), o_l2k = take(NM_LIST * 4);
    const size_t o_terms = take((size_t)NM_TERM_PAIRS * KP * 4), o_sig = take(KP * 4), o_tri = take(2016 * 2), o_fl = take(64 * 4);
    if (cand_start) *cand_start = off;
    const size_t o_score = take((size_t)Ccap * 4), o_rec = take((size_t)Ccap * 4), o_order = take((size_t)Ccap * 4);
    const size_t o_kx = take(KP * CS * 4), o_ky = take(KP * CS * 4);
    const size_t o_vis = take((size_t)Ccap * 4), o_box = take(4 * CS * 4), o_area = take((size_t)Ccap * 4), o_ext = take(4 * CS * 4);
    const size_t o_sup = take((size_t)((Ccap + 31) / 32) * 4 + 8), o_have = take((size_t)((Ccap + 31) / 32) * 4 + 8);
    if (s) {
        s->tmask = (unsigned long long*)(base + o_tmask); s->acc = (unsigned long long*)(base + o_acc);
        s->score = (float*)(base + o_score); s->recidx = (int*)(base + o_rec); s->order = (int*)(base + o_order);
        s->kx = (float*)(base + o_kx); s->ky = (float*)(base + o_ky);
        s->vis = (unsigned*)(base + o_vis); s->box = (float*)(base + o_box); s->area = (float*)(base + o_area);
        s->ext = (float*)(base + o_ext);
        s->sup = (unsigned*)(base + o_sup); s->keep = (int*)(base + o_keep); s->tk = (int*)(base + o_tk);
        s->misc = (int*)(base + o_misc);
        s->l1_key = (unsigned*)(base + o_l1k); s->l2_key = (unsigned*)(base + o_l2k);
        s->terms = (float*)(base + o_terms); s->sig = (float*)(base + o_sig);
        s->tri = (unsigned short*)(base + o_tri);
        s->have = (unsigned*)(base + o_have); s->fl = (int*)(base + o_fl); s->spec = (unsigned long long*)(base + o_spec);
    }
    return off;
}

// Layout as byte offsets, computed on the host and passed as a launch parameter (see tk_from_offsets in tracker_body.cuh).
// off[0..11]: fixed arrays (always shared memory, relative to `fixed`); off[12..22]: per-candidate arrays relative to `cand`
// (== `fixed` unless the layout was made with split = true).
__device__ __forceinline__ void nm_from_offsets(unsigned char* fixed, unsigned char* cand, const SmemOffsets& o, NmSmem& s) {
    s.tmask = reinterpret_cast<unsigned long long*>(fixed + o.off[0]);
    s.acc = reinterpret_cast<unsigned long long*>(fixed + o.off[1]);
    s.spec = reinterpret_cast<unsigned long long*>(fixed + o.off[2]);
    s.keep = reinterpret_cast<int*>(fixed + o.off[3]);
    s.tk = reinterpret_cast<int*>(fixed + o.off[4]);
    s.misc = reinterpret_cast<int*>(fixed + o.off[5]);
    s.l1_key = reinterpret_cast<unsigned*>(fixed + o.off[6]);
    s.l2_key = reinterpret_cast<unsigned*>(fixed + o.off[7]);
    s.terms = reinterpret_cast<float*>(fixed + o.off[8]);
    s.sig = reinterpret_cast<float*>(fixed + o.off[9]);
    s.tri = reinterpret_cast<unsigned short*>(fixed + o.off[10]);
    s.fl = reinterpret_cast<int*>(fixed + o.off[11]);
    s.score = reinterpret_cast<float*>(cand + o.off[12]);
    s.recidx = reinterpret_cast<int*>(cand + o.off[13]);
    s.order = reinterpret_cast<int*>(cand + o.off[14]);
    s.kx = reinterpret_cast<float*>(cand + o.off[15]);
    s.ky = reinterpret_cast<float*>(cand + o.off[16]);
    s.vis = reinterpret_cast<unsigned*>(cand + o.off[17]);
    s.box = reinterpret_cast<float*>(cand + o.off[18]);
    s.area = reinterpret_cast<float*>(cand + o.off[19]);
    s.ext = reinterpret_cast<float*>(cand + o.off[20]);
    s.sup = reinterpret_cast<unsigned*>(cand + o.off[21]);
    s.have = reinterpret_cast<unsigned*>(cand + o.off[22]);
}
// split: offsets of the per-candidate arrays relative to the first of them (for a separate base); returns the size of the
// fixed part in *fixed_bytes and of the per-candidate part in *cand_bytes
inline SmemOffsets nm_offsets(int Ccap, int Kcap, bool split = false, size_t* fixed_bytes = nullptr, size_t* cand_bytes = nullptr) {
    SmemOffsets o{};
    NmSmem t;
    size_t cs = 0;
    const size_t total = nm_carve(nullptr, Ccap, Kcap, &t, &cs);
    const void* arr[23] = {t.tmask, t.acc, t.spec, t.keep, t.tk, t.misc, t.l1_key, t.l2_key, t.terms, t.sig, t.tri, t.fl,
                           t.score, t.recidx, t.order, t.kx, t.ky, t.vis, t.box, t.area, t.ext, t.sup, t.have};
    for (int i = 0; i < 23; ++i) {
        size_t v = (size_t)(uintptr_t)arr[i];
        if (split && i >= NM_FIXED_ARRAYS) v -= cs;
        o.off[i] = (unsigned)v;
    }
    if (fixed_bytes) *fixed_bytes = cs;
    if (cand_bytes) *cand_bytes = total - cs;
    return o;
}


__device__ __forceinline__ bool is_sup(const unsigned* sup, int r) { return (sup[r >> 5] >> (r & 31)) & 1u; }

// Warp-aggregated append to a work list.  Must be reached by all 32 lanes of the warp.
__device__ __forceinline__ void list_push(int* counter, unsigned* keys, bool want, unsigned key) {
    const unsigned bm = __ballot_sync(FULL, want);
    if (bm == 0u) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(bm) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(bm));
    base = __shfl_sync(FULL, base, leader);
    if (want) {
        const int pos = base + __popc(bm & ((1u << lane) - 1u));
        keys[pos] = key;
    }
}

// Box IoU exactly as gpu_postprocess.cu:113-131 (recomputed where needed instead of stored).
__device__ __forceinline__ float pair_iou(const NmSmem& s, int CS, int i, int j) {
    const float xi1 = s.box[0 * CS + i], yi1 = s.box[1 * CS + i], xi2 = s.box[2 * CS + i], yi2 = s.box[3 * CS + i];
    const float xj1 = s.box[0 * CS + j], yj1 = s.box[1 * CS + j], xj2 = s.box[2 * CS + j], yj2 = s.box[3 * CS + j];
    const float ix1 = pb_max(xi1, xj1), iy1 = pb_max(yi1, yj1);
    const float ix2 = pb_min(xi2, xj2), iy2 = pb_min(yi2, yj2);
    const float iw = pb_max(0.0f, ix2 - ix1), ih = pb_max(0.0f, iy2 - iy1);
    const float inter = iw * ih;
    const float uni = s.area[i] + s.area[j] - inter;
    // inter == +0 gives iou == +0 whenever uni > 0 (and 0 otherwise): skip the division then.
    return (uni > 0 && inter > 0.0f) ? (inter / uni) : 0.0f;
}

// Stage A of the reference's pair test (gpu_postprocess.cu:113-137) on ranks i, j:
// returns 1 (IoU > thr: overlap), 0 (provably no overlap) or 2 (undecided).  Symmetric in (i, j).
// iou, t8 (= 2 * scale^2 * 4) and the common visibility mask are handed back for the later stages.
__device__ __forceinline__ int nms_stage_a2(const NmSmem& s, int CS, int i, int j, float thr, float& iou, float& t8, unsigned& vp) {
    const float area_i = s.area[i], area_j = s.area[j];
    iou = pair_iou(s, CS, i, j);
    if (iou > thr) return 1;
    vp = s.vis[i] & s.vis[j];
    const int cnt = __popc(vp);
    if (cnt < 3) return 0;                                               // :162
    float scale_sq = pb_max(area_i, area_j);
    if (scale_sq < 32.0f * 32.0f) scale_sq = 32.0f * 32.0f;
    t8 = 2.0f * scale_sq * 4.0f;
    const float need = (iou > 0.2f) ? pb_min(thr, 0.4f) : thr;           // smallest OKS that can fire (:165)
    // Geometric bound.  If the keypoint extents of i and j are separated by more than
    // r = sqrt(3.1 * t8 * sigma_max^2) along x or y, every keypoint pair has
    // d2 > 3.003 * t8 * sigma_k^2 and contributes < 0.05: oks < 0.05 <= need - 0.002.
    if (need > 0.06f) {
        const float r2 = 3.1f * t8 * (0.107f * 0.107f);
        const float gx = pb_max(s.ext[0 * CS + i] - s.ext[1 * CS + j], s.ext[0 * CS + j] - s.ext[1 * CS + i]);
        const float gy = pb_max(s.ext[2 * CS + i] - s.ext[3 * CS + j], s.ext[2 * CS + j] - s.ext[3 * CS + i]);
        if ((gx > 0.0f && gx * gx > r2) || (gy > 0.0f && gy * gy > r2)) return 0;
    }
    return 2;
}
__device__ __forceinline__ int nms_stage_a(const NmSmem& s, int CS, int i, int j, float thr) {
    float iou = 0.0f, t8 = 0.0f;
    unsigned vp = 0u;
    return nms_stage_a2(s, CS, i, j, thr, iou, t8, vp);
}

// Stage B: a keypoint with d2 >= 3.003*den contributes exp(-d2/den) < 0.05, any other at most 1,
// so oks <= (m + 0.05*(cnt-m))/cnt.  If that is below the smallest threshold that could fire
// (minus a margin far above fp32 rounding) the reference's test is false.  true = still undecided.
__device__ __forceinline__ bool nms_stage_b(const NmSmem& s, int CS, int i, int j, float thr) {
    const float iou = pair_iou(s, CS, i, j);
    const unsigned vis = s.vis[i] & s.vis[j];
    const int cnt = __popc(vis);
    float scale_sq = pb_max(s.area[i], s.area[j]);
    if (scale_sq < 32.0f * 32.0f) scale_sq = 32.0f * 32.0f;
    const float t8 = 2.0f * scale_sq * 4.0f;
    const float need = (iou > 0.2f) ? pb_min(thr, 0.4f) : thr;
    int m = 0;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        if (vis & (1u << k)) {
            const float dx = s.kx[k * CS + i] - s.kx[k * CS + j];
            const float dy = s.ky[k * CS + i] - s.ky[k * CS + j];
            const float d2 = dx * dx + dy * dy;
            const float sg = kSigmas[k];
            m += (d2 < 3.003f * (t8 * sg * sg)) ? 1 : 0;
        }
    }
    return !((float)m + 0.05f * (float)(cnt - m) < (need - 0.002f) * (float)cnt);
}

// Stages B and C over the pairs queued in list 1.  MODE 0: pairs inside the tile -> tile mask;
// MODE 1: survivor x later rank -> suppressed bitmap; MODE 2: verification of the fast path
// (any overlapping pair raises misc[5]).  All threads of the CTA call this.
template <int NM_THREADS, int MODE>
__device__ __forceinline__ void resolve_lists(const NmSmem& s, int CS, int t0, float thr, int tid) {
    const int n1 = s.misc[3];
    // stage B (dense lanes); uniform trip count because list_push uses ballots
    for (int e0 = 0; e0 < n1; e0 += NM_THREADS) {
        const int e = e0 + tid;
        bool und = false;
        unsigned key = 0u;
        if (e < n1) {
            key = s.l1_key[e];
            und = nms_stage_b(s, CS, (int)(key >> 16), (int)(key & 0xffffu), thr);
        }
        list_push(&s.misc[4], s.l2_key, und, key);
    }
    __syncthreads();
    // stage C: one thread per (pair, keypoint) evaluates exp(-d2 / (2*scale*4*sigma^2))
    // (gpu_postprocess.cu:151-157); one thread per pair adds the 17 terms in keypoint order (the
    // reference's summation order; an invisible keypoint adds an exact +0) and applies :162-167.
    const int n2 = s.misc[4];
    for (int cbase = 0; cbase < n2; cbase += NM_TERM_PAIRS) {
        const int ncur = (n2 - cbase) < NM_TERM_PAIRS ? (n2 - cbase) : NM_TERM_PAIRS;
        for (int idx = tid; idx < ncur * KP; idx += NM_THREADS) {
            const int e = idx / KP, k = idx - e * KP;
            const unsigned key = s.l2_key[cbase + e];
            const int i = (int)(key >> 16), j = (int)(key & 0xffffu);
            float term = 0.0f;
            if ((s.vis[i] & s.vis[j]) & (1u << k)) {
                float scale_sq = pb_max(s.area[i], s.area[j]);
                if (scale_sq < 32.0f * 32.0f) scale_sq = 32.0f * 32.0f;
                const float t8 = 2.0f * scale_sq * 4.0f;
                const float dx = s.kx[k * CS + i] - s.kx[k * CS + j];
                const float dy = s.ky[k * CS + i] - s.ky[k * CS + j];
                const float d2 = dx * dx + dy * dy;
                const float sg = s.sig[k];
                term = pb_expf(-d2 / (t8 * sg * sg));
            }
            s.terms[idx] = term;
        }
        __syncthreads();
        for (int e = tid; e < ncur; e += NM_THREADS) {
            const unsigned key = s.l2_key[cbase + e];
            const int i = (int)(key >> 16), j = (int)(key & 0xffffu);
            const int cnt = __popc(s.vis[i] & s.vis[j]);
            float sum = 0.0f;
#pragma unroll
            for (int k = 0; k < KP; ++k) sum += s.terms[e * KP + k];
            const float oks = sum / (float)cnt;
            const float iou = pair_iou(s, CS, i, j);
            if ((oks > thr) || (oks > 0.4f && iou > 0.2f)) {
                if (MODE == 0) atomicOr(&s.tmask[i - t0], 1ull << (j - t0));
                else if (MODE == 1) atomicOr(&s.sup[j >> 5], 1u << (j & 31));
                else s.misc[5] = 1;
            }
        }
        __syncthreads();
    }
}

// One stream's rank + NMS, executed by the NM_THREADS threads of one CTA.  `s`: the working arrays (shared memory; in spill
// mode the per-candidate ones live in a global scratch and Ccap is the full candidate cap).  fuse_det / fuse_score (may be
// null): shared-memory arrays of the tracker stage that receive the first fuse_cap kept detections (fused per-stream kernel).
// Returns the number of kept detections (uniform over the CTA).
template <int NM_THREADS>
__device__ __forceinline__ int nms_body(const NmSmem& s, const float* __restrict__ heads, int N, int lazy, const CandScratch& cs, int nseg,
                                        int segcap, int Ccap, int Kcap, float nms_thr, const PostBuffers& out, const int b, const int nstreams,
                                        float* fuse_det, float* fuse_score, int fuse_cap) {
    const int CS = Ccap + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* recs = cs.records + (size_t)b * nseg * segcap * HEAD_ROWS;
    const int* ancs = cs.anchors + (size_t)b * nseg * segcap;
    const float* head = heads + (size_t)b * HEAD_ROWS * N;
    float* o_pose = out.det_poses + (size_t)b * Kcap * POSE_F;
    float* o_box = out.det_bboxes + (size_t)b * Kcap * 4;
    float* o_score = out.det_scores + (size_t)b * Kcap;
    int* o_slot = out.keep_slots + (size_t)b * Kcap;
    int* o_anchor = out.keep_anchors + (size_t)b * Kcap;

    unsigned long long t_stamp = 0;
    if (tid == 0) t_stamp = globaltimer_ns();
    if (tid == 0 && out.dbg) out.dbg[((size_t)out.dbg_slot * nstreams + b) * 6 + 3] = t_stamp;
    auto stamp = [&](int slot) {   // thread 0 accumulates in shared memory, flushed once at the end
#ifndef PB_NO_STAMPS
        if (tid == 0) { const unsigned long long now = globaltimer_ns(); s.acc[slot] += now - t_stamp; t_stamp = now; }
#endif
    };

    // ---------------- 0. candidate list = concatenation of the segment lists (anchor order) -----
    if (tid < 16) { s.misc[tid] = 0; s.acc[tid] = 0ull; }
    if (tid < KP) s.sig[tid] = kSigmas[tid];
#pragma unroll 1
    for (int i = tid; i < (Ccap + 31) / 32 + 2; i += NM_THREADS) s.sup[i] = 0u;
#pragma unroll 1
    for (int i = tid; i < Ccap; i += NM_THREADS) s.vis[i] = 0u;
#pragma unroll 1
    for (int p = tid; p < 2016; p += NM_THREADS) {       // triangular pair table: row a holds the 63-a pairs (a, b > a)
        int a = (int)((127.0f - sqrtf(16129.0f - 8.0f * (float)p)) * 0.5f);    // first pair of row a is a*(127-a)/2
        if (a < 0) a = 0;
        while (a > 0 && a * (127 - a) / 2 > p) --a;
        while ((a + 1) * (126 - a) / 2 <= p) ++a;
        s.tri[p] = (unsigned short)((a << 8) | (a + 1 + (p - a * (127 - a) / 2)));
    }
    int* segstart = reinterpret_cast<int*>(s.l1_key);    // [nseg + 1]; the pair list is idle until the sweep
    if (warp == 0) {
        int run = 0;
        for (int sg0 = 0; sg0 < nseg; sg0 += 32) {
            const int sg = sg0 + lane;
            const int n = (sg < nseg) ? cs.counts[b * nseg + sg] : 0;
            int incl = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
            if (sg < nseg) segstart[sg] = run + incl - n;
            run += __shfl_sync(FULL, incl, 31);
        }
        if (lane == 0) { segstart[nseg] = run; s.misc[0] = run < Ccap ? run : Ccap; }
    }
    __syncthreads();
    const int C = s.misc[0];
#pragma unroll 1
    for (int c = tid; c < C; c += NM_THREADS) {          // slot c -> record of its segment, and its score
        int sg = 0;
        while (segstart[sg + 1] <= c) ++sg;
        const int ri = sg * segcap + (c - segstart[sg]);
        s.recidx[c] = ri;
        s.score[c] = recs[(size_t)ri * HEAD_ROWS + 4];
    }
    __syncthreads();
    stamp(0);

    // ---------------- 1. rank by (score desc, slot asc)  (:178-203, R2) ----------------
    // NaN confidences pass the filter (:51) and the reference's insertion sort never moves anything across one
    // (`score < key` is false on either side): a NaN stays at its slot index and the runs between NaNs are sorted
    // on their own.  The all-pairs count below is that order only without NaNs, so their presence selects the
    // segmented count (still a permutation: rank = slots of earlier segments + rank inside the own segment).
    {
        bool nan_here = false;
        for (int i = tid; i < C; i += NM_THREADS) nan_here |= (s.score[i] != s.score[i]);
        if (nan_here) s.misc[8] = 1;                 // misc[] was cleared above, before two barriers
    }
    __syncthreads();
    if (s.misc[8] == 0) {
        for (int i = tid; i < C; i += NM_THREADS) {
            const float si = s.score[i];
            int rank = 0;
            for (int j = 0; j < C; ++j) {
                const float sj = s.score[j];
                rank += (sj > si || (sj == si && j < i)) ? 1 : 0;
            }
            s.order[rank] = i;
        }
    } else {
        for (int i = tid; i < C; i += NM_THREADS) {
            const float si = s.score[i];
            int rank = i;
            if (si == si) {
                int lo = i, hi = i;                  // the NaN-free run [lo, hi] around slot i
                while (lo > 0 && s.score[lo - 1] == s.score[lo - 1]) --lo;
                while (hi + 1 < C && s.score[hi + 1] == s.score[hi + 1]) ++hi;
                rank = lo;
                for (int j = lo; j <= hi; ++j) {
                    const float sj = s.score[j];
                    rank += (sj > si || (sj == si && j < i)) ? 1 : 0;
                }
            }
            s.order[rank] = i;
        }
    }
    __syncthreads();
    stamp(1);

    // ---------------- L. lazy sweep: IoU first, keypoints only where an OKS test is unavoidable ----------------
    // Same result as the complete sweep below (sections 2-3), different evaluation order.  A candidate
    // needs its keypoints only if no kept candidate removes it by the IoU rule (:113-137) — with 5-9
    // near-duplicate anchors per person that is ~15 % of the candidates — so the decode kernel has
    // gathered box rows only and keypoints are fetched from the head tensor here, on demand.  Used when
    // the head lives in page-locked host memory (every sector is a PCIe read).  Per tile of 64 ranks:
    //   live   = tile ranks no kept candidate of an earlier tile removes by IoU (struck eagerly, step 5);
    //   round 1: S = survivors of a box-only greedy pass over live; fetch S's keypoints; complete tests of
    //            S against every earlier kept candidate and inside S; walk the tile in rank order: a
    //            member of S is kept unless struck; a live rank outside S must be struck by a kept rank
    //            of this tile — if one is not (the rank that shadowed it fell to an OKS rule), round 1 is
    //            abandoned;
    //   round 2: fetch the keypoints of all live ranks and decide the tile with every pair tested, exactly
    //            as the complete sweep does.
    // Every decision is the one the reference's sweep takes (:209-242): a rank is dropped only when a
    // kept earlier rank overlaps it by the complete rule, kept only when all kept earlier ranks were
    // tested against it.
    if (lazy) {
#pragma unroll 1
    for (int it = tid; it < C * 4; it += NM_THREADS) {
        const int r = it >> 2, e = it & 3;
        s.box[e * CS + r] = recs[(size_t)s.recidx[s.order[r]] * HEAD_ROWS + e];
    }
#pragma unroll 1
    for (int i = tid; i < (Ccap + 31) / 32 + 2; i += NM_THREADS) s.have[i] = 0u;
    __syncthreads();
#pragma unroll 1
    for (int r = tid; r < C; r += NM_THREADS) {   // cx,cy,w,h -> corners (:66-69), area (:128)
        const float cx = s.box[0 * CS + r], cy = s.box[1 * CS + r], w = s.box[2 * CS + r], h = s.box[3 * CS + r];
        const float x1 = cx - w * 0.5f, y1 = cy - h * 0.5f, x2 = cx + w * 0.5f, y2 = cy + h * 0.5f;
        s.box[0 * CS + r] = x1; s.box[1 * CS + r] = y1; s.box[2 * CS + r] = x2; s.box[3 * CS + r] = y2;
        s.area[r] = (x2 - x1) * (y2 - y1);
    }
    __syncthreads();
    if (lazy == 2) {
        // "deferred" flavour: the decode kernel gathered complete records (head in HBM), so all keypoints
        // go to shared memory now and nothing is fetched later; what remains of the lazy sweep is its
        // evaluation ORDER — IoU strikes first, OKS tests only for the ranks still alive at their own tile.
#pragma unroll 1
        for (int it = tid; it < C * POSE_F; it += NM_THREADS) {
            const int r = it / POSE_F, e = it - r * POSE_F;
            const float v = recs[(size_t)s.recidx[s.order[r]] * HEAD_ROWS + 5 + e];
            const int k = e / 3, comp = e - 3 * k;
            if (comp == 0) s.kx[k * CS + r] = v;
            else if (comp == 1) s.ky[k * CS + r] = v;
            else if (v > 0.2f) atomicOr(&s.vis[r], 1u << k);
        }
        __syncthreads();
#pragma unroll 1
        for (int r = tid; r < C; r += NM_THREADS) {
            float lx = s.kx[r], hx = lx, ly = s.ky[r], hy = ly;
#pragma unroll
            for (int k = 1; k < KP; ++k) {
                const float x = s.kx[k * CS + r], y = s.ky[k * CS + r];
                lx = fminf(lx, x); hx = fmaxf(hx, x); ly = fminf(ly, y); hy = fmaxf(hy, y);
            }
            s.ext[0 * CS + r] = lx; s.ext[1 * CS + r] = hx; s.ext[2 * CS + r] = ly; s.ext[3 * CS + r] = hy;
        }
#pragma unroll 1
        for (int i = tid; i < (Ccap + 31) / 32 + 2; i += NM_THREADS) s.have[i] = 0xffffffffu;
        __syncthreads();
    }
    stamp(2);
    // fetch the keypoints of the ranks listed in s.fl[0..nf): head -> candidate record (for the output
    // stage) + shared memory (rank-indexed), then their keypoint extents
    auto fetch_list = [&](int nf) {
#pragma unroll 1
        for (int it = tid; it < nf * POSE_F; it += NM_THREADS) {
            const int p = it / POSE_F, e = it - p * POSE_F;
            const int r = s.fl[p];
            const int ri = s.recidx[s.order[r]];
            const float v = ldg_stream_f(head + (size_t)(5 + e) * N + ancs[ri]);
            recs[(size_t)ri * HEAD_ROWS + 5 + e] = v;                 // verbatim (:75-80)
            const int k = e / 3, comp = e - 3 * k;
            if (comp == 0) s.kx[k * CS + r] = v;
            else if (comp == 1) s.ky[k * CS + r] = v;
            else if (v > 0.2f) atomicOr(&s.vis[r], 1u << k);
        }
        __syncthreads();
#pragma unroll 1
        for (int p = tid; p < nf; p += NM_THREADS) {
            const int r = s.fl[p];
            float lx = s.kx[r], hx = lx, ly = s.ky[r], hy = ly;
#pragma unroll
            for (int k = 1; k < KP; ++k) {
                const float x = s.kx[k * CS + r], y = s.ky[k * CS + r];
                lx = fminf(lx, x); hx = fmaxf(hx, x); ly = fminf(ly, y); hy = fmaxf(hy, y);
            }
            s.ext[0 * CS + r] = lx; s.ext[1 * CS + r] = hx; s.ext[2 * CS + r] = ly; s.ext[3 * CS + r] = hy;
            atomicOr(&s.have[r >> 5], 1u << (r & 31));
        }
        if (tid == 0) s.acc[9] += (unsigned long long)nf;
        __syncthreads();
    };
    // complete tests of the ranks in s.fl[0..nf) against the kept candidates keep[0..nk): IoU is known to
    // be <= thr for these pairs (the eager IoU strikes), the OKS rules may still remove the later rank
    auto cross_tests = [&](int nf, int nk) {
        const int pairs = nf * nk;
        for (int pbase = 0; pbase < pairs; pbase += NM_LIST) {
            const int pend = (pairs - pbase) < NM_LIST ? (pairs - pbase) : NM_LIST;
#pragma unroll 1
            for (int p0 = 0; p0 < pend; p0 += NM_THREADS) {
                int q = 0, i = 0, j = 0;
                if (p0 + tid < pend) {
                    const int p = pbase + p0 + tid;
                    const int a = p / nf;
                    i = s.keep[a]; j = s.fl[p - a * nf];
                    if (!is_sup(s.sup, j)) q = nms_stage_a(s, CS, i, j, nms_thr);
                }
                if (q == 1) atomicOr(&s.sup[j >> 5], 1u << (j & 31));
                list_push(&s.misc[3], s.l1_key, q == 2, ((unsigned)i << 16) | (unsigned)j);
            }
            __syncthreads();
            resolve_lists<NM_THREADS, 1>(s, CS, 0, nms_thr, tid);
            if (tid == 0) { s.misc[3] = 0; s.misc[4] = 0; }
            __syncthreads();
        }
    };
    for (int t0 = 0; t0 < C; t0 += 64) {
        const int tl = (C - t0) < 64 ? (C - t0) : 64;
        const int nk0 = s.misc[1];                                    // kept before this tile
        if (tid < 64) s.tmask[tid] = 0ull;
        if (tid == 0) { s.misc[3] = 0; s.misc[4] = 0; }
        __syncthreads();
        const unsigned long long valid = (tl == 64) ? ~0ull : ((1ull << tl) - 1ull);
        const unsigned long long live = valid & ~((unsigned long long)s.sup[t0 >> 5] | ((unsigned long long)s.sup[(t0 >> 5) + 1] << 32));
        // (1) IoU mask among the live ranks of the tile
#pragma unroll 1
        for (int p = tid; p < 2016; p += NM_THREADS) {
            const unsigned short ab = s.tri[p];
            const int a = ab >> 8, bb = ab & 0xff;
            if (((live >> a) & 1ull) && ((live >> bb) & 1ull) && pair_iou(s, CS, t0 + a, t0 + bb) > nms_thr)
                atomicOr(&s.tmask[a], 1ull << bb);
        }
        __syncthreads();
        if (lazy == 2) {
            if (tid == 0) s.misc[7] = 1;                              // straight to the full tile step
            __syncthreads();
        } else {
        // (2) round 1: speculative box-only greedy pass -> S
        if (tid == 0) {
            unsigned long long rem = live, spec = 0ull;
            int room = Kcap - nk0;
            while (rem != 0ull && room > 0) {
                const int a = __ffsll((long long)rem) - 1;
                spec |= 1ull << a;
                --room;
                rem &= ~s.tmask[a];
                rem &= ~(1ull << a);
            }
            s.spec[0] = spec;          // ranks whose keypoints are (being) fetched
            s.spec[1] = spec;          // the ones added in this attempt
        }
        __syncthreads();
        // Two attempts: S, then S plus every rank the walk found unshadowed together with the ranks it
        // shadows itself (its own near-duplicates, which are needed as soon as it falls too).
        for (int attempt = 0; attempt < 2; ++attempt) {
            const unsigned long long spec = s.spec[0], added = s.spec[1];
            if (tid == 0) {
                int n = 0;
                for (unsigned long long m = added; m; m &= m - 1ull) s.fl[n++] = t0 + __ffsll((long long)m) - 1;
                s.misc[6] = n;
            }
            __syncthreads();
            const int na_ = s.misc[6];
            fetch_list(na_);
            cross_tests(na_, nk0);                                    // (3) the added ranks against every earlier kept candidate
            // (4) pairs inside the fetched set that involve an added rank: IoU is in the mask already, the OKS rules may fire
#pragma unroll 1
            for (int p0 = 0; p0 < 2016; p0 += NM_THREADS) {
                const int p = p0 + tid;
                int q = 0, a = 0, bb = 0;
                if (p < 2016) {
                    const unsigned short ab = s.tri[p];
                    a = ab >> 8; bb = ab & 0xff;
                    if (((spec >> a) & 1ull) && ((spec >> bb) & 1ull) && (((added >> a) | (added >> bb)) & 1ull) &&
                        !((s.tmask[a] >> bb) & 1ull) && !is_sup(s.sup, t0 + a) && !is_sup(s.sup, t0 + bb))
                        q = nms_stage_a(s, CS, t0 + a, t0 + bb, nms_thr);
                }
                list_push(&s.misc[3], s.l1_key, q == 2, ((unsigned)(t0 + a) << 16) | (unsigned)(t0 + bb));
            }
            __syncthreads();
            resolve_lists<NM_THREADS, 0>(s, CS, t0, nms_thr, tid);
            // (5) walk the tile in rank order
            if (tid == 0) {
                unsigned long long supt = (unsigned long long)s.sup[t0 >> 5] | ((unsigned long long)s.sup[(t0 >> 5) + 1] << 32);
                unsigned long long extra = 0ull;
                int nk = nk0, ntk = 0;
                for (unsigned long long m = live; m && nk < Kcap; m &= m - 1ull) {
                    const int a = __ffsll((long long)m) - 1;
                    if ((supt >> a) & 1ull) continue;                // struck: by an earlier tile (OKS) or by a kept rank of this tile
                    if (!((spec >> a) & 1ull)) {                     // shadowed only by a rank that fell: needs its own tests,
                        extra |= (1ull << a) | (s.tmask[a] & live);  // and so may the ranks it shadows itself
                        supt |= s.tmask[a];
                        continue;
                    }
                    s.keep[nk++] = t0 + a;
                    s.tk[ntk++] = t0 + a;
                    supt |= s.tmask[a];
                }
                extra &= ~spec;
                s.misc[7] = (extra != 0ull) ? 1 : 0;
                if (extra == 0ull) {
                    s.sup[t0 >> 5] = (unsigned)(supt | ~live);
                    s.sup[(t0 >> 5) + 1] = (unsigned)((supt | ~live) >> 32);
                    s.misc[1] = nk; s.misc[2] = ntk;
                } else {
                    s.spec[0] = spec | extra;
                    s.spec[1] = extra;
                }
                s.misc[3] = 0; s.misc[4] = 0;
            }
            __syncthreads();
            if (!s.misc[7]) break;
        }
        }   // speculative attempts
        if (s.misc[7]) {
            // round 2: keypoints for every live rank, every pair tested (the complete sweep's tile step)
            if (tid == 0) {
                int nf = 0;
                for (unsigned long long m = live; m; m &= m - 1ull) {
                    const int a = __ffsll((long long)m) - 1;
                    if (lazy == 2 || !is_sup(s.have, t0 + a)) s.fl[nf++] = t0 + a;
                }
                s.misc[6] = nf;
                s.acc[8] += 1ull;
            }
            __syncthreads();
            const int nf2 = s.misc[6];
            if (lazy != 2) fetch_list(nf2);
            cross_tests(nf2, nk0);                                    // the newly fetched ranks against the earlier kept
            if (tid < 64) s.tmask[tid] = 0ull;
            __syncthreads();
#pragma unroll 1
            for (int p0 = 0; p0 < 2016; p0 += NM_THREADS) {
                const int p = p0 + tid;
                int q = 0, a = 0, bb = 0;
                if (p < 2016) {
                    const unsigned short ab = s.tri[p];
                    a = ab >> 8; bb = ab & 0xff;
                    if (bb < tl && !is_sup(s.sup, t0 + a) && !is_sup(s.sup, t0 + bb))
                        q = nms_stage_a(s, CS, t0 + a, t0 + bb, nms_thr);
                }
                if (q == 1) atomicOr(&s.tmask[a], 1ull << bb);
                list_push(&s.misc[3], s.l1_key, q == 2, ((unsigned)(t0 + a) << 16) | (unsigned)(t0 + bb));
            }
            __syncthreads();
            resolve_lists<NM_THREADS, 0>(s, CS, t0, nms_thr, tid);
            if (tid == 0) {
                unsigned long long supt = (unsigned long long)s.sup[t0 >> 5] | ((unsigned long long)s.sup[(t0 >> 5) + 1] << 32);
                unsigned long long rem = valid & ~supt;
                int nk = nk0, ntk = 0;
                while (rem != 0ull && nk < Kcap) {                   // :224 "num_keep < 256"
                    const int a = __ffsll((long long)rem) - 1;
                    s.keep[nk++] = t0 + a;
                    s.tk[ntk++] = t0 + a;
                    supt |= s.tmask[a];
                    rem &= ~supt;
                    rem &= ~(1ull << a);
                }
                s.sup[t0 >> 5] = (unsigned)supt;
                s.sup[(t0 >> 5) + 1] = (unsigned)(supt >> 32);
                s.misc[1] = nk; s.misc[2] = ntk; s.misc[3] = 0; s.misc[4] = 0;
            }
            __syncthreads();
        }
        const int ntk = s.misc[2];
        if (s.misc[1] >= Kcap) break;
        // (6) the tile's kept ranks strike the later ranks by the IoU rule (their OKS tests wait for the keypoints)
        const int j0 = t0 + 64, rem_n = C - j0;
        if (rem_n > 0 && ntk > 0) {
#pragma unroll 1
            for (int p = tid; p < ntk * rem_n; p += NM_THREADS) {
                const int ai = p / rem_n, j = j0 + (p - ai * rem_n);
                if (!is_sup(s.sup, j) && pair_iou(s, CS, s.tk[ai], j) > nms_thr) atomicOr(&s.sup[j >> 5], 1u << (j & 31));
            }
        }
        __syncthreads();
    }
    __syncthreads();
    stamp(3);
    } else {

    // ---------------- 2. records -> shared memory SoA in rank order ----------------
    // eight independent loads per thread in flight (the stores to shared memory would otherwise keep the compiler
    // from moving the next load above them: one L2 round trip per element)
    {
        const int total = C * HEAD_ROWS;
#pragma unroll 1
        for (int it0 = tid; it0 < total; it0 += NM_THREADS * 8) {
            float v8[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int it = it0 + u * NM_THREADS;
                v8[u] = 0.0f;
                if (it < total) {
                    const int r = it / HEAD_ROWS, row = it - r * HEAD_ROWS;
                    v8[u] = recs[(size_t)s.recidx[s.order[r]] * HEAD_ROWS + row];
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int it = it0 + u * NM_THREADS;
                if (it < total) {
                    const int r = it / HEAD_ROWS, row = it - r * HEAD_ROWS;
                    const float v = v8[u];
                    if (row < 4) {
                        s.box[row * CS + r] = v;
                    } else if (row > 4) {
                        const int k = (row - 5) / 3, comp = (row - 5) - 3 * k;
                        if (comp == 0) s.kx[k * CS + r] = v;
                        else if (comp == 1) s.ky[k * CS + r] = v;
                        else if (v > 0.2f) atomicOr(&s.vis[r], 1u << k);
                    }
                }
            }
        }
    }
    __syncthreads();
#pragma unroll 1
    for (int r = tid; r < C; r += NM_THREADS) {   // cx,cy,w,h -> corners (:66-69), area (:128), keypoint extents
        const float cx = s.box[0 * CS + r], cy = s.box[1 * CS + r], w = s.box[2 * CS + r], h = s.box[3 * CS + r];
        const float x1 = cx - w * 0.5f, y1 = cy - h * 0.5f, x2 = cx + w * 0.5f, y2 = cy + h * 0.5f;
        s.box[0 * CS + r] = x1; s.box[1 * CS + r] = y1; s.box[2 * CS + r] = x2; s.box[3 * CS + r] = y2;
        s.area[r] = (x2 - x1) * (y2 - y1);
        float lx = s.kx[r], hx = lx, ly = s.ky[r], hy = ly;
#pragma unroll
        for (int k = 1; k < KP; ++k) {
            const float x = s.kx[k * CS + r], y = s.ky[k * CS + r];
            lx = fminf(lx, x); hx = fmaxf(hx, x); ly = fminf(ly, y); hy = fmaxf(hy, y);
        }
        s.ext[0 * CS + r] = lx; s.ext[1 * CS + r] = hx; s.ext[2 * CS + r] = ly; s.ext[3 * CS + r] = hy;
    }
    stamp(2);

    // ---------------- 3. greedy suppression in rank order (:88-172 + :209-242) ----------------
    for (int t0 = 0; t0 < C; t0 += 64) {
        const int tl = (C - t0) < 64 ? (C - t0) : 64;
        if (tid < 64) s.tmask[tid] = 0ull;
        if (tid == 0) { s.misc[3] = 0; s.misc[4] = 0; }
        __syncthreads();
        // (a) pairs inside the tile
#pragma unroll 1
        for (int p0 = 0; p0 < 2016; p0 += NM_THREADS) {
            const int p = p0 + tid;
            int q = 0, a = 0, bb = 0;
            if (p < 2016) {
                const unsigned short ab = s.tri[p];
                a = ab >> 8; bb = ab & 0xff;
                if (bb < tl && !is_sup(s.sup, t0 + a) && !is_sup(s.sup, t0 + bb))
                    q = nms_stage_a(s, CS, t0 + a, t0 + bb, nms_thr);
            }
            if (q == 1) atomicOr(&s.tmask[a], 1ull << bb);
            list_push(&s.misc[3], s.l1_key, q == 2, ((unsigned)(t0 + a) << 16) | (unsigned)(t0 + bb));
        }
        __syncthreads();
        resolve_lists<NM_THREADS, 0>(s, CS, t0, nms_thr, tid);
        // (b) serial greedy over the tile, one step per survivor
        if (tid == 0) {
            unsigned long long supt = (unsigned long long)s.sup[t0 >> 5] | ((unsigned long long)s.sup[(t0 >> 5) + 1] << 32);
            const unsigned long long valid = (tl == 64) ? ~0ull : ((1ull << tl) - 1ull);
            unsigned long long rem = valid & ~supt;
            int nk = s.misc[1], ntk = 0;
            while (rem != 0ull && nk < Kcap) {                   // :224 "num_keep < 256"
                const int a = __ffsll((long long)rem) - 1;
                s.keep[nk++] = t0 + a;
                s.tk[ntk++] = t0 + a;
                supt |= s.tmask[a];                              // bits b > a only
                rem &= ~supt;
                rem &= ~(1ull << a);
            }
            s.sup[t0 >> 5] = (unsigned)supt;
            s.sup[(t0 >> 5) + 1] = (unsigned)(supt >> 32);
            s.misc[1] = nk; s.misc[2] = ntk; s.misc[3] = 0; s.misc[4] = 0;
        }
        __syncthreads();
        const int ntk = s.misc[2];
        if (s.misc[1] >= Kcap) break;
        // (c) the tile's survivors strike the remaining ranks, NM_LIST candidate pairs per round
        const int j0 = t0 + 64, rem_n = C - j0;
        if (rem_n > 0 && ntk > 0) {
            const int pairs = ntk * rem_n;
            for (int pbase = 0; pbase < pairs; pbase += NM_LIST) {
                const int pend = (pairs - pbase) < NM_LIST ? (pairs - pbase) : NM_LIST;
#pragma unroll 1
                for (int p0 = 0; p0 < pend; p0 += NM_THREADS) {
                    int q = 0, i = 0, j = 0;
                    if (p0 + tid < pend) {
                        const int p = pbase + p0 + tid;
                        const int ai = p / rem_n;
                        j = j0 + (p - ai * rem_n);
                        i = s.tk[ai];
                        if (!is_sup(s.sup, j)) q = nms_stage_a(s, CS, i, j, nms_thr);
                    }
                    if (q == 1) atomicOr(&s.sup[j >> 5], 1u << (j & 31));
                    list_push(&s.misc[3], s.l1_key, q == 2, ((unsigned)i << 16) | (unsigned)j);
                }
                __syncthreads();
                resolve_lists<NM_THREADS, 1>(s, CS, t0, nms_thr, tid);
                if (tid == 0) { s.misc[3] = 0; s.misc[4] = 0; }
                __syncthreads();
            }
        }
    }
    __syncthreads();
    stamp(3);
    }   // complete sweep
    const int nkeep = s.misc[1];

    // ---------------- 4. kept detections in score order ----------------
#pragma unroll 1
    for (int it = tid; it < nkeep * POSE_F; it += NM_THREADS) {
        const int k = it / POSE_F, e = it - k * POSE_F;
        const float v = recs[(size_t)s.recidx[s.order[s.keep[k]]] * HEAD_ROWS + 5 + e];   // verbatim (:75-80)
        o_pose[it] = v;
        if (fuse_det != nullptr && k < fuse_cap) fuse_det[it] = v;
    }
#pragma unroll 1
    for (int it = tid; it < nkeep * 4; it += NM_THREADS) {
        const int k = it >> 2, e = it & 3;
        o_box[it] = s.box[e * CS + s.keep[k]];
    }
#pragma unroll 1
    for (int k = tid; k < nkeep; k += NM_THREADS) {
        const int slot = s.order[s.keep[k]];
        const float sc = s.score[slot];
        o_score[k] = sc;
        if (fuse_score != nullptr && k < fuse_cap) fuse_score[k] = sc;
        o_slot[k] = slot;
        o_anchor[k] = ancs[s.recidx[slot]];
    }
    stamp(4);
    if (tid == 0) { out.num_keep[b] = nkeep; out.num_cand[b] = C; s.acc[7] = 1ull; }
    if (tid == 0 && out.dbg) out.dbg[((size_t)out.dbg_slot * nstreams + b) * 6 + 4] = globaltimer_ns();
    __syncthreads();
    if (tid < 16 && s.acc[tid] != 0ull) out.stage_ns[(size_t)b * 16 + tid] += s.acc[tid];
    return nkeep;
}

}  // namespace pb
