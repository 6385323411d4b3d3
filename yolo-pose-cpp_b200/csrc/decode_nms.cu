// decode_nms.cu — head decode + confidence filter + OKS/IoU pose-NMS, one CTA per stream.
//
// Replaces GPUPostprocess::process (reference src/cuda/gpu_postprocess.cu:366-476) and
// its six kernels (:30-313) plus three host synchronisations with ONE launch for all B
// streams.  Per stream:
//   1. the confidence row [N] is streamed with 128-bit loads; hits are compacted in
//      ascending anchor order with warp popcounts + one block scan (rule R1 replaces the
//      reference's atomicAdd arrival order, :54);
//   2. candidates are ranked by (score desc, slot asc) == the reference's stable insertion
//      sort (:178-203, rule R2);
//   3. the other 55 head rows are gathered only at the candidate anchors (32-byte
//      sectors), straight into shared memory in rank order, SoA;
//   4. greedy suppression runs tile-by-tile (64 ranks): intra-tile bitmask via 64-bit
//      shared atomics, a serial resolve of the tile, then the tile's survivors strike
//      the remaining ranks in parallel.  Only mask bits the reference's sweep (:209-242)
//      would actually read are evaluated; the pair test is the reference's (:88-172),
//      guarded by an exact upper-bound filter that skips the 17 exponentials when the
//      pair provably cannot reach the threshold;
//   5. kept detections are written in score order (== kernelCompactDetections/CopyBack,
//      :248-313) together with their candidate slots and anchor ids.
// Arithmetic goes through pb_math.h (see there), compiled with --fmad=false.
#include "pb_common.cuh"

namespace pb {

constexpr int DN_THREADS = 512;
constexpr int DN_WARPS = DN_THREADS / 32;
constexpr int DN_MAX_CHUNKS = 32;            // N <= 32 * 512 * 4 = 65536 anchors
constexpr unsigned FULL = 0xffffffffu;
constexpr int DN_LIST = 2048;                // undecided pairs per round (>= 64*63/2)
constexpr int DN_TERM_PAIRS = 128;           // pairs whose 17 terms are evaluated at once

struct DnSmem {
    float* score;      // [Ccap] by slot
    int* anchor;       // [Ccap] by slot
    int* order;        // [Ccap] rank -> slot
    float* kx;         // [17][Ccap] by rank
    float* ky;         // [17][Ccap]
    unsigned* vis;     // [Ccap] bit k: conf_k > 0.2
    float* box;        // [4][Ccap] cx,cy,w,h then x1,y1,x2,y2
    float* area;       // [Ccap]
    float* ext;        // [4][Ccap] keypoint extents lx, hx, ly, hy over all 17 keypoints
    unsigned* sup;     // [Ccap/32]
    int* keep;         // [Kcap] kept ranks
    unsigned long long* tmask;  // [64]
    int* cnt;          // [DN_MAX_CHUNKS * DN_WARPS]
    int* tk;           // [64] ranks kept in the current tile
    int* misc;         // [8]: 0 total, 1 nkeep, 2 ntk, 3 list length
    unsigned* list_key; // [DN_LIST] undecided pairs (rank_i << 16 | rank_j)
    float* list_iou;    // [DN_LIST]
    float* terms;       // [DN_TERM_PAIRS * 17] per-keypoint OKS terms of the pairs being resolved
    float* sig;         // [17] COCO sigmas (shared copy: indexed per lane)
    unsigned long long* acc;  // [16] stage telemetry accumulators
};

__host__ __device__ inline size_t dn_align(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline size_t dn_carve(unsigned char* base, int Ccap, int Kcap, DnSmem* s) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = dn_align(off + bytes); return o; };
    size_t o_tmask = take(64 * 8), o_acc = take(16 * 8);
    size_t o_score = take((size_t)Ccap * 4), o_anchor = take((size_t)Ccap * 4), o_order = take((size_t)Ccap * 4);
    size_t o_kx = take((size_t)KP * Ccap * 4), o_ky = take((size_t)KP * Ccap * 4);
    size_t o_vis = take((size_t)Ccap * 4), o_box = take((size_t)4 * Ccap * 4), o_area = take((size_t)Ccap * 4);
    size_t o_ext = take((size_t)4 * Ccap * 4);
    size_t o_sup = take((size_t)((Ccap + 31) / 32) * 4 + 8), o_keep = take((size_t)Kcap * 4);
    size_t o_cnt = take((size_t)DN_MAX_CHUNKS * DN_WARPS * 4), o_tk = take(64 * 4), o_misc = take(8 * 4);
    size_t o_lk = take((size_t)DN_LIST * 4), o_li = take((size_t)DN_LIST * 4);
    size_t o_terms = take((size_t)DN_TERM_PAIRS * KP * 4), o_sig = take(KP * 4);
    if (s) {
        s->tmask = (unsigned long long*)(base + o_tmask);
        s->acc = (unsigned long long*)(base + o_acc);
        s->score = (float*)(base + o_score); s->anchor = (int*)(base + o_anchor); s->order = (int*)(base + o_order);
        s->kx = (float*)(base + o_kx); s->ky = (float*)(base + o_ky);
        s->vis = (unsigned*)(base + o_vis); s->box = (float*)(base + o_box); s->area = (float*)(base + o_area);
        s->ext = (float*)(base + o_ext);
        s->sup = (unsigned*)(base + o_sup); s->keep = (int*)(base + o_keep);
        s->cnt = (int*)(base + o_cnt); s->tk = (int*)(base + o_tk); s->misc = (int*)(base + o_misc);
        s->list_key = (unsigned*)(base + o_lk); s->list_iou = (float*)(base + o_li);
        s->terms = (float*)(base + o_terms); s->sig = (float*)(base + o_sig);
    }
    return off;
}

size_t decode_nms_smem_bytes(int max_cand, int max_keep) { return dn_carve(nullptr, max_cand, max_keep, nullptr); }

// The reference's pair test (gpu_postprocess.cu:113-168) on shared-memory SoA data, split in
// two so that the rare expensive part can be compacted and run with full lanes:
//   nms_quick  -> 1 overlap (IoU alone decides), 0 no overlap (exact upper bounds decide),
//                 2 undecided: the 17-exponential OKS has to be evaluated (nms_exact).
// i, j are ranks.  Both functions are symmetric in (i, j) bit for bit.
__device__ __forceinline__ int nms_quick(const DnSmem& s, int Ccap, int i, int j, float thr, float* iou_out) {
    const float xi1 = s.box[0 * Ccap + i], yi1 = s.box[1 * Ccap + i], xi2 = s.box[2 * Ccap + i], yi2 = s.box[3 * Ccap + i];
    const float xj1 = s.box[0 * Ccap + j], yj1 = s.box[1 * Ccap + j], xj2 = s.box[2 * Ccap + j], yj2 = s.box[3 * Ccap + j];
    const float ix1 = pb_max(xi1, xj1), iy1 = pb_max(yi1, yj1);
    const float ix2 = pb_min(xi2, xj2), iy2 = pb_min(yi2, yj2);
    const float iw = pb_max(0.0f, ix2 - ix1), ih = pb_max(0.0f, iy2 - iy1);
    const float inter = iw * ih;
    const float area_i = s.area[i], area_j = s.area[j];
    const float uni = area_i + area_j - inter;
    // inter == +0 gives iou == +0 whenever uni > 0 (and 0 otherwise): skip the division then.
    const float iou = (uni > 0 && inter > 0.0f) ? (inter / uni) : 0.0f;
    *iou_out = iou;
    if (iou > thr) return 1;

    const unsigned vis = s.vis[i] & s.vis[j];
    const int cnt = __popc(vis);
    if (cnt < 3) return 0;
    float scale_sq = pb_max(area_i, area_j);
    if (scale_sq < 32.0f * 32.0f) scale_sq = 32.0f * 32.0f;
    const float t8 = 2.0f * scale_sq * 4.0f;
    const float need = (iou > 0.2f) ? pb_min(thr, 0.4f) : thr;

    // Bound 1.  If the keypoint extents of i and j are separated by more than
    // r = sqrt(3.1 * t8 * sigma_max^2) along x or y, every keypoint pair has
    // d2 > 3.003 * t8 * sigma_k^2, i.e. contributes < 0.05: oks < 0.05 <= need - 0.002.
    if (need > 0.06f) {
        const float r2 = 3.1f * t8 * (0.107f * 0.107f);
        const float gx = pb_max(s.ext[0 * Ccap + i] - s.ext[1 * Ccap + j], s.ext[0 * Ccap + j] - s.ext[1 * Ccap + i]);
        const float gy = pb_max(s.ext[2 * Ccap + i] - s.ext[3 * Ccap + j], s.ext[2 * Ccap + j] - s.ext[3 * Ccap + i]);
        if ((gx > 0.0f && gx * gx > r2) || (gy > 0.0f && gy * gy > r2)) return 0;
    }
    // Bound 2.  A keypoint with d2 >= 3.003*den contributes exp(-d2/den) < 0.05, any other at
    // most 1, so oks <= (m + 0.05*(cnt-m))/cnt.  If that is below the smallest threshold that
    // could fire (minus a margin far above fp32 rounding) the reference's test is false.
    int m = 0;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        if (vis & (1u << k)) {
            const float dx = s.kx[k * Ccap + i] - s.kx[k * Ccap + j];
            const float dy = s.ky[k * Ccap + i] - s.ky[k * Ccap + j];
            const float d2 = dx * dx + dy * dy;
            const float sg = kSigmas[k];
            m += (d2 < 3.003f * (t8 * sg * sg)) ? 1 : 0;
        }
    }
    if ((float)m + 0.05f * (float)(cnt - m) < (need - 0.002f) * (float)cnt) return 0;
    return 2;
}

// The undecided pairs are resolved with the exponentials spread over threads: one thread per
// (pair, keypoint) evaluates exp(-d2 / (2*scale*4*sigma^2)) (gpu_postprocess.cu:151-157), then one
// thread per pair adds the 17 terms in keypoint order (the reference's summation order; a
// keypoint that is not visible on both sides contributes an exact +0) and applies :162-167.
// MODE 0: pairs inside the tile -> tile mask; MODE 1: survivor x later rank -> suppressed bitmap.
template <int MODE>
__device__ __forceinline__ void exact_phase(const DnSmem& s, int Ccap, int t0, float thr, int tid) {
    const int nlist = s.misc[3];
    for (int cbase = 0; cbase < nlist; cbase += DN_TERM_PAIRS) {
        const int ncur = (nlist - cbase) < DN_TERM_PAIRS ? (nlist - cbase) : DN_TERM_PAIRS;
        for (int idx = tid; idx < ncur * KP; idx += DN_THREADS) {
            const int e = idx / KP, k = idx - e * KP;
            const unsigned key = s.list_key[cbase + e];
            const int i = (int)(key >> 16), j = (int)(key & 0xffffu);
            float term = 0.0f;
            if ((s.vis[i] & s.vis[j]) & (1u << k)) {
                float scale_sq = pb_max(s.area[i], s.area[j]);
                if (scale_sq < 32.0f * 32.0f) scale_sq = 32.0f * 32.0f;
                const float t8 = 2.0f * scale_sq * 4.0f;
                const float dx = s.kx[k * Ccap + i] - s.kx[k * Ccap + j];
                const float dy = s.ky[k * Ccap + i] - s.ky[k * Ccap + j];
                const float d2 = dx * dx + dy * dy;
                const float sg = s.sig[k];
                term = pb_expf(-d2 / (t8 * sg * sg));
            }
            s.terms[idx] = term;
        }
        __syncthreads();
        for (int e = tid; e < ncur; e += DN_THREADS) {
            const unsigned key = s.list_key[cbase + e];
            const int i = (int)(key >> 16), j = (int)(key & 0xffffu);
            const int cnt = __popc(s.vis[i] & s.vis[j]);
            float sum = 0.0f;
#pragma unroll
            for (int k = 0; k < KP; ++k) sum += s.terms[e * KP + k];
            const float oks = sum / (float)cnt;
            const float iou = s.list_iou[cbase + e];
            if ((oks > thr) || (oks > 0.4f && iou > 0.2f)) {
                if (MODE == 0) atomicOr(&s.tmask[i - t0], 1ull << (j - t0));
                else atomicOr(&s.sup[j >> 5], 1u << (j & 31));
            }
        }
        __syncthreads();
    }
}

// Warp-aggregated append of an undecided pair to the work list.
__device__ __forceinline__ void list_push(const DnSmem& s, bool want, unsigned key, float iou) {
    const unsigned bm = __ballot_sync(FULL, want);
    if (bm == 0u) return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs(bm) - 1) base = atomicAdd(&s.misc[3], __popc(bm));
    base = __shfl_sync(FULL, base, __ffs(bm) - 1);
    if (want) {
        const int pos = base + __popc(bm & ((1u << lane) - 1u));
        s.list_key[pos] = key;
        s.list_iou[pos] = iou;
    }
}

__device__ __forceinline__ bool is_sup(const unsigned* sup, int r) { return (sup[r >> 5] >> (r & 31)) & 1u; }

__global__ void __launch_bounds__(DN_THREADS, 1)
pb_decode_nms_kernel(const float* __restrict__ heads, int N, int Ccap, int Kcap,
                     float conf_thr, float nms_thr, PostBuffers out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DnSmem s;
    dn_carve(smem_raw, Ccap, Kcap, &s);

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* head = heads + (size_t)b * HEAD_ROWS * N;
    const float* conf_row = head + 4 * (size_t)N;
    unsigned long long t_stamp = 0;
    if (tid == 0) t_stamp = globaltimer_ns();
    // stage telemetry: thread 0 accumulates globaltimer deltas in shared memory (cheap) and
    // flushes them once at the end of the kernel
    auto stamp = [&](int slot) {
        if (tid == 0) { const unsigned long long now = globaltimer_ns(); s.acc[slot] += now - t_stamp; t_stamp = now; }
    };

    // ---------------- 1. confidence scan + ordered compaction (A1, R1) ----------------
    const bool vec_ok = ((N & 3) == 0) && ((((uintptr_t)conf_row) & 15) == 0);
    const int ngroups = (N + 3) >> 2;
    const int chunks = (ngroups + DN_THREADS - 1) / DN_THREADS;
    unsigned m[4] = {0u, 0u, 0u, 0u};
    constexpr int SU = 8;                      // independent 128-bit loads in flight per thread
    for (int c0 = 0; c0 < chunks; c0 += SU) {
        float4 v[SU];
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            const int g = (c0 + u) * DN_THREADS + tid;
            const float ninf = -__int_as_float(0x7f800000);
            v[u] = make_float4(ninf, ninf, ninf, ninf);
            if (c0 + u < chunks && g < ngroups) {
                if (vec_ok) {
                    v[u] = ldg_stream_f4(reinterpret_cast<const float4*>(conf_row) + g);
                } else {
                    const int a = g * 4;
                    v[u].x = conf_row[a];
                    if (a + 1 < N) v[u].y = conf_row[a + 1];
                    if (a + 2 < N) v[u].z = conf_row[a + 2];
                    if (a + 3 < N) v[u].w = conf_row[a + 3];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < SU; ++u) {
            const int c = c0 + u;
            if (c < chunks) {
                // keep iff !(conf < thr)  (gpu_postprocess.cu:51)
                const unsigned hm = (!(v[u].x < conf_thr) ? 1u : 0u) | (!(v[u].y < conf_thr) ? 2u : 0u) |
                                    (!(v[u].z < conf_thr) ? 4u : 0u) | (!(v[u].w < conf_thr) ? 8u : 0u);
                m[c >> 3] |= hm << ((c & 7) * 4);
                const int wsum = __reduce_add_sync(FULL, __popc(hm));
                if (lane == 0) s.cnt[c * DN_WARPS + warp] = wsum;
            }
        }
    }
    if (tid < 8) s.misc[tid] = 0;
    if (tid < 16) s.acc[tid] = 0ull;
    if (tid < KP) s.sig[tid] = kSigmas[tid];
    for (int i = tid; i < (Ccap + 31) / 32 + 2; i += DN_THREADS) s.sup[i] = 0u;
    for (int i = tid; i < Ccap; i += DN_THREADS) s.vis[i] = 0u;
    __syncthreads();
    if (warp == 0) {   // exclusive scan of the (chunk, warp) counts in anchor order
        const int n = chunks * DN_WARPS;
        const int per = (n + 31) / 32;
        int local = 0;
        for (int i = 0; i < per; ++i) { const int e = lane * per + i; if (e < n) local += s.cnt[e]; }
        int incl = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
        int run = incl - local;
        for (int i = 0; i < per; ++i) {
            const int e = lane * per + i;
            if (e < n) { const int t = s.cnt[e]; s.cnt[e] = run; run += t; }
        }
        if (lane == 31) s.misc[0] = incl;
    }
    __syncthreads();
    for (int c = 0; c < chunks; ++c) {
        const unsigned hm = (m[c >> 3] >> ((c & 7) * 4)) & 15u;
        if (__ballot_sync(FULL, hm != 0u) == 0u) continue;
        const int n = __popc(hm);
        int incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
        int pos = s.cnt[c * DN_WARPS + warp] + incl - n;
        const int a0 = (c * DN_THREADS + tid) * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (hm & (1u << e)) {
                if (pos < Ccap) { s.anchor[pos] = a0 + e; s.score[pos] = conf_row[a0 + e]; }
                ++pos;
            }
        }
    }
    __syncthreads();
    const int total = s.misc[0];
    const int C = total < Ccap ? total : Ccap;
    stamp(0);

    // ---------------- 2. rank by (score desc, slot asc)  (A3 sort, R2) ----------------
    for (int i = tid; i < C; i += DN_THREADS) {
        const float si = s.score[i];
        int rank = 0;
        for (int j = 0; j < C; ++j) {
            const float sj = s.score[j];
            rank += (sj > si || (sj == si && j < i)) ? 1 : 0;
        }
        s.order[rank] = i;
    }
    __syncthreads();
    stamp(1);

    // ---------------- 3. gather the other head rows at the candidate anchors ----------
    {
        const int items = 55 * C;
        constexpr int GU = 8;                  // independent sector loads in flight per thread
        for (int it0 = tid; it0 < items; it0 += DN_THREADS * GU) {
            float v[GU];
            int rr[GU], rw[GU];
#pragma unroll
            for (int u = 0; u < GU; ++u) {
                const int it = it0 + u * DN_THREADS;
                rr[u] = -1;
                if (it < items) {
                    const int ri = it / C;
                    const int r = it - ri * C;
                    const int row = ri < 4 ? ri : ri + 1;
                    const int a = s.anchor[s.order[r]];
                    v[u] = ldg_stream_f(head + (size_t)row * N + a);
                    rr[u] = r; rw[u] = row;
                }
            }
#pragma unroll
            for (int u = 0; u < GU; ++u) {
                if (rr[u] < 0) continue;
                const int r = rr[u], row = rw[u];
                if (row < 4) {
                    s.box[row * Ccap + r] = v[u];
                } else {
                    const int k = (row - 5) / 3, comp = (row - 5) - 3 * k;
                    if (comp == 0) s.kx[k * Ccap + r] = v[u];
                    else if (comp == 1) s.ky[k * Ccap + r] = v[u];
                    else if (v[u] > 0.2f) atomicOr(&s.vis[r], 1u << k);
                }
            }
        }
    }
    __syncthreads();
    for (int r = tid; r < C; r += DN_THREADS) {   // cx,cy,w,h -> corners (:66-69) + area (:128)
        const float cx = s.box[0 * Ccap + r], cy = s.box[1 * Ccap + r], w = s.box[2 * Ccap + r], h = s.box[3 * Ccap + r];
        const float x1 = cx - w * 0.5f, y1 = cy - h * 0.5f, x2 = cx + w * 0.5f, y2 = cy + h * 0.5f;
        s.box[0 * Ccap + r] = x1; s.box[1 * Ccap + r] = y1; s.box[2 * Ccap + r] = x2; s.box[3 * Ccap + r] = y2;
        s.area[r] = (x2 - x1) * (y2 - y1);
        float lx = s.kx[r], hx = lx, ly = s.ky[r], hy = ly;
#pragma unroll
        for (int k = 1; k < KP; ++k) {
            const float x = s.kx[k * Ccap + r], y = s.ky[k * Ccap + r];
            lx = fminf(lx, x); hx = fmaxf(hx, x); ly = fminf(ly, y); hy = fmaxf(hy, y);
        }
        s.ext[0 * Ccap + r] = lx; s.ext[1 * Ccap + r] = hx; s.ext[2 * Ccap + r] = ly; s.ext[3 * Ccap + r] = hy;
    }

    stamp(2);
    // ---------------- 4. greedy suppression in rank order (A2 + A3 sweep) --------------
    int nkeep = 0;
    for (int t0 = 0; t0 < C; t0 += 64) {
        const int tl = (C - t0) < 64 ? (C - t0) : 64;
        if (tid < 64) s.tmask[tid] = 0ull;
        if (tid == 0) s.misc[3] = 0;
        __syncthreads();
        // (a) pairs inside the tile: decide cheaply, queue the undecided ones
        for (int p = tid; p < 64 * 64; p += DN_THREADS) {
            const int a = p >> 6, bb = p & 63;
            int q = 0;
            float iou = 0.0f;
            if (a < bb && bb < tl && !is_sup(s.sup, t0 + a) && !is_sup(s.sup, t0 + bb))
                q = nms_quick(s, Ccap, t0 + a, t0 + bb, nms_thr, &iou);
            if (q == 1) atomicOr(&s.tmask[a], 1ull << bb);
            list_push(s, q == 2, ((unsigned)(t0 + a) << 16) | (unsigned)(t0 + bb), iou);
        }
        __syncthreads();
        stamp(8);
        exact_phase<0>(s, Ccap, t0, nms_thr, tid);
        stamp(9);
        if (warp == 0) {
            // serial greedy over the tile with the 64 mask rows held in registers (2 per lane)
            const unsigned long long row_lo = s.tmask[lane], row_hi = s.tmask[lane + 32];
            unsigned long long supt = (unsigned long long)s.sup[t0 >> 5] | ((unsigned long long)s.sup[(t0 >> 5) + 1] << 32);
            unsigned long long kept = 0ull;
            int nk = s.misc[1];
            const int nk0 = nk;
            for (int a = 0; a < tl && nk < Kcap; ++a) {          // :224 "num_keep < 256"
                const unsigned long long ra = __shfl_sync(FULL, a < 32 ? row_lo : row_hi, a & 31);
                if ((supt >> a) & 1ull) continue;
                kept |= 1ull << a;
                ++nk;
                supt |= ra;
            }
            for (int half = 0; half < 2; ++half) {               // lanes write the kept ranks in order
                const int a = half * 32 + lane;
                if ((kept >> a) & 1ull) {
                    const int pos = __popcll(kept & ((1ull << a) - 1ull));
                    s.keep[nk0 + pos] = t0 + a;
                    s.tk[pos] = t0 + a;
                }
            }
            if (lane == 0) {
                s.sup[t0 >> 5] = (unsigned)supt;
                s.sup[(t0 >> 5) + 1] = (unsigned)(supt >> 32);
                s.misc[1] = nk; s.misc[2] = nk - nk0; s.misc[3] = 0;
            }
        }
        __syncthreads();
        stamp(10);
        nkeep = s.misc[1];
        const int ntk = s.misc[2];
        if (nkeep >= Kcap) break;
        // (c) the tile's survivors strike the remaining ranks, DN_LIST candidate pairs per round
        const int j0 = t0 + 64, rem = C - j0;
        if (rem > 0 && ntk > 0) {
            const int pairs = ntk * rem;
            for (int pbase = 0; pbase < pairs; pbase += DN_LIST) {
                const int pend = (pairs - pbase) < DN_LIST ? (pairs - pbase) : DN_LIST;
                for (int p0 = 0; p0 < pend; p0 += DN_THREADS) {   // uniform trip count: list_push uses ballots
                    const int p = pbase + p0 + tid;
                    int q = 0, ai = 0, j = 0;
                    float iou = 0.0f;
                    if (p0 + tid < pend) {
                        ai = p / rem;
                        j = j0 + (p - ai * rem);
                        if (!is_sup(s.sup, j)) q = nms_quick(s, Ccap, s.tk[ai], j, nms_thr, &iou);
                    }
                    if (q == 1) atomicOr(&s.sup[j >> 5], 1u << (j & 31));
                    list_push(s, q == 2, ((unsigned)s.tk[ai] << 16) | (unsigned)j, iou);
                }
                __syncthreads();
                exact_phase<1>(s, Ccap, t0, nms_thr, tid);
                if (tid == 0) s.misc[3] = 0;
                __syncthreads();
            }
        }
        stamp(11);
    }
    __syncthreads();
    nkeep = s.misc[1];
    stamp(3);

    // ---------------- 5. kept detections in score order -------------------------------
    float* o_pose = out.det_poses + (size_t)b * Kcap * POSE_F;
    float* o_box = out.det_bboxes + (size_t)b * Kcap * 4;
    float* o_score = out.det_scores + (size_t)b * Kcap;
    int* o_slot = out.keep_slots + (size_t)b * Kcap;
    int* o_anchor = out.keep_anchors + (size_t)b * Kcap;
    for (int it = tid; it < nkeep * POSE_F; it += DN_THREADS) {
        const int k = it / POSE_F, e = it - k * POSE_F;
        const int r = s.keep[k];
        const int kp = e / 3, comp = e - 3 * kp;
        float v;
        if (comp == 0) v = s.kx[kp * Ccap + r];
        else if (comp == 1) v = s.ky[kp * Ccap + r];
        else v = head[(size_t)(7 + 3 * kp) * N + s.anchor[s.order[r]]];
        o_pose[it] = v;
    }
    for (int it = tid; it < nkeep * 4; it += DN_THREADS) {
        const int k = it >> 2, e = it & 3;
        o_box[it] = s.box[e * Ccap + s.keep[k]];
    }
    for (int k = tid; k < nkeep; k += DN_THREADS) {
        const int slot = s.order[s.keep[k]];
        o_score[k] = s.score[slot];
        o_slot[k] = slot;
        o_anchor[k] = s.anchor[slot];
    }
    stamp(4);
    if (tid == 0) { out.num_keep[b] = nkeep; out.num_cand[b] = C; s.acc[7] = 1ull; }
    if (tid < 16 && s.acc[tid] != 0ull) out.stage_ns[(size_t)b * 16 + tid] += s.acc[tid];
}

cudaError_t launch_decode_nms(const float* d_heads, int B, int N, int max_cand, int max_keep,
                              float conf_thr, float nms_thr, const PostBuffers& out,
                              cudaStream_t stream) {
    const size_t smem = decode_nms_smem_bytes(max_cand, max_keep);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(pb_decode_nms_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    pb_decode_nms_kernel<<<B, DN_THREADS, smem, stream>>>(d_heads, N, max_cand, max_keep, conf_thr,
                                                          nms_thr, out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace pb
