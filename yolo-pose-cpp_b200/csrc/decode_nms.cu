// decode_nms.cu — head decode + confidence filter + OKS/IoU pose-NMS for B streams.
//
// Replaces GPUPostprocess::process (reference src/cuda/gpu_postprocess.cu:366-476): six
// kernels (:30-313), two of them single-threaded, and three host synchronisations per frame
// and stream, by TWO launches for all B streams:
//
// K1  pb_decode_gather_kernel   grid (NSEG, B): every CTA owns one segment of one stream's
//     confidence row.  The segment is streamed with 128-bit loads (the only dense read of the
//     head tensor), hits are compacted in ascending anchor order with warp popcounts + one
//     block scan (rule R1 replaces the reference's atomicAdd arrival order, :54), and the
//     other 55 head rows are gathered only at the hit anchors (32-byte sectors) and written
//     as one contiguous 56-float record per candidate to an L2-resident scratch.  Splitting a
//     stream over NSEG CTAs puts the sector gathers of 64 streams on all 148 SMs' miss queues.
//
// K2  pb_nms_kernel             grid B, one CTA per stream: candidates are ranked by
//     (score desc, slot asc) == the reference's stable insertion sort (:178-203, rule R2), their
//     records are loaded (coalesced) into shared memory in rank order as SoA, and greedy
//     suppression runs tile-by-tile over 64 ranks: pairs inside the tile -> bitmask, an
//     ffs-driven serial resolve of the tile (one step per SURVIVOR), then the survivors strike
//     the later ranks in parallel.  Only mask bits the reference's sweep (:209-242) would read
//     are evaluated.  The pair test is the reference's (:88-172), staged so that lanes stay
//     dense: (A) IoU + an exact geometric bound, (B) an exact per-keypoint bound, (C) the 17
//     exponentials, one thread per (pair, keypoint), summed in keypoint order.  Kept
//     detections leave in score order (== kernelCompactDetections/CopyBack, :248-313) with
//     their candidate slots and anchor ids.
//
// Arithmetic goes through pb_math.h (see there), compiled with --fmad=false.
#include <cstdlib>
#include "nms_body.cuh"

namespace pb {

// =======================================================================================
// K1: decode + gather
// =======================================================================================
constexpr int DG_THREADS = 256;
constexpr int DG_WARPS = DG_THREADS / 32;
constexpr int DG_MAX_CHUNKS = 16;            // float4 groups per segment <= 16 * 256
constexpr int BOX_ROWS = 5;                  // cx, cy, w, h, confidence

template <int ROWS>
__global__ void __launch_bounds__(DG_THREADS)
pb_decode_gather_kernel(const float* __restrict__ heads, int N, int nseg, int groups_per_seg, int segcap,
                        float conf_thr, CandScratch cs) {
    constexpr int rows = ROWS;
    __shared__ int s_cnt[DG_MAX_CHUNKS * DG_WARPS];
    __shared__ int s_total;
    extern __shared__ int s_anchor[];            // [segcap]

    const int seg = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* head = heads + (size_t)b * HEAD_ROWS * N;
    const float* conf_row = head + 4 * (size_t)N;
    const bool vec_ok = ((N & 3) == 0) && ((((uintptr_t)conf_row) & 15) == 0);
    const int ngroups = (N + 3) >> 2;
    const int g0 = seg * groups_per_seg;
    int g1 = g0 + groups_per_seg; if (g1 > ngroups) g1 = ngroups;
    const int nloc = g1 > g0 ? g1 - g0 : 0;
    const int chunks = (nloc + DG_THREADS - 1) / DG_THREADS;

    // ---- confidence scan: keep iff !(conf < thr)  (gpu_postprocess.cu:51) ----
    unsigned hm[2] = {0u, 0u};                   // 4 bits per chunk
    const float ninf = -__int_as_float(0x7f800000);
    for (int c0 = 0; c0 < chunks; c0 += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int g = g0 + (c0 + u) * DG_THREADS + tid;
            v[u] = make_float4(ninf, ninf, ninf, ninf);
            if (c0 + u < chunks && g < g1) {
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
        for (int u = 0; u < 8; ++u) {
            const int c = c0 + u;
            if (c < chunks) {
                const int g = g0 + c * DG_THREADS + tid;
                unsigned m = 0u;
                if (g < g1) {
                    const int a = g * 4;
                    m = (!(v[u].x < conf_thr) ? 1u : 0u) | ((a + 1 < N && !(v[u].y < conf_thr)) ? 2u : 0u) |
                        ((a + 2 < N && !(v[u].z < conf_thr)) ? 4u : 0u) | ((a + 3 < N && !(v[u].w < conf_thr)) ? 8u : 0u);
                }
                hm[c >> 3] |= m << ((c & 7) * 4);
                const int wsum = __reduce_add_sync(FULL, __popc(m));
                if (lane == 0) s_cnt[c * DG_WARPS + warp] = wsum;
            }
        }
    }
    __syncthreads();
    if (warp == 0) {                             // exclusive scan of (chunk, warp) counts = anchor order
        const int n = chunks * DG_WARPS;
        const int per = (n + 31) / 32;
        int local = 0;
        for (int i = 0; i < per; ++i) { const int e = lane * per + i; if (e < n) local += s_cnt[e]; }
        int incl = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
        int run = incl - local;
        for (int i = 0; i < per; ++i) {
            const int e = lane * per + i;
            if (e < n) { const int t = s_cnt[e]; s_cnt[e] = run; run += t; }
        }
        if (lane == 31) s_total = incl;
    }
    __syncthreads();
    for (int c = 0; c < chunks; ++c) {
        const unsigned m = (hm[c >> 3] >> ((c & 7) * 4)) & 15u;
        if (__ballot_sync(FULL, m != 0u) == 0u) continue;
        const int n = __popc(m);
        int incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += t; }
        int pos = s_cnt[c * DG_WARPS + warp] + incl - n;
        const int a0 = (g0 + c * DG_THREADS + tid) * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (m & (1u << e)) { if (pos < segcap) s_anchor[pos] = a0 + e; ++pos; }
        }
    }
    __syncthreads();
    const int total = s_total;
    const int n_s = total < segcap ? total : segcap;   // only the first Ccap overall can matter (R1)

    // ---- gather head rows [0, rows) at the hit anchors into one record per candidate.  rows = 56: the
    // complete column (head in HBM: sector reads are cheap, NMS then never touches the head again);
    // rows = 5: box + confidence only (head in page-locked host memory: the 51 keypoint rows are
    // fetched later, by the NMS kernel, only for the candidates that can survive) ----
    float* rec = cs.records + ((size_t)(b * nseg + seg) * segcap) * HEAD_ROWS;
    int* anc = cs.anchors + (size_t)(b * nseg + seg) * segcap;
    const int items = n_s * rows;
    for (int it0 = tid; it0 < items; it0 += DG_THREADS * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int it = it0 + u * DG_THREADS;
            if (it < items) {
                const int h = it / rows, row = it - h * rows;
                v[u] = ldg_stream_f(head + (size_t)row * N + s_anchor[h]);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int it = it0 + u * DG_THREADS;
            if (it < items) { const int h = it / rows, row = it - h * rows; rec[h * HEAD_ROWS + row] = v[u]; }
        }
    }
    for (int h = tid; h < n_s; h += DG_THREADS) anc[h] = s_anchor[h];
    if (tid == 0) cs.counts[b * nseg + seg] = n_s;
}

// =======================================================================================
// K2: rank + NMS (nms_body.cuh), one CTA per stream
// =======================================================================================
constexpr int NM_THREADS = 1024;

size_t decode_nms_smem_bytes(int max_cand, int max_keep) { return nm_carve(nullptr, max_cand, max_keep, nullptr); }

#ifdef PB_NMS_MAXNREG
__global__ void __maxnreg__(PB_NMS_MAXNREG)
#else
__global__ void __launch_bounds__(NM_THREADS, 1)
#endif
pb_nms_kernel(const float* __restrict__ heads, int N, int lazy, CandScratch cs, int nseg, int segcap, int Ccap, int Kcap, float nms_thr,
              PostBuffers out, SmemOffsets so) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NmSmem s;
    nm_from_offsets(smem_raw, smem_raw, so, s);
    nms_body<NM_THREADS>(s, heads, N, lazy, cs, nseg, segcap, Ccap, Kcap, nms_thr, out, blockIdx.x, gridDim.x, nullptr, nullptr, 0);
    if (out.ready) {
        // resident tracker (tracker.cu: pb_tracker_seq_kernel): this stream-frame's kept detections are complete
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(out.ready + blockIdx.x), "r"(out.ready_seq) : "memory");
        }
    }
}

// Tiered variants: CTAs that share an SM — half-SM (512 threads, at most ~113 KB of shared memory: two of them, or one and a
// compact resident tracker CTA, tracker.cu: pb_tracker_seq_kernel), third-SM (384 or 256 threads, ~75 KB) and quarter-SM
// (256 threads, ~56 KB).  The NMS stage is not on a video stream's chain of dependent frames, so what counts for it is SM-time
// per stream-frame, not latency: a 1024-thread CTA keeps an SM's issue slots 39 % busy for 33 us at 136 candidates, and most
// of its warps only walk through barriers; smaller CTAs packed three or four to an SM fill the slots.  The shared memory holds
// the working set of CT candidates; a stream with more (up to Ccap, rule R1) keeps its per-candidate arrays in a global scratch
// instead (spill path: the same code on generic pointers, same results), exactly as the fused per-stream kernel does (fused.cu).
struct NmsTierParams {
    const float* heads;
    int N, lazy;
    CandScratch cs;
    int nseg, segcap, Ccap, CT, Kcap;
    float nms_thr;
    PostBuffers out;
    SmemOffsets so_small, so_big;
    unsigned char* spill;       // [B, spill_stride] (nullptr when CT == Ccap)
    size_t spill_stride;
};

// (cold path, a function of its own: kept out of the register allocation of the common path; the parameters stay in constant memory)
template <int NT>
static __device__ __noinline__ void nms_tier_spill(const NmsTierParams& F, int b) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NmSmem s;
    nm_from_offsets(smem_raw, F.spill + (size_t)b * F.spill_stride, F.so_big, s);
    nms_body<NT>(s, F.heads, F.N, F.lazy, F.cs, F.nseg, F.segcap, F.Ccap, F.Kcap, F.nms_thr, F.out, b, (int)gridDim.x, nullptr, nullptr, 0);
}

template <int NT, int PER_SM>
__global__ void __launch_bounds__(NT, PER_SM)
pb_nms_tier_kernel(const __grid_constant__ NmsTierParams F) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    int total = 0;
    for (int sg = 0; sg < F.nseg; ++sg) total += F.cs.counts[b * F.nseg + sg];
    if (total <= F.CT || F.spill == nullptr) {
        NmSmem s;
        nm_from_offsets(smem_raw, smem_raw, F.so_small, s);
        nms_body<NT>(s, F.heads, F.N, F.lazy, F.cs, F.nseg, F.segcap, F.CT, F.Kcap, F.nms_thr, F.out, b, gridDim.x, nullptr, nullptr, 0);
    } else {
        nms_tier_spill<NT>(F, b);
    }
    if (F.out.ready) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(F.out.ready + b), "r"(F.out.ready_seq) : "memory");
        }
    }
}

static const void* nms_tier_func(int threads, int per_sm) {
    if (threads == 1024) return (const void*)pb_nms_tier_kernel<1024, 1>;
    if (threads == 512) return (const void*)pb_nms_tier_kernel<512, 2>;
    if (threads == 384) return (const void*)pb_nms_tier_kernel<384, 3>;
    if (threads == 256 && per_sm == 3) return (const void*)pb_nms_tier_kernel<256, 3>;
    if (threads == 256 && per_sm == 4) return (const void*)pb_nms_tier_kernel<256, 4>;
    return nullptr;
}

// per_sm CTAs of `threads` threads per SM: 2 x 512 (the default), 3 x 384, 3 x 256 or 4 x 256
// target_kb > 0: shared memory of a CTA (1024 threads: one CTA per SM, sized below the full candidate cap so that the SM keeps
// the shared-memory / L1 split of the kernels around it)
NmsTierPlan nms_tier_plan(int max_cand, int max_keep, size_t smem_optin, int threads, int per_sm, int target_kb) {
    NmsTierPlan p{};
    p.ok = false;
    p.threads = threads; p.per_sm = per_sm;
    if (!nms_tier_func(threads, per_sm)) return p;
    size_t target = (size_t)(228 - per_sm) * 1024 / (size_t)per_sm;      // 228 KB per SM, 1 KB reserved per resident CTA
    target &= ~(size_t)1023;
    if (target_kb > 0 && (size_t)target_kb * 1024 < target) target = (size_t)target_kb * 1024;
    if (target > smem_optin) target = smem_optin;
    int ct = 0;
    for (int c = 64; c <= max_cand; c += 16) {
        if (nm_carve(nullptr, c, max_keep, nullptr) <= target) ct = c; else break;
    }
    if (nm_carve(nullptr, max_cand, max_keep, nullptr) <= target) ct = max_cand;
    if (ct < 64) return p;
    p.CT = ct;
    p.so_small = nm_offsets(ct, max_keep);
    p.smem_bytes = nm_carve(nullptr, ct, max_keep, nullptr);
    size_t fixed_bytes = 0, cand_bytes = 0;
    p.so_big = nm_offsets(max_cand, max_keep, true, &fixed_bytes, &cand_bytes);
    p.spill_stride = ct < max_cand ? ((cand_bytes + 255) & ~(size_t)255) : 0;
    if (fixed_bytes > p.smem_bytes) return p;
    p.ok = true;
    return p;
}

cudaError_t launch_nms_tier(const NmsTierPlan& tp, const float* d_heads, int N, int sweep, int B, int max_cand, int max_keep, float nms_thr,
                            const DecodePlan& plan, const CandScratch& cs, const PostBuffers& out, unsigned char* spill, cudaStream_t stream) {
    NmsTierParams F{};
    F.heads = d_heads; F.N = N; F.lazy = sweep; F.cs = cs; F.nseg = plan.nseg; F.segcap = plan.segcap;
    F.Ccap = max_cand; F.CT = tp.CT; F.Kcap = max_keep; F.nms_thr = nms_thr; F.out = out;
    F.so_small = tp.so_small; F.so_big = tp.so_big; F.spill = tp.spill_stride ? spill : nullptr; F.spill_stride = tp.spill_stride;
    if (tp.spill_stride && !spill) return cudaErrorInvalidValue;
    const void* fn = nms_tier_func(tp.threads, tp.per_sm);
    if (!fn) return cudaErrorInvalidValue;
    const cudaError_t e = ensure_dyn_smem(fn, tp.smem_bytes);
    if (e != cudaSuccess) return e;
    if (tp.threads == 1024) pb_nms_tier_kernel<1024, 1><<<B, 1024, tp.smem_bytes, stream>>>(F);
    else if (tp.threads == 512) pb_nms_tier_kernel<512, 2><<<B, 512, tp.smem_bytes, stream>>>(F);
    else if (tp.threads == 384) pb_nms_tier_kernel<384, 3><<<B, 384, tp.smem_bytes, stream>>>(F);
    else if (tp.per_sm == 3) pb_nms_tier_kernel<256, 3><<<B, 256, tp.smem_bytes, stream>>>(F);
    else pb_nms_tier_kernel<256, 4><<<B, 256, tp.smem_bytes, stream>>>(F);
    count_launch();
    return cudaGetLastError();
}

// The resident tracker kernel waits on the device for these kernels.  With CUDA's lazy module loading the FIRST launch of a
// kernel loads its code, which may have to wait for running kernels: a kernel that spins until that launch has run would
// wait for its time-out.  The host layer therefore loads (and sizes) them before it enqueues a resident tracker launch.
cudaError_t preload_post_kernels(int max_cand, int max_keep, const NmsTierPlan* tier) {
    cudaFuncAttributes a{};
    cudaError_t e = cudaFuncGetAttributes(&a, (const void*)pb_decode_gather_kernel<HEAD_ROWS>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, (const void*)pb_decode_gather_kernel<BOX_ROWS>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, (const void*)pb_nms_kernel);
    if (e == cudaSuccess) e = ensure_dyn_smem((const void*)pb_nms_kernel, decode_nms_smem_bytes(max_cand, max_keep));
    if (e == cudaSuccess && tier && tier->ok) {
        const void* fn = nms_tier_func(tier->threads, tier->per_sm);
        e = fn ? cudaFuncGetAttributes(&a, fn) : cudaErrorInvalidValue;
        if (e == cudaSuccess) e = ensure_dyn_smem(fn, tier->smem_bytes);
    }
    return e;
}

// =======================================================================================
// launch
// =======================================================================================
DecodePlan decode_plan(int B, int N, int max_cand) {
    DecodePlan p{};
    const int ngroups = (N + 3) / 4;
    int nseg = (2 * 148 + B - 1) / B;                      // >= 2 CTAs per SM in total
    if (nseg < 1) nseg = 1;
    if (nseg > 16) nseg = 16;
    if (const char* e = getenv("PB_DECODE_NSEG")) { const int v = atoi(e); if (v >= 1 && v <= 16) nseg = v; }   // experiment
    const int min_seg = (ngroups + DG_MAX_CHUNKS * DG_THREADS - 1) / (DG_MAX_CHUNKS * DG_THREADS);
    if (nseg < min_seg) nseg = min_seg;
    p.nseg = nseg;
    p.groups_per_seg = (ngroups + nseg - 1) / nseg;
    p.segcap = max_cand;
    return p;
}

cudaError_t launch_decode_gather(const float* d_heads, int B, int N, float conf_thr, bool lazy_keypoints, const DecodePlan& plan,
                                 const CandScratch& cs, cudaStream_t stream) {
    const size_t smem1 = (size_t)plan.segcap * sizeof(int);
    if (lazy_keypoints)
        pb_decode_gather_kernel<BOX_ROWS><<<dim3(plan.nseg, B), DG_THREADS, smem1, stream>>>(d_heads, N, plan.nseg, plan.groups_per_seg,
                                                                                            plan.segcap, conf_thr, cs);
    else
        pb_decode_gather_kernel<HEAD_ROWS><<<dim3(plan.nseg, B), DG_THREADS, smem1, stream>>>(d_heads, N, plan.nseg, plan.groups_per_seg,
                                                                                             plan.segcap, conf_thr, cs);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_nms(const float* d_heads, int N, int sweep, int B, int max_cand, int max_keep, float nms_thr,
                       const DecodePlan& plan, const CandScratch& cs, const PostBuffers& out, cudaStream_t stream) {
    const size_t smem2 = decode_nms_smem_bytes(max_cand, max_keep);
    const SmemOffsets so = nm_offsets(max_cand, max_keep);
    {
        const cudaError_t e = ensure_dyn_smem((const void*)pb_nms_kernel, smem2);
        if (e != cudaSuccess) return e;
    }
    pb_nms_kernel<<<B, NM_THREADS, smem2, stream>>>(d_heads, N, sweep, cs, plan.nseg, plan.segcap, max_cand, max_keep, nms_thr, out, so);
    count_launch();
    return cudaGetLastError();
}

}  // namespace pb
