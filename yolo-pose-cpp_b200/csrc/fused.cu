// fused.cu — the per-stream pipeline: rank + pose-NMS and the tracker update of one stream-frame in ONE CTA.
//
// Replaces the hand-over GPUPostprocess::process -> GPUTracker::update of the reference's frame loop
// (src/main.cpp:207-224; gpu_postprocess.cu:366-476 -> gpu_tracker.cu:1057-1158): the kept detections go from the
// NMS stage to the tracker stage through shared memory, the two stages share one launch and one residency period on
// the SM, and the shared memory is sized by the candidates a stream actually has, so that two stream-CTAs fit an SM:
//
//   * layout  [ detection prefix | union( NMS working set for CT candidates , tracker working set ) ]
//     The prefix (this frame's detection scores and poses, tracker_body.cuh: tk_carve) is written by the NMS stage's
//     output step and read by the tracker stage; everything behind it is the NMS stage's first and the tracker's after.
//   * tier    CT = the largest candidate count whose NMS working set fits beside the tracker's (plan_fused).  A stream
//     with more candidates than CT (up to max_candidates, rule R1) takes the spill path: the per-candidate arrays of
//     the NMS stage live in a per-stream global scratch (L2) instead, the same code on generic pointers — slower, same
//     results, no second launch and no effect on the other streams of the batch.
//   * order   the NMS stage depends on this frame's candidates only and runs at once; the tracker stage waits (acquire
//     spin on seq_done[b], tracker_body.cuh) for the same video stream's previous frame, which may still be in flight
//     in another CTA on another lane.  The wait of step i+1 is hidden behind its own NMS stage.
#include <cstdlib>
#include "nms_body.cuh"
#include "tracker_body.cuh"

namespace pb {

struct FusedParams {
    const float* heads;
    int N, sweep;
    CandScratch cs;
    int nseg, segcap;
    int Ccap, CT, Kcap;                 // candidate cap (max_candidates), shared-memory tier, keep cap
    float nms_thr;
    PostBuffers out;
    SmemOffsets so_small, so_big;       // NMS layouts: all in shared memory (CT) / per-candidate arrays in the spill scratch (Ccap)
    unsigned nms_base;                  // byte offset of the NMS layouts inside the CTA's shared memory (= detection prefix)
    unsigned char* spill;               // [B, spill_stride] or nullptr when CT == Ccap
    size_t spill_stride;
};

template <int NT>
static __device__ __noinline__ int nms_spill(unsigned char* fixed, const FusedParams& F, int b, int nstreams, float* fdet, float* fscore, int fcap) {
    NmSmem s;
    nm_from_offsets(fixed, F.spill + (size_t)b * F.spill_stride, F.so_big, s);
    return nms_body<NT>(s, F.heads, F.N, F.sweep, F.cs, F.nseg, F.segcap, F.Ccap, F.Kcap, F.nms_thr, F.out, b, nstreams, fdet, fscore, fcap);
}

template <int NT>
__global__ void __launch_bounds__(NT, NT <= 512 ? 2 : 1)
pb_stream_kernel(const FusedParams F, const TrackBuffers tb, const TrackParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    // candidates of this stream-frame (sum of the decode kernel's segment counts): selects the layout
    int total = 0;
    for (int sg = 0; sg < F.nseg; ++sg) total += F.cs.counts[b * F.nseg + sg];
    if (total > F.Ccap) total = F.Ccap;
    float* fscore = reinterpret_cast<float*>(smem_raw + P.so.off[16]);     // TkSmem::dscore
    float* fdet = reinterpret_cast<float*>(smem_raw + P.so.off[34]);       // TkSmem::det
    unsigned char* fixed = smem_raw + F.nms_base;
    int nkeep;
    if (total <= F.CT) {
        NmSmem s;
        nm_from_offsets(fixed, fixed, F.so_small, s);
        nkeep = nms_body<NT>(s, F.heads, F.N, F.sweep, F.cs, F.nseg, F.segcap, F.CT, F.Kcap, F.nms_thr, F.out, b, (int)gridDim.x, fdet, fscore, P.Dm);
    } else {
        nkeep = nms_spill<NT>(fixed, F, b, (int)gridDim.x, fdet, fscore, P.Dm);
    }
    __syncthreads();                    // the NMS stage's shared memory is the tracker's from here on
    DetSource none{nullptr, nullptr, nullptr, 0};
    tracker_body<NT, true, true>(tb, P, none, b, smem_raw, nkeep);
}

// Shared-memory plan of the fused kernel for one handle; ok == false: this configuration keeps the three-kernel step.
FusedPlan plan_fused(int T, int Dm, int max_cand, int max_keep, size_t smem_optin) {
    FusedPlan f{};
    f.ok = false;
    f.tk = tracker_plan(T, Dm, true);
    if (!(f.tk.cost_in_smem && f.tk.det_in_smem && f.tk.pred_in_smem)) return f;       // large tables: tracker_body<.., false, ..>
    if ((long)T * Dm > 16384) return f;                                                // CTA-wide auction: wants its own thread count
    f.threads = 512;
    if (const char* e = getenv("PB_FUSED_THREADS")) { const int t = atoi(e); if (t == 256 || t == 512 || t == 1024) f.threads = t; }
    const size_t two_per_sm = 113 * 1024;                                              // (228 KB - 2 x 1 KB reserved) / 2
    size_t target = f.tk.smem_bytes > two_per_sm ? f.tk.smem_bytes : two_per_sm;
    if (f.threads == 1024) target = smem_optin;                                        // one CTA per SM anyway
    if (target > smem_optin) target = smem_optin;
    const size_t prefix = f.tk.prefix_bytes;
    int ct = 0;
    for (int c = 64; c <= max_cand; c += 32) {
        if (prefix + nm_carve(nullptr, c, max_keep, nullptr) <= target) ct = c; else break;
    }
    if (prefix + nm_carve(nullptr, max_cand, max_keep, nullptr) <= target) ct = max_cand;
    if (const char* e = getenv("PB_FUSED_TIER")) { const int t = atoi(e); if (t >= 32 && t <= ct) ct = t; }
    if (ct < 64) return f;
    f.CT = ct;
    size_t small_bytes = 0, fixed_bytes = 0, cand_bytes = 0;
    f.so_small = nm_offsets(ct, max_keep);
    small_bytes = nm_carve(nullptr, ct, max_keep, nullptr);
    f.so_big = nm_offsets(max_cand, max_keep, true, &fixed_bytes, &cand_bytes);
    f.spill_stride = ct < max_cand ? ((cand_bytes + 255) & ~(size_t)255) : 0;
    f.nms_base = (unsigned)prefix;
    f.smem_bytes = prefix + small_bytes;
    if (f.smem_bytes < f.tk.smem_bytes) f.smem_bytes = f.tk.smem_bytes;
    if (f.smem_bytes > smem_optin) return f;
    f.ok = true;
    return f;
}

cudaError_t launch_fused(const FusedPlan& fp, const float* d_heads, int N, int sweep, int B, int max_cand, int max_keep, float nms_thr,
                         const DecodePlan& dp, const CandScratch& cs, const PostBuffers& out, unsigned char* spill,
                         const TrackBuffers& tb, TrackParams p, cudaStream_t stream) {
    FusedParams F{};
    F.heads = d_heads; F.N = N; F.sweep = sweep; F.cs = cs; F.nseg = dp.nseg; F.segcap = dp.segcap;
    F.Ccap = max_cand; F.CT = fp.CT; F.Kcap = max_keep; F.nms_thr = nms_thr; F.out = out;
    F.so_small = fp.so_small; F.so_big = fp.so_big; F.nms_base = fp.nms_base; F.spill = spill; F.spill_stride = fp.spill_stride;
    p.cost_in_smem = 1; p.det_in_smem = 1; p.pred_in_smem = 1;
    p.term_floats = fp.tk.term_floats; p.cell_cap = fp.tk.cell_cap; p.so = fp.tk.so;
    const void* fn = fp.threads == 256 ? (const void*)pb_stream_kernel<256> : fp.threads == 512 ? (const void*)pb_stream_kernel<512>
                                                                                                : (const void*)pb_stream_kernel<1024>;
    const cudaError_t e = ensure_dyn_smem(fn, fp.smem_bytes);
    if (e != cudaSuccess) return e;
    if (fp.threads == 256) pb_stream_kernel<256><<<B, 256, fp.smem_bytes, stream>>>(F, tb, p);
    else if (fp.threads == 512) pb_stream_kernel<512><<<B, 512, fp.smem_bytes, stream>>>(F, tb, p);
    else pb_stream_kernel<1024><<<B, 1024, fp.smem_bytes, stream>>>(F, tb, p);
    count_launch();
    return cudaGetLastError();
}

}  // namespace pb
