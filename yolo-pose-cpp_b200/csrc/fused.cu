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
//   * order   the NMS stage depends on this frame's candidates only and runs at once.  The tracker stages of one video
//     stream must run in frame order: a per-stream chain word hands them from CTA to CTA without anybody waiting (below).
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

// The spill path (more candidates than the shared-memory tier holds) is a function of its own: it is cold, and kept out of
// the register allocation of the common path.  The kernel parameters stay in constant memory (__grid_constant__).
template <int NT>
static __device__ __noinline__ int nms_spill_stage(const FusedParams& F, const SmemOffsets& tso, int Dm, int b) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* fscore = reinterpret_cast<float*>(smem_raw + tso.off[16]);     // TkSmem::dscore
    float* fdet = reinterpret_cast<float*>(smem_raw + tso.off[34]);       // TkSmem::det
    unsigned char* fixed = smem_raw + F.nms_base;
    NmSmem s;
    nm_from_offsets(fixed, F.spill + (size_t)b * F.spill_stride, F.so_big, s);
    return nms_body<NT>(s, F.heads, F.N, F.sweep, F.cs, F.nseg, F.segcap, F.Ccap, F.Kcap, F.nms_thr, F.out, b, (int)gridDim.x, fdet, fscore, Dm);
}

template <int NT>
__global__ void __launch_bounds__(NT, NT <= 512 ? 2 : 1)
pb_stream_kernel(const __grid_constant__ FusedParams F, const __grid_constant__ TrackBuffers tb, const __grid_constant__ TrackParams P,
                 const __grid_constant__ RingTable R) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_flag;
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    // candidates of this stream-frame (sum of the decode kernel's segment counts): selects the layout
    int total = 0;
    for (int sg = 0; sg < F.nseg; ++sg) total += F.cs.counts[b * F.nseg + sg];
    if (total > F.Ccap) total = F.Ccap;
    float* fscore = reinterpret_cast<float*>(smem_raw + P.so.off[16]);     // TkSmem::dscore
    float* fdet = reinterpret_cast<float*>(smem_raw + P.so.off[34]);       // TkSmem::det
    int nkeep;
    if (total <= F.CT) {
        unsigned char* fixed = smem_raw + F.nms_base;
        NmSmem s;
        nm_from_offsets(fixed, fixed, F.so_small, s);
        nkeep = nms_body<NT>(s, F.heads, F.N, F.sweep, F.cs, F.nseg, F.segcap, F.CT, F.Kcap, F.nms_thr, F.out, b, (int)gridDim.x, fdet, fscore, P.Dm);
    } else {
        nkeep = nms_spill_stage<NT>(F, P.so, P.Dm, b);
    }
    int slot = P.seq % R.depth;
    if (tid == 0) R.frame_id[slot][b] = P.frame_id;
    __syncthreads();                    // kept detections + frame id written by all threads; the NMS stage's shared memory is free
    if (tid == 0) s_flag = chain_arrive(tb.chain + b, P.seq, R.wait_window);
    __syncthreads();
    const int how = s_flag;             // 0: the chain's owner runs this frame's tracker stage from the ring slot; 1: owner; 2: waits for its turn
    if (how == 0) return;
    int seq = P.seq;
    int wait_mode = (how == 2) ? 1 : 0; // waits inside the tracker stage, after its detection-only part, for the chain
    bool first = true;                  // first frame this CTA runs: the previous owner may still be assembling its records
    for (;;) {
        slot = seq % R.depth;
        int frame_id = P.frame_id, D = nkeep;
        if (!first) {
            // a later frame of the stream: its NMS stage ran in another CTA and left the kept detections in its ring slot
            frame_id = R.frame_id[slot][b];
            D = R.num_keep[slot][b];
            const int Dc = D < P.Dm ? D : P.Dm;
            const float* gp = R.det_poses[slot] + (size_t)b * F.Kcap * POSE_F;
            const float* gs = R.det_scores[slot] + (size_t)b * F.Kcap;
#pragma unroll 1
            for (int i = tid; i < Dc * POSE_F; i += NT) fdet[i] = gp[i];
#pragma unroll 1
            for (int d = tid; d < Dc; d += NT) fscore[d] = gs[d];
            __syncthreads();
        }
        if (tb.dbg && tid == 0) {       // PB_TIMELINE: tracker stage begin, and which step's kernel runs it
            unsigned long long* q = tb.dbg + ((size_t)(seq & 63) * gridDim.x + b) * 6;
            q[0] = globaltimer_ns(); q[1] = (unsigned long long)P.seq; q[5] = (unsigned long long)seq;
        }
        DetSource none{nullptr, nullptr, nullptr, 0};
        const int go = tracker_body<NT, true, true>(tb, P, none, b, smem_raw, D, seq, frame_id, R.outputs[slot], R.num_outputs[slot], wait_mode, first);
        if (go <= 0) return;            // the chain was handed on (or given up) when the state was released
        ++seq;
        wait_mode = 0;
        first = false;
        __syncthreads();
    }
}

// Shared-memory plan of the fused kernel for one handle; ok == false: this configuration keeps the three-kernel step.
FusedPlan plan_fused(int T, int Dm, int max_cand, int max_keep, size_t smem_optin) {
    FusedPlan f{};
    f.ok = false;
    f.tk = tracker_plan(T, Dm, getenv("PB_FUSED_COMPACT") ? atoi(getenv("PB_FUSED_COMPACT")) != 0 : true);
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
                         const TrackBuffers& tb, TrackParams p, const RingTable& ring, cudaStream_t stream) {
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
    if (fp.threads == 256) pb_stream_kernel<256><<<B, 256, fp.smem_bytes, stream>>>(F, tb, p, ring);
    else if (fp.threads == 512) pb_stream_kernel<512><<<B, 512, fp.smem_bytes, stream>>>(F, tb, p, ring);
    else pb_stream_kernel<1024><<<B, 1024, fp.smem_bytes, stream>>>(F, tb, p, ring);
    count_launch();
    return cudaGetLastError();
}

}  // namespace pb
