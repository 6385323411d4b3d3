// pb_common.cuh — shared declarations of the sm_100a kernels (internal).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pb_math.h"
#include "../../include/posebyte_b200.h"

namespace pb {

constexpr int KP = 17;
constexpr int POSE_F = 51;                 // floats per pose (17 x (x,y,conf))
constexpr int ST_TENTATIVE = 0, ST_CONFIRMED = 1, ST_LOST = 2;   // reference gpu_tracker.cu:23-25
constexpr int HEAD_ROWS = 56;

__device__ __constant__ const float kSigmas[KP] = {
    0.026f, 0.025f, 0.025f, 0.035f, 0.035f, 0.079f, 0.079f, 0.072f, 0.072f,
    0.062f, 0.062f, 0.107f, 0.107f, 0.087f, 0.087f, 0.089f, 0.089f};

// ---- post-process (decode + NMS) buffers, one slab per stream ------------------------
struct PostBuffers {
    float* det_poses;    // [B, Kcap, 51]  kept detections in score order
    float* det_bboxes;   // [B, Kcap, 4]
    float* det_scores;   // [B, Kcap]
    int* keep_slots;     // [B, Kcap]  candidate slot (anchor-ordered list index)
    int* keep_anchors;   // [B, Kcap]  anchor index
    int* num_keep;       // [B]
    int* num_cand;       // [B]
    unsigned long long* stage_ns;  // [B, 16] list, rank, load, nms, output (ns sums), [7] = launches
    unsigned long long* dbg;       // optional timeline [64 launches][B][6] (PB_TIMELINE=1): [3] NMS begin, [4] NMS end
    int dbg_slot;
    int* ready;          // [B] or nullptr: the stand-alone NMS kernel stores ready_seq here (release) once a stream's kept detections are
    int ready_seq;       //     written — the resident tracker kernel (tracker.cu: pb_tracker_seq_kernel) acquires it per stream-frame
};

// ---- candidate scratch between the decode+gather kernel and the NMS kernel (L2 resident) ----
struct CandScratch {
    float* records;      // [B, nseg, segcap, 56]  one contiguous head column per candidate
    int* anchors;        // [B, nseg, segcap]
    int* counts;         // [B, nseg]
};
struct DecodePlan { int nseg, groups_per_seg, segcap; };

// ---- tracker state, struct-of-arrays over streams (all persistent across frames) ------
struct TrackBuffers {
    float* poses;        // [B, T, 51]
    float* vel;          // [B, T, 34]
    float* scores;       // [B, T]
    float* predicted;    // [B, T, 51]
    float* tcent;        // [B, T, 4]
    float* dcent;        // [B, Dm, 4]
    float* cost;         // [B, T*Dm]   flat, stride = this frame's D (reference quirk Q1)
    float* det_scores;   // [B, Dm]
    int* states;         // [B, T]
    int* ids;            // [B, T]
    int* hits;           // [B, T]
    int* ages;           // [B, T]
    int* last_frame;     // [B, T]
    int* active;         // [B, T]
    int* pred_dirty;     // [B, T]  predicted[t] changed since tcent[t] was last derived from it
    int* row_assign;     // [B, T]
    int* col_assign;     // [B, Dm]
    int* scalars;        // [B, 4]  next_id, slot_hint, D, num_active
    const float* out_xform;  // [B, 4] scale_x, scale_y, pad_x, pad_y applied to the output records, or nullptr
    void* outputs;       // [B, Dm] TrackOutput (228 B)
    int* num_outputs;    // [B]
    float* det_poses_scratch;  // [B, Dm, 51]  used when the detections do not fit in shared memory
    unsigned long long* stage_ns;  // [B, 12] globaltimer stamps per stage (telemetry)
    int* seq_done;       // [B] sequence number of the last tracker launch whose STATE update this stream has completed (see pb_tracker_kernel)
    int* out_done;       // [B] ... whose TrackOutput records are written as well (second release)
    int* error_flag;     // [1] set when a stream's predecessor did not finish within the time-out
    unsigned* gate_g;    // [B, T*Dw] large tables: the gates the row-sliced pre-kernel (tracker.cu) computed for this frame
    unsigned* lgate_g;   // [B, T*Dw]
    float* tarea_g;      // [B, T] keypoint-box areas of the predicted poses (pre-kernel)
    unsigned long long* chain;   // [B] fused path: who runs the stream's next tracker stage — (published-NMS mask << 32) | next seq << 1 | busy
    unsigned long long* dbg;   // optional timeline [64 launches][B][6] (PB_TIMELINE=1): [0] begin, [1] state acquired, [2] end
};

// byte offsets of a kernel's shared-memory arrays, in the declaration order of its layout struct
struct SmemOffsets { unsigned off[48]; };

struct TrackParams {
    int B, T, Dm;
    float new_track_thresh;
    int max_age, min_hits, gating_enabled;
    int frame_id;
    int seq;             // sequence number of this launch; the CTA of stream b starts once seq_done[b] == seq - 1
    // where the large per-frame arrays live (1 = shared memory, 0 = global scratch / state)
    int cost_in_smem, det_in_smem, pred_in_smem, term_floats, cell_cap;
    int precomputed;     // predict, centres, gates and the tier-1 cost pass of this frame were done by the pre-kernel (large tables)
    int bulk_off;        // 1: the stand-alone tracker kernel loads the state slabs element by element instead of by bulk copy (A/B switch, PB_NO_BULK)
    int sub_solve_off;   // 1: tiers 2 and 3 of larger tables keep the CTA-wide / wide solve (A/B switch, PB_NO_SUB_SOLVE)
    SmemOffsets so;      // shared-memory layout (tracker_plan)
};

struct DetSource {
    const float* poses;   // [B, stride, 51]
    const float* scores;  // [B, stride]
    const int* num;       // [B]
    int stride;
};

// ---- warp helpers ---------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_stream_f(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// host-side launchers (defined in the .cu files, called from pb_api.cu)
size_t decode_nms_smem_bytes(int max_cand, int max_keep);
DecodePlan decode_plan(int B, int N, int max_cand);
cudaError_t launch_decode_gather(const float* d_heads, int B, int N, float conf_thr, bool lazy_keypoints, const DecodePlan& plan,
                                 const CandScratch& cs, cudaStream_t stream);
cudaError_t launch_nms(const float* d_heads, int N, int sweep /*0 complete, 1 lazy, 2 deferred*/, int B, int max_cand, int max_keep, float nms_thr,
                       const DecodePlan& plan, const CandScratch& cs, const PostBuffers& out, cudaStream_t stream);

// tiered NMS kernel (decode_nms.cu: pb_nms_tier_kernel): per_sm CTAs of `threads` threads share an SM (2 x 512, 3 x 384, 3 x 256,
// 4 x 256), shared memory for CT candidates, spill path beyond
struct NmsTierPlan { bool ok; int CT; size_t smem_bytes, spill_stride; SmemOffsets so_small, so_big; int threads, per_sm; };
NmsTierPlan nms_tier_plan(int max_cand, int max_keep, size_t smem_optin, int threads = 512, int per_sm = 2, int target_kb = 0);
cudaError_t launch_nms_tier(const NmsTierPlan& tp, const float* d_heads, int N, int sweep, int B, int max_cand, int max_keep, float nms_thr,
                            const DecodePlan& plan, const CandScratch& cs, const PostBuffers& out, unsigned char* spill, cudaStream_t stream);

cudaError_t preload_post_kernels(int max_cand, int max_keep, const NmsTierPlan* tier);

struct TrackerPlan { size_t smem_bytes, prefix_bytes; int threads; int cost_in_smem, det_in_smem, pred_in_smem, term_floats, cell_cap; SmemOffsets so;
                     int pre_slices; size_t pre_smem; int resident; };   // pre_slices > 0: large tables, row-sliced pre-kernel with this many CTAs per stream
TrackerPlan tracker_plan(int T, int Dm, bool compact = false, bool resident = false);
cudaError_t launch_tracker(const TrackBuffers& tb, TrackParams p, const DetSource& src,
                           const TrackerPlan& plan, cudaStream_t stream);
// large tables: predict + centres + gates + tier-1 OKS cost pass, row-sliced over plan.pre_slices CTAs per stream
cudaError_t launch_tracker_pre(const TrackBuffers& tb, const TrackParams& p, const DetSource& src, const TrackerPlan& plan, cudaStream_t stream);
void tracker_plan_pre(TrackerPlan& plan, int B, int T, int Dm, int sm_count, size_t smem_optin);
cudaError_t launch_tracker_reset(const TrackBuffers& tb, int B, int T, int Dm, int seq, cudaStream_t stream);

// Ring of per-step buffers as the fused kernel sees it: a CTA that owns its stream's chain goes on with the following
// frames, whose kept detections were published by other CTAs in their steps' slots (slot = seq % depth).
constexpr int PB_MAX_RING = 12;            // <= 15: the chain word holds a 15-bit mask of published frames
struct RingTable {
    const float* det_poses[PB_MAX_RING];   // [B, Kcap, 51]
    const float* det_scores[PB_MAX_RING];  // [B, Kcap]
    const int* num_keep[PB_MAX_RING];      // [B]
    void* outputs[PB_MAX_RING];            // [B, Dm] TrackOutput
    int* num_outputs[PB_MAX_RING];         // [B]
    int* frame_id[PB_MAX_RING];            // [B] frame id of the step whose NMS stage published into this slot
    int depth;
    int wait_window;                       // a CTA whose frame is fewer than this many frames from the head of its chain waits for its turn (tracker_body.cuh)
};

// Resident tracker (tracker.cu: pb_tracker_seq_kernel): one CTA per video stream runs the tracker stages of `n` consecutive
// frames in ONE launch; frame i's kept detections come from slot i of this table as soon as the NMS kernel of that step has
// published them (PostBuffers::ready).  Frame i has sequence number TrackParams::seq + i and frame id TrackParams::frame_id + i.
constexpr int PB_SEQ_MAX = 32;
struct SeqTable {
    const float* det_poses[PB_SEQ_MAX];    // [B, stride, 51]
    const float* det_scores[PB_SEQ_MAX];   // [B, stride]
    const int* num_keep[PB_SEQ_MAX];       // [B]
    const int* ready[PB_SEQ_MAX];          // [B]
    void* outputs[PB_SEQ_MAX];             // [B, Dm] TrackOutput
    int* num_outputs[PB_SEQ_MAX];          // [B]
    int n, stride;
};
cudaError_t launch_tracker_seq(const TrackBuffers& tb, TrackParams p, const SeqTable& q, const TrackerPlan& plan, cudaStream_t stream);

// fused per-stream kernel (fused.cu): NMS + tracker of one stream-frame in one CTA
struct FusedPlan {
    bool ok;
    int threads, CT;                    // CTA size; candidates whose NMS working set fits in shared memory (more: spill path)
    size_t smem_bytes, spill_stride;    // dynamic shared memory per CTA; bytes of spill scratch per stream (0: never spills)
    unsigned nms_base;
    SmemOffsets so_small, so_big;
    TrackerPlan tk;                     // compact tracker layout (detections first)
};
FusedPlan plan_fused(int T, int Dm, int max_cand, int max_keep, size_t smem_optin);
cudaError_t launch_fused(const FusedPlan& fp, const float* d_heads, int N, int sweep, int B, int max_cand, int max_keep, float nms_thr,
                         const DecodePlan& dp, const CandScratch& cs, const PostBuffers& out, unsigned char* spill,
                         const TrackBuffers& tb, TrackParams p, const RingTable& ring, cudaStream_t stream);

void count_launch(int n = 1);
// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: raise it for `func` on the
// current device when `bytes` exceeds what was set there before (thread-safe; pb_api.cu).
cudaError_t ensure_dyn_smem(const void* func, size_t bytes);

}  // namespace pb
