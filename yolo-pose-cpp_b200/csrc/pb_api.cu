// pb_api.cu — handle management and the C ABI of include/posebyte_b200.h (host C++17).
//
// A handle owns the device buffers of B independent streams (the reference allocates the
// same buffers per GPUPostprocess / GPUTracker object: gpu_postprocess.cu:319-347,
// gpu_tracker.cu:925-1010) as struct-of-arrays slabs, plus a pinned staging ring for the
// host-buffer entry point.  There is no CPU fallback: pb_create fails without a device.
#include <nvtx3/nvToolsExt.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <atomic>
#include <map>
#include <mutex>
#include <new>
#include <utility>
#include <vector>

#include "pb_common.cuh"

namespace pb {
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

cudaError_t ensure_dyn_smem(const void* func, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    static std::mutex mu;
    static std::map<std::pair<const void*, int>, size_t> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = done[{func, dev}];
    if (bytes <= have) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}
}  // namespace pb

static thread_local char g_err[512] = "";
void pb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

using namespace pb;

// NVTX range around the entry points of the path (header-only NVTX 3: costs a null check when no tool is attached; the
// reference marks nothing — SURVEY.md 5 lists the ranges as the tracing aid a profile of the frame loop needs)
namespace {
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace

// One slot of the step pipeline: candidate scratch (decode+gather -> NMS) and kept detections
// (NMS -> tracker) with the events that order their reuse.
struct PipeSlot {
    unsigned char* spill = nullptr;   // resident-tracker path: spill scratch of the tiered NMS kernel (per slot: NMS launches of several steps overlap)
    int* ready = nullptr;          // [B] resident-tracker path: sequence number of the step whose kept detections the slot holds (per stream)
    PostBuffers post{};
    CandScratch cand{};
    void* outputs = nullptr;       // [B,Dm] TrackOutput of the step that used this slot (a read-back may still be copying them
    int* num_outputs = nullptr;    //  while the next step's tracker already assembles its own)
    cudaEvent_t ev_gather = nullptr, ev_nms = nullptr, ev_trk = nullptr;
    bool used = false;
};

struct pb_handle_st {
    pb_config cfg;
    PostBuffers post{};            // == ring[cur].post (the slot of the most recent step)
    CandScratch cand{};            // == ring[cur].cand
    std::vector<PipeSlot> ring;    // cfg.pipeline_depth slots
    int cur = 0;
    bool inflight = false;         // work of a pipelined pb_step may still run on the internal streams
    cudaStream_t s_nms = nullptr, s_trk = nullptr, s_trk2 = nullptr, s_trk3 = nullptr;   // lanes: step i's NMS and tracker run on lane i % lanes
    cudaStream_t f_lane[2 * PB_MAX_RING] = {};   // fused path: step i's kernel runs on f_lane[i % f_lanes]
    int f_lanes = 1;
    cudaStream_t s_rb = nullptr;   // fused path: read-back copies (pb_submit_host), ordered behind every kernel that can write the step's records
    RingTable rtab{};
    bool borrow_until_wait = false;   // pb_submit_host: the head buffer belongs to the library until pb_wait, no ordering on the caller's stream needed
    int lanes = 0;                 // number of lanes (0: NMS on s_nms, trackers alternate between s_trk and s_trk2)
    int trk_seq = 0;               // sequence number of the last tracker launch (TrackParams::seq)
    cudaStream_t last_trk_stream = nullptr;
    bool last_was_readback = false;
    bool overlap_trackers = false; // two tracker grids (one CTA per SM each) always fit the device: 2 * num_streams <= SM count
    DecodePlan dplan{};
    TrackBuffers trk{};
    TrackerPlan plan{};
    FusedPlan fplan{};             // fused per-stream kernel (fused.cu); fplan.ok == false: separate NMS and tracker kernels
    unsigned char* d_spill = nullptr;
    // resident-tracker path of pb_step_seq (tracker.cu: pb_tracker_seq_kernel): its own ring of 2 * seq_chunk slots, allocated at the
    // first use; a chunk of steps = decode + NMS launches per step and ONE tracker launch; the two halves of the ring alternate
    std::vector<PipeSlot> seq_ring;
    int seq_chunk = 0;             // steps per tracker launch (0: path not available for this handle)
    int seq_half = 0;              // half of the ring the next chunk uses
    bool seq_half_used[2] = {false, false};
    bool seq_inflight = false;     // kernels of the resident path may still be running
    cudaEvent_t ev_seq_trk[2] = {nullptr, nullptr};   // per half: the tracker launch of the last chunk that used it
    static constexpr int SEQ_MAX_LANES = 4;
    int seq_lanes = 3;             // step i's decode and NMS kernels run back to back on lane i % seq_lanes
    cudaEvent_t ev_seq_nms[SEQ_MAX_LANES] = {};       // per lane: its last NMS launch
    cudaEvent_t ev_seq_dec[SEQ_MAX_LANES] = {};       // per lane: its last decode launch (the last reader of a borrowed head tensor)
    cudaEvent_t ev_seq_start = nullptr;
    cudaStream_t s_seq_nms[SEQ_MAX_LANES] = {}, s_seq_trk = nullptr;
    unsigned char* seq_spill[SEQ_MAX_LANES] = {};     // per lane: spill scratch of the tiered NMS kernel (its launches on a lane are serial)
    TrackerPlan seq_plan{};
    NmsTierPlan exp_tier{}; unsigned char* exp_spill = nullptr;   // PB_NMS_TIER experiment (serial path)
    int sub_solve_off = 0, bulk_off = 0;   // A/B switches read at pb_create: PB_NO_SUB_SOLVE, PB_NO_BULK (1 none, 2 centres + cost matrix only, 3 slabs only)
    NmsTierPlan pipe_tier{};       // ok: the pipelined three-kernel step launches the tiered NMS kernel (CTAs that share an SM)
    NmsTierPlan seq_nms{};         // ok: the steps of the resident path use the tiered (half-SM) NMS kernel
    std::vector<void*> allocs;
    // host-buffer path
    float* d_stage = nullptr;      // [B,56,N] device staging
    void* h_out_pinned = nullptr;  // [B,Dm] TrackOutput
    int* h_cnt_pinned = nullptr;   // [B]
    cudaStream_t own_stream = nullptr;
    int frames = 0;
    bool lazy_keypoints = false;   // set by pb_step_host while the head is read in place from host memory
    float* d_xform = nullptr;      // [B,4] output transform (pb_set_output_transform)
    void* rb_tracks = nullptr;     // page-locked read-back targets of the step being enqueued (pb_submit_host)
    int* rb_counts = nullptr;
    // optional per-kernel event timing (pb_set_profiling)
    bool profiling = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<std::pair<int, int>> ev_post, ev_track, ev_gather;   // indices into ev_pool (begin, end)
    size_t ev_used = 0;
};

static int prof_event(pb_handle_st* h) {
    if (h->ev_used == h->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return -1;
        h->ev_pool.push_back(e);
    }
    return (int)h->ev_used++;
}

// Every entry point that takes a handle runs on the handle's device and gives the caller's current device back.
struct DevGuard {
    int prev = -1;
    bool switched = false;
    explicit DevGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DevGuard() { if (switched) cudaSetDevice(prev); }
    DevGuard(const DevGuard&) = delete;
    DevGuard& operator=(const DevGuard&) = delete;
};

#define PB_CUDA(call)                                                                         \
    do { cudaError_t e_ = (call);                                                             \
         if (e_ != cudaSuccess) { pb_set_error("%s failed: %s", #call, cudaGetErrorString(e_)); return PB_ERR_CUDA; } } while (0)

template <typename T>
static int dev_alloc(pb_handle_st* h, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 16);
    if (e != cudaSuccess) { pb_set_error("cudaMalloc(%zu) failed: %s", count * sizeof(T), cudaGetErrorString(e)); return PB_ERR_CUDA; }
    e = cudaMemset(q, 0, count * sizeof(T) + 16);
    if (e != cudaSuccess) { pb_set_error("cudaMemset failed: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    h->allocs.push_back(q);
    *p = static_cast<T*>(q);
    return PB_OK;
}
#define PB_TRY(x) do { int r_ = (x); if (r_ != PB_OK) return r_; } while (0)

// Tracker state snapshot: header + every persistent slab, in one host blob.
struct SnapHeader { unsigned magic, version; int B, T, Dm, frames; };
static const unsigned kSnapMagic = 0x50425354u;   // "PBST"

template <typename F>
static int for_each_state_slab(pb_handle_st* h, F&& f) {
    const size_t B = h->cfg.num_streams, T = h->cfg.max_tracks, Dm = h->cfg.max_detections;
    TrackBuffers& t = h->trk;
    PB_TRY(f(t.poses, B * T * POSE_F * 4)); PB_TRY(f(t.vel, B * T * 34 * 4)); PB_TRY(f(t.scores, B * T * 4));
    PB_TRY(f(t.predicted, B * T * POSE_F * 4)); PB_TRY(f(t.tcent, B * T * 16)); PB_TRY(f(t.dcent, B * Dm * 16));
    PB_TRY(f(t.cost, B * T * Dm * 4)); PB_TRY(f(t.det_scores, B * Dm * 4));
    PB_TRY(f(t.states, B * T * 4)); PB_TRY(f(t.ids, B * T * 4)); PB_TRY(f(t.hits, B * T * 4)); PB_TRY(f(t.ages, B * T * 4));
    PB_TRY(f(t.last_frame, B * T * 4)); PB_TRY(f(t.active, B * T * 4)); PB_TRY(f(t.pred_dirty, B * T * 4));
    PB_TRY(f(t.row_assign, B * T * 4)); PB_TRY(f(t.col_assign, B * Dm * 4)); PB_TRY(f(t.scalars, B * 4 * 4));
    return PB_OK;
}

extern "C" {

static int check_order_flag(pb_handle_st* h);

const char* pb_last_error(void) { return g_err; }
const char* pb_version(void) { return "posebyte-b200 0.1 (sm_100a)"; }
long long pb_launch_count(void) { return g_launches.load(); }

void pb_default_config(pb_config* c) {
    if (!c) return;
    c->num_streams = 1;
    c->num_anchors = 8400;            // gpu_postprocess.h:13
    c->max_candidates = 1024;         // gpu_postprocess.h:13
    c->max_keep = 256;                // gpu_postprocess.cu:224
    c->max_tracks = 128;              // gpu_tracker.h:17-25
    c->max_detections = 64;
    c->match_threshold = 0.5f;
    c->high_thresh = 0.30f;
    c->low_thresh = 0.15f;
    c->new_track_thresh = 0.30f;
    c->max_age = 10;
    c->min_hits = 3;
    c->use_cuda_graph = 0;
    c->gating_enabled = 1;
    c->device = 0;
    c->pipeline_depth = 1;
    c->keypoint_fetch = 0;
    c->fuse_stages = 2;
}

static void seq_configure(pb_handle_st* h, int sm_count);
static NmsTierPlan tier_plan_by_id(const pb_config& c, int id);
static int smem_config_kb(size_t bytes);

static int build_handle(pb_handle_st* h) {
    const pb_config& c = h->cfg;
    const size_t B = c.num_streams, T = c.max_tracks, Dm = c.max_detections, K = c.max_keep;
    h->dplan = decode_plan(c.num_streams, c.num_anchors, c.max_candidates);
    const size_t nsc = B * (size_t)h->dplan.nseg * (size_t)h->dplan.segcap;
    unsigned long long* post_ns = nullptr;
    PB_TRY(dev_alloc(h, &post_ns, B * 16));
    h->ring.resize((size_t)c.pipeline_depth);
    for (PipeSlot& sl : h->ring) {
        PB_TRY(dev_alloc(h, &sl.post.det_poses, B * K * POSE_F));
        PB_TRY(dev_alloc(h, &sl.post.det_bboxes, B * K * 4));
        PB_TRY(dev_alloc(h, &sl.post.det_scores, B * K));
        PB_TRY(dev_alloc(h, &sl.post.keep_slots, B * K));
        PB_TRY(dev_alloc(h, &sl.post.keep_anchors, B * K));
        PB_TRY(dev_alloc(h, &sl.post.num_keep, B));
        PB_TRY(dev_alloc(h, &sl.post.num_cand, B));
        sl.post.stage_ns = post_ns;
        PB_TRY(dev_alloc(h, &sl.cand.records, nsc * HEAD_ROWS));
        PB_TRY(dev_alloc(h, &sl.cand.anchors, nsc));
        PB_TRY(dev_alloc(h, &sl.cand.counts, B * (size_t)h->dplan.nseg));
        unsigned char* outp = nullptr;
        PB_TRY(dev_alloc(h, &outp, B * Dm * 228));
        sl.outputs = outp;
        PB_TRY(dev_alloc(h, &sl.num_outputs, B));
        if (h->pipe_tier.ok && h->pipe_tier.spill_stride) PB_TRY(dev_alloc(h, &sl.spill, B * h->pipe_tier.spill_stride));
        PB_CUDA(cudaEventCreateWithFlags(&sl.ev_gather, cudaEventDisableTiming));
        PB_CUDA(cudaEventCreateWithFlags(&sl.ev_nms, cudaEventDisableTiming));
        PB_CUDA(cudaEventCreateWithFlags(&sl.ev_trk, cudaEventDisableTiming));
    }
    h->cur = 0; h->post = h->ring[0].post; h->cand = h->ring[0].cand;
    if (c.pipeline_depth > 1) {
        PB_CUDA(cudaStreamCreateWithFlags(&h->s_nms, cudaStreamNonBlocking));
        PB_CUDA(cudaStreamCreateWithFlags(&h->s_trk, cudaStreamNonBlocking));
        PB_CUDA(cudaStreamCreateWithFlags(&h->s_trk2, cudaStreamNonBlocking));
        PB_CUDA(cudaStreamCreateWithFlags(&h->s_trk3, cudaStreamNonBlocking));
        for (int i = 0; i < h->f_lanes; ++i) PB_CUDA(cudaStreamCreateWithFlags(&h->f_lane[i], cudaStreamNonBlocking));
        PB_CUDA(cudaStreamCreateWithFlags(&h->s_rb, cudaStreamNonBlocking));
    }
    TrackBuffers& t = h->trk;
    PB_TRY(dev_alloc(h, &t.poses, B * T * POSE_F));
    PB_TRY(dev_alloc(h, &t.vel, B * T * 34));
    PB_TRY(dev_alloc(h, &t.scores, B * T));
    PB_TRY(dev_alloc(h, &t.predicted, B * T * POSE_F));
    PB_TRY(dev_alloc(h, &t.tcent, B * T * 4));
    PB_TRY(dev_alloc(h, &t.dcent, B * Dm * 4));
    PB_TRY(dev_alloc(h, &t.cost, B * T * Dm));
    PB_TRY(dev_alloc(h, &t.det_scores, B * Dm));
    PB_TRY(dev_alloc(h, &t.states, B * T));
    PB_TRY(dev_alloc(h, &t.ids, B * T));
    PB_TRY(dev_alloc(h, &t.hits, B * T));
    PB_TRY(dev_alloc(h, &t.ages, B * T));
    PB_TRY(dev_alloc(h, &t.last_frame, B * T));
    PB_TRY(dev_alloc(h, &t.active, B * T));
    PB_TRY(dev_alloc(h, &t.pred_dirty, B * T));
    PB_TRY(dev_alloc(h, &t.row_assign, B * T));
    PB_TRY(dev_alloc(h, &t.col_assign, B * Dm));
    PB_TRY(dev_alloc(h, &t.scalars, B * 4));
    t.outputs = h->ring[0].outputs; t.num_outputs = h->ring[0].num_outputs;
    PB_TRY(dev_alloc(h, &t.det_poses_scratch, B * Dm * POSE_F));
    PB_TRY(dev_alloc(h, &t.stage_ns, B * 20));
    PB_TRY(dev_alloc(h, &t.seq_done, B));
    PB_TRY(dev_alloc(h, &t.out_done, B));
    PB_TRY(dev_alloc(h, &t.error_flag, 1));
    PB_TRY(dev_alloc(h, &t.chain, B));
    h->rtab.depth = (int)h->ring.size();
    {   // CTAs that wait for their turn hold an SM slot each: wait_window * num_streams stays below the SM count
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int w = 2;
        while (w > 0 && (long)w * c.num_streams >= sms) --w;
        if (h->rtab.depth < 2) w = 0;
        if (const char* e = getenv("PB_FUSED_WAIT")) { const int a = atoi(e); if (a >= 0 && a <= 14) w = a; }
        h->rtab.wait_window = w;
    }
    for (size_t i = 0; i < h->ring.size(); ++i) {
        PipeSlot& sl = h->ring[i];
        h->rtab.det_poses[i] = sl.post.det_poses; h->rtab.det_scores[i] = sl.post.det_scores; h->rtab.num_keep[i] = sl.post.num_keep;
        h->rtab.outputs[i] = sl.outputs; h->rtab.num_outputs[i] = sl.num_outputs;
        PB_TRY(dev_alloc(h, &h->rtab.frame_id[i], B));
    }
    if (getenv("PB_TIMELINE")) {                       // development aid: absolute begin/end times of the last 64 launches
        PB_TRY(dev_alloc(h, &t.dbg, (size_t)64 * B * 6));
        for (PipeSlot& sl : h->ring) sl.post.dbg = t.dbg;
        h->post.dbg = t.dbg;
    }
    h->plan = tracker_plan(c.max_tracks, c.max_detections);
    {
        int dev = 0, sms = 0, optin = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        tracker_plan_pre(h->plan, c.num_streams, c.max_tracks, c.max_detections, sms, (size_t)optin);
        if (h->plan.pre_slices > 0) {
            const size_t Dw = (Dm + 31) / 32;
            PB_TRY(dev_alloc(h, &t.gate_g, B * T * Dw));
            PB_TRY(dev_alloc(h, &t.lgate_g, B * T * Dw));
            PB_TRY(dev_alloc(h, &t.tarea_g, B * T));
        }
    }
    {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        seq_configure(h, sms);
    }
    if (h->fplan.ok && h->fplan.spill_stride) PB_TRY(dev_alloc(h, &h->d_spill, B * h->fplan.spill_stride));
    PB_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    PB_CUDA(launch_tracker_reset(h->trk, c.num_streams, c.max_tracks, c.max_detections, h->trk_seq, h->own_stream));
    PB_CUDA(cudaStreamSynchronize(h->own_stream));
    return PB_OK;
}

int pb_create(const pb_config* cfg, pb_handle_t* out) {
    if (!cfg || !out) { pb_set_error("pb_create: null argument"); return PB_ERR_INVALID; }
    *out = nullptr;
    const pb_config& c = *cfg;
    if (c.num_streams < 1 || c.num_anchors < 1 || c.max_candidates < 1 || c.max_keep < 1 ||
        c.max_tracks < 1 || c.max_detections < 1) { pb_set_error("pb_create: sizes must be positive"); return PB_ERR_INVALID; }
    if (c.num_anchors > 65536) { pb_set_error("pb_create: num_anchors > 65536 unsupported"); return PB_ERR_UNSUPPORTED; }
    if (c.max_tracks >= 65536 || c.max_detections >= 65536) { pb_set_error("pb_create: max_tracks/max_detections too large"); return PB_ERR_UNSUPPORTED; }
    if ((unsigned long long)c.max_tracks * (unsigned long long)c.max_detections * (unsigned long long)c.max_detections >= (1ull << 32)) {
        // the tracker's stage loops divide cell indices by the frame's detection count with a 32-bit magic multiplier
        pb_set_error("pb_create: max_tracks * max_detections^2 must stay below 2^32");
        return PB_ERR_UNSUPPORTED;
    }
    if (c.max_keep > c.max_candidates) { pb_set_error("pb_create: max_keep > max_candidates"); return PB_ERR_INVALID; }
    if (c.fuse_stages < 0 || c.fuse_stages > 2) { pb_set_error("pb_create: fuse_stages must be 0, 1 or 2"); return PB_ERR_INVALID; }
    if (c.keypoint_fetch < 0 || c.keypoint_fetch > 3) { pb_set_error("pb_create: keypoint_fetch must be 0..3"); return PB_ERR_INVALID; }
    if (c.pipeline_depth < 1 || c.pipeline_depth > PB_MAX_RING) { pb_set_error("pb_create: pipeline_depth must be 1..%d", PB_MAX_RING); return PB_ERR_INVALID; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= c.device) {
        (void)cudaGetLastError();
        pb_set_error("pb_create: CUDA device %d not available (this library has no CPU path)", c.device);
        return PB_ERR_NO_DEVICE;
    }
    DevGuard dev_guard(c.device);           // the caller's current device is restored on return
    { int cur = -1; if (cudaGetDevice(&cur) != cudaSuccess || cur != c.device) { pb_set_error("pb_create: cannot select device %d", c.device); return PB_ERR_CUDA; } }
    cudaDeviceProp prop{};
    PB_CUDA(cudaGetDeviceProperties(&prop, c.device));
    if (prop.major < 10) { pb_set_error("pb_create: device sm_%d%d, built for sm_100a only", prop.major, prop.minor); return PB_ERR_NO_DEVICE; }
    if (decode_nms_smem_bytes(c.max_candidates, c.max_keep) > (size_t)prop.sharedMemPerBlockOptin) {
        pb_set_error("pb_create: max_candidates=%d needs %zu B of shared memory (> %zu)", c.max_candidates,
                     decode_nms_smem_bytes(c.max_candidates, c.max_keep), (size_t)prop.sharedMemPerBlockOptin);
        return PB_ERR_UNSUPPORTED;
    }
    {
        const TrackerPlan tp = tracker_plan(c.max_tracks, c.max_detections);
        if (tp.smem_bytes > (size_t)prop.sharedMemPerBlockOptin) {
            pb_set_error("pb_create: max_tracks=%d / max_detections=%d need %zu B of shared memory (> %zu)", c.max_tracks,
                         c.max_detections, tp.smem_bytes, (size_t)prop.sharedMemPerBlockOptin);
            return PB_ERR_UNSUPPORTED;
        }
    }
    if (const char* g = getenv("PB_L2_FETCH")) {       // experiment: 32 / 64 / 128 byte L2 fetch granularity
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g));
    }
    pb_handle_st* h = new (std::nothrow) pb_handle_st();
    if (!h) { pb_set_error("pb_create: out of host memory"); return PB_ERR_INVALID; }
    h->cfg = c;
    h->lazy_keypoints = (c.keypoint_fetch == 1);
    if (getenv("PB_NO_SUB_SOLVE")) h->sub_solve_off = 1;
    if (const char* e = getenv("PB_NO_BULK")) h->bulk_off = atoi(e);
    // fused per-stream kernel: always when asked for (1); by default (2) where it measured faster than the three-kernel step —
    // handles whose tracker grids cannot overlap (2 * num_streams > SM count): 128 streams 1.95 M against 1.55 M stream-frames/s,
    // 148 streams 2.30 M against 2.02 M; 64 streams 1.88 M against 2.04 M, one stream 25 us against 22 us per frame
    const bool want_fused = c.fuse_stages == 1 || (c.fuse_stages == 2 && 2 * c.num_streams > prop.multiProcessorCount) || getenv("PB_FORCE_FUSED");
    if (want_fused && !getenv("PB_NO_FUSED"))
        h->fplan = plan_fused(c.max_tracks, c.max_detections, c.max_candidates, c.max_keep, (size_t)prop.sharedMemPerBlockOptin);
    if (getenv("PB_DEBUG_PLAN"))
        fprintf(stderr, "[pb] fused plan: ok %d threads %d tier %d smem %zu (tracker %zu, prefix %zu) spill/stream %zu | separate: nms %zu tracker %zu\n",
                (int)h->fplan.ok, h->fplan.threads, h->fplan.CT, h->fplan.smem_bytes, h->fplan.tk.smem_bytes, h->fplan.tk.prefix_bytes,
                h->fplan.spill_stride, decode_nms_smem_bytes(c.max_candidates, c.max_keep), tracker_plan(c.max_tracks, c.max_detections).smem_bytes);
    h->overlap_trackers = 2 * c.num_streams <= prop.multiProcessorCount && !getenv("PB_NO_TRACKER_OVERLAP");
    // Lanes: the oldest tracker grid in flight never waits, the younger ones may spin on their predecessors and
    // hold one SM per CTA while they do: (lanes - 1) * num_streams must stay below the SM count.
    h->lanes = 0;
    // Measured at 64 streams: three lanes 37.0 us per batch, two lanes 44.8 (a step's NMS then waits for the tracker two
    // steps back), the separate NMS stream 38-40.
    if (h->overlap_trackers && 2 * c.num_streams < prop.multiProcessorCount && c.pipeline_depth >= 4) h->lanes = 3;
    if (h->fplan.ok && c.pipeline_depth >= 2) {
        // fused path: nobody waits on the device (chain hand-off, fused.cu), so the number of steps in flight is not bounded
        // by the SM count; one ring slot stays free so that a step's decode does not wait for the oldest step in flight
        h->f_lanes = 2 * c.pipeline_depth;         // streams are cheap: a lane is never what a step waits for (the ring slot is)
        if (const char* e = getenv("PB_FUSED_LANES")) { const int l = atoi(e); if (l >= 1 && l <= 2 * PB_MAX_RING) h->f_lanes = l; }
    }
    if (const char* e = getenv("PB_LANES")) {
        const int l = atoi(e);
        if (l >= 0 && l <= 3 && (l <= 1 || (l - 1) * c.num_streams < prop.multiProcessorCount)) h->lanes = l;
    }
    if (!h->fplan.ok) {
        // The NMS kernel in the shared-memory configuration of the tracker kernel.  An SM splits its 256 KB between L1 and
        // shared memory in steps (8 ... 100, 132, 164, 196, 228 KB) and changes the split only when it is empty; NMS CTAs
        // sized for the full candidate cap (224 KB at 1024 candidates) and tracker CTAs (137 KB: the 164 KB configuration)
        // then keep draining each other's SMs.  Measured at 64 streams: 31.5 us per step with the 224 KB kernel, 30.6 us at
        // 200 KB, 26.8 us at 160 KB and below.  So the NMS kernel takes the tracker's configuration: shared memory for the
        // candidates that fit (704 at 163 KB), the spill path for a stream with more (same results).  PB_PIPE_NMS_TIER: 0 off,
        // 1-4 the small-CTA variants, >= 64 that many KB.
        int id = -1;
        if (const char* e = getenv("PB_PIPE_NMS_TIER")) id = atoi(e);
        if (id < 0) {
            const TrackerPlan tp = tracker_plan(c.max_tracks, c.max_detections);
            const int cls = smem_config_kb(tp.smem_bytes);
            if (cls >= 100 && (size_t)cls * 1024 < decode_nms_smem_bytes(c.max_candidates, c.max_keep) + 1024) id = cls - 1;
        }
        if (id > 0) h->pipe_tier = tier_plan_by_id(c, id);
        if (h->pipe_tier.ok && h->pipe_tier.CT < 128 && h->pipe_tier.CT < c.max_candidates) h->pipe_tier = NmsTierPlan{};
    }
    int r = build_handle(h);
    if (r != PB_OK) { pb_destroy(h); return r; }
    *out = h;
    return PB_OK;
}

int pb_destroy(pb_handle_t h) {
    if (!h) return PB_OK;
    DevGuard dev_guard(h->cfg.device);
    cudaDeviceSynchronize();          // pipelined steps may still be running on the internal streams
    for (void* p : h->allocs) cudaFree(p);
    if (h->d_stage) cudaFree(h->d_stage);
    if (h->h_out_pinned) cudaFreeHost(h->h_out_pinned);
    if (h->h_cnt_pinned) cudaFreeHost(h->h_cnt_pinned);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->s_nms) cudaStreamDestroy(h->s_nms);
    if (h->s_trk) cudaStreamDestroy(h->s_trk);
    if (h->s_trk2) cudaStreamDestroy(h->s_trk2);
    if (h->s_trk3) cudaStreamDestroy(h->s_trk3);
    for (cudaStream_t q : h->f_lane) if (q) cudaStreamDestroy(q);
    if (h->s_rb) cudaStreamDestroy(h->s_rb);
    for (int i = 0; i < 2; ++i) if (h->ev_seq_trk[i]) cudaEventDestroy(h->ev_seq_trk[i]);
    for (int i = 0; i < pb_handle_st::SEQ_MAX_LANES; ++i) {
        if (h->s_seq_nms[i]) cudaStreamDestroy(h->s_seq_nms[i]);
        if (h->ev_seq_nms[i]) cudaEventDestroy(h->ev_seq_nms[i]);
        if (h->ev_seq_dec[i]) cudaEventDestroy(h->ev_seq_dec[i]);
    }
    if (h->s_seq_trk) cudaStreamDestroy(h->s_seq_trk);
    if (h->ev_seq_start) cudaEventDestroy(h->ev_seq_start);
    for (PipeSlot& sl : h->seq_ring) {
        if (sl.ev_gather) cudaEventDestroy(sl.ev_gather);
        if (sl.ev_nms) cudaEventDestroy(sl.ev_nms);
        if (sl.ev_trk) cudaEventDestroy(sl.ev_trk);
    }
    for (PipeSlot& sl : h->ring) {
        if (sl.ev_gather) cudaEventDestroy(sl.ev_gather);
        if (sl.ev_nms) cudaEventDestroy(sl.ev_nms);
        if (sl.ev_trk) cudaEventDestroy(sl.ev_trk);
    }
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    delete h;
    return PB_OK;
}

// Make `stream` wait for everything a pipelined pb_step left running on the internal streams.
// ... and for the kernels of the resident-tracker path (pb_step_seq)
static int join_seq(pb_handle_st* h, cudaStream_t stream) {
    if (!h->seq_inflight) return PB_OK;
    for (int i = 0; i < 2; ++i)
        if (h->seq_half_used[i]) PB_CUDA(cudaStreamWaitEvent(stream, h->ev_seq_trk[i], 0));
    for (int i = 0; i < h->seq_lanes; ++i) PB_CUDA(cudaStreamWaitEvent(stream, h->ev_seq_nms[i], 0));
    h->seq_inflight = false;
    return PB_OK;
}

static int join_on(pb_handle_st* h, cudaStream_t stream) {
    PB_TRY(join_seq(h, stream));
    if (!h->inflight) return PB_OK;
    for (PipeSlot& sl : h->ring)
        if (sl.used) { PB_CUDA(cudaStreamWaitEvent(stream, sl.ev_nms, 0)); PB_CUDA(cudaStreamWaitEvent(stream, sl.ev_trk, 0)); }
    h->inflight = false;
    return PB_OK;
}

int pb_reset(pb_handle_t h, pb_stream_t stream) {
    if (!h) { pb_set_error("pb_reset: null handle"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    PB_TRY(join_on(h, (cudaStream_t)stream));
    PB_CUDA(launch_tracker_reset(h->trk, h->cfg.num_streams, h->cfg.max_tracks, h->cfg.max_detections, h->trk_seq, (cudaStream_t)stream));
    h->frames = 0;
    return PB_OK;
}

int pb_join(pb_handle_t h, pb_stream_t stream) {
    if (!h) { pb_set_error("pb_join: null handle"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    return join_on(h, (cudaStream_t)stream);
}

static int nms_sweep_mode(const pb_handle_st* h) { return h->lazy_keypoints ? 1 : (h->cfg.keypoint_fetch == 3 ? 2 : 0); }

int pb_postprocess(pb_handle_t h, const float* d_heads, float conf, float nms, pb_stream_t stream) {
    NvtxRange nvtx_range("pb_postprocess");
    if (!h || !d_heads) { pb_set_error("pb_postprocess: null argument"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    const pb_config& c = h->cfg;
    PB_TRY(join_on(h, (cudaStream_t)stream));
    int e0 = -1, e1 = -1, em = -1;
    if (h->profiling && (e0 = prof_event(h)) >= 0) cudaEventRecord(h->ev_pool[e0], (cudaStream_t)stream);
    PB_CUDA(launch_decode_gather(d_heads, c.num_streams, c.num_anchors, conf, h->lazy_keypoints, h->dplan, h->cand, (cudaStream_t)stream));
    if (h->profiling && e0 >= 0 && (em = prof_event(h)) >= 0) {
        cudaEventRecord(h->ev_pool[em], (cudaStream_t)stream);
        h->ev_gather.push_back({e0, em});
    }
    static const bool tier_exp = getenv("PB_NMS_TIER") != nullptr;          // experiment: the tiered (half-SM) kernel on the serial path
    if (tier_exp) {
        if (!h->exp_tier.ok) {
            h->exp_tier = tier_plan_by_id(c, atoi(getenv("PB_NMS_TIER")) > 0 ? atoi(getenv("PB_NMS_TIER")) : 1);
            if (h->exp_tier.ok && h->exp_tier.spill_stride) PB_TRY(dev_alloc(h, &h->exp_spill, (size_t)c.num_streams * h->exp_tier.spill_stride));
        }
        PB_CUDA(launch_nms_tier(h->exp_tier, d_heads, c.num_anchors, nms_sweep_mode(h), c.num_streams, c.max_candidates, c.max_keep, nms, h->dplan,
                                h->cand, h->post, h->exp_spill, (cudaStream_t)stream));
    } else if (h->pipe_tier.ok) {
        PB_CUDA(launch_nms_tier(h->pipe_tier, d_heads, c.num_anchors, nms_sweep_mode(h), c.num_streams, c.max_candidates, c.max_keep, nms, h->dplan,
                                h->cand, h->post, h->ring[h->cur].spill, (cudaStream_t)stream));
    } else
    PB_CUDA(launch_nms(d_heads, c.num_anchors, nms_sweep_mode(h), c.num_streams, c.max_candidates, c.max_keep, nms, h->dplan, h->cand, h->post, (cudaStream_t)stream));
    if (h->profiling && e0 >= 0 && (e1 = prof_event(h)) >= 0) {
        cudaEventRecord(h->ev_pool[e1], (cudaStream_t)stream);
        h->ev_post.push_back({e0, e1});
    }
    return PB_OK;
}

static TrackParams track_params(pb_handle_st* h, int frame_id) {
    const pb_config& c = h->cfg;
    TrackParams p{};
    p.seq = h->trk_seq + 1;                 // committed (h->trk_seq = p.seq) once the launch has succeeded: no gap on failure
    p.B = c.num_streams; p.T = c.max_tracks; p.Dm = c.max_detections;
    p.sub_solve_off = h->sub_solve_off;
    p.bulk_off = h->bulk_off;
    p.new_track_thresh = c.new_track_thresh; p.max_age = c.max_age; p.min_hits = c.min_hits;
    p.gating_enabled = c.gating_enabled; p.frame_id = frame_id;
    return p;
}

int pb_tracker_update(pb_handle_t h, const float* d_det_poses, const float* d_det_scores,
                      const int* d_num_dets, int det_stride, int frame_id, pb_stream_t stream) {
    NvtxRange nvtx_range("pb_tracker_update");
    if (!h) { pb_set_error("pb_tracker_update: null handle"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    const pb_config& c = h->cfg;
    PB_TRY(join_on(h, (cudaStream_t)stream));
    DetSource src;
    if (d_det_poses == nullptr && d_det_scores == nullptr && d_num_dets == nullptr) {
        src = {h->post.det_poses, h->post.det_scores, h->post.num_keep, c.max_keep};
    } else {
        if (!d_det_poses || !d_det_scores || !d_num_dets || det_stride < 1) {
            pb_set_error("pb_tracker_update: detection pointers must be all NULL or all set");
            return PB_ERR_INVALID;
        }
        src = {d_det_poses, d_det_scores, d_num_dets, det_stride};
    }
    TrackParams p = track_params(h, frame_id);
    int e0 = -1, e1 = -1;
    if (h->profiling && (e0 = prof_event(h)) >= 0) cudaEventRecord(h->ev_pool[e0], (cudaStream_t)stream);
    if (h->plan.pre_slices > 0) {            // large tables: row-sliced predict / gates / tier-1 costs ahead of the per-stream kernel
        PB_CUDA(launch_tracker_pre(h->trk, p, src, h->plan, (cudaStream_t)stream));
        p.precomputed = 1;
    }
    PB_CUDA(launch_tracker(h->trk, p, src, h->plan, (cudaStream_t)stream));
    h->trk_seq = p.seq;
    if (h->profiling && e0 >= 0 && (e1 = prof_event(h)) >= 0) {
        cudaEventRecord(h->ev_pool[e1], (cudaStream_t)stream);
        h->ev_track.push_back({e0, e1});
    }
    h->frames++;
    return PB_OK;
}

// Device-to-host copies of a step's TrackOutput records, enqueued right behind its tracker launch.
static int enqueue_readback(pb_handle_st* h, cudaStream_t stream) {
    if (!h->rb_tracks) return PB_OK;
    const size_t B = h->cfg.num_streams, Dm = h->cfg.max_detections;
    PB_CUDA(cudaMemcpyAsync(h->rb_counts, h->trk.num_outputs, B * sizeof(int), cudaMemcpyDeviceToHost, stream));
    PB_CUDA(cudaMemcpyAsync(h->rb_tracks, h->trk.outputs, B * Dm * 228, cudaMemcpyDeviceToHost, stream));
    return PB_OK;
}


// Fused step: decode+gather on the caller's stream (the only reader of the borrowed head tensor unless the sweep is lazy),
// then ONE launch per step — NMS and tracker of every stream-frame in one CTA (fused.cu) — on lane (step mod lanes).  The
// kernels of up to `lanes` steps are in flight; each video stream's frames stay in order through its chain word: a CTA
// whose stream is still busy with an earlier frame leaves its kept detections in the step's ring slot and exits, the CTA
// that holds the chain goes on with them.  Hence: every tracker stage of steps <= i has run when the kernels of ALL steps
// <= i have completed, which is what the events below express (ev_trk: the step's kernel; ev_nms: the step is complete
// for the host, including a read-back).  Ring slot = sequence number mod depth; the caller's stream waits for the slot's
// previous step — and, one step after the other, for every older one — before the decode kernel reuses its scratch.
static int step_fused(pb_handle_st* h, const float* d_heads, float conf, float nms, int frame_id, cudaStream_t stream) {
    const pb_config& c = h->cfg;
    const bool piped = c.pipeline_depth > 1;
    PB_TRY(join_seq(h, stream));            // (the lane's kernel follows the decode on the caller's stream)
    TrackParams tp = track_params(h, frame_id);
    const int depth = (int)h->ring.size();
    const int pos = tp.seq % depth;
    PipeSlot& sl = h->ring[pos];
    if (piped && sl.used) PB_CUDA(cudaStreamWaitEvent(stream, sl.ev_nms, 0));      // the slot's scratch, kept detections and records are free again
    cudaStream_t ts = stream;
    static const bool decode_on_lane = getenv("PB_DECODE_ON_LANE") != nullptr;      // experiment: relaxed borrow of the head tensor
    if (!(piped && decode_on_lane))
        PB_CUDA(launch_decode_gather(d_heads, c.num_streams, c.num_anchors, conf, h->lazy_keypoints, h->dplan, sl.cand, stream));
    if (piped) {
        ts = h->f_lane[tp.seq % h->f_lanes];
        PB_CUDA(cudaEventRecord(sl.ev_gather, stream));
        PB_CUDA(cudaStreamWaitEvent(ts, sl.ev_gather, 0));
        if (decode_on_lane)
            PB_CUDA(launch_decode_gather(d_heads, c.num_streams, c.num_anchors, conf, h->lazy_keypoints, h->dplan, sl.cand, ts));
    }
    sl.post.dbg_slot = tp.seq & 63;
    PB_CUDA(launch_fused(h->fplan, d_heads, c.num_anchors, nms_sweep_mode(h), c.num_streams, c.max_candidates, c.max_keep, nms, h->dplan,
                         sl.cand, sl.post, h->d_spill, h->trk, tp, h->rtab, ts));
    h->trk_seq = tp.seq;
    h->trk.outputs = sl.outputs; h->trk.num_outputs = sl.num_outputs;
    if (piped) {
        PB_CUDA(cudaEventRecord(sl.ev_trk, ts));
        cudaStream_t done_on = ts;
        if (h->rb_tracks) {
            // the step's records may have been written by the CTA of an older step's kernel on another lane
            for (int j = 0; j < h->f_lanes && j < depth; ++j) {
                PipeSlot& o = h->ring[((tp.seq - j) % depth + depth) % depth];
                if (o.used || j == 0) PB_CUDA(cudaStreamWaitEvent(h->s_rb, o.ev_trk, 0));
            }
            PB_TRY(enqueue_readback(h, h->s_rb));
            done_on = h->s_rb;
        }
        PB_CUDA(cudaEventRecord(sl.ev_nms, done_on));
        // a lazy sweep fetches keypoints from the borrowed head tensor inside the fused kernel: later work on the caller's
        // stream is ordered behind it (not needed when the buffer is the library's until pb_wait: pb_submit_host)
        if (h->lazy_keypoints && !h->borrow_until_wait) PB_CUDA(cudaStreamWaitEvent(stream, sl.ev_trk, 0));
        h->inflight = true;
    } else {
        PB_TRY(enqueue_readback(h, ts));
    }
    sl.used = true;
    h->cur = pos; h->post = sl.post; h->cand = sl.cand;
    h->frames++;
    return PB_OK;
}

// Pipelined step (pipeline_depth > 1): the three kernels of one step run on three streams —
// decode+gather on the caller's (it is the only reader of the borrowed head tensor), NMS and
// tracker on internal ones — chained by events, with `depth` slots of candidate scratch and
// kept-detection buffers.  Step i+1's decode+gather and NMS overlap step i's tracker; tracker
// launches stay in frame order on one stream (each stream's state depends on its previous
// frame).  Results are complete for the caller after pb_join or any pb_get_* call.
static int step_pipelined(pb_handle_st* h, const float* d_heads, float conf, float nms, int frame_id, cudaStream_t stream) {
    const pb_config& c = h->cfg;
    PB_TRY(join_seq(h, stream));            // (NMS and tracker follow the decode on the caller's stream)
    const int pos = h->inflight || h->ring[h->cur].used ? (h->cur + 1) % (int)h->ring.size() : h->cur;
    PipeSlot& sl = h->ring[pos];
    if (sl.used) PB_CUDA(cudaStreamWaitEvent(stream, sl.ev_nms, 0));          // scratch still being read
    PB_CUDA(launch_decode_gather(d_heads, c.num_streams, c.num_anchors, conf, h->lazy_keypoints, h->dplan, sl.cand, stream));
    PB_CUDA(cudaEventRecord(sl.ev_gather, stream));
    const TrackParams tp = track_params(h, frame_id);
    // Lanes: the NMS and the tracker launch of step i run back to back on internal stream i % lanes, so a
    // step's tracker follows its NMS without an event hop and `lanes` steps are in flight with one kernel
    // each (at most lanes * num_streams CTAs, one SM each: NMS launches of consecutive steps overlap
    // without starving the tracker grids).  Inside the tracker kernel the per-stream sequence flags keep
    // every video stream's frames in order (tracker.cu): the oldest grid never waits, so the spin cannot
    // deadlock while (lanes - 1) * num_streams < SM count.  lanes == 0: NMS launches on one stream,
    // tracker launches alternating between two.  With a read-back of the records behind every launch
    // (pb_submit_host) the output buffer of step i must not be overwritten before it is copied: one lane then.
    cudaStream_t lane[3] = {h->s_trk, h->s_trk2, h->s_trk3};
    const bool single = h->rb_tracks || !h->overlap_trackers || h->plan.pre_slices > 0;   // (the pre-kernel needs the previous frame's final state: stream order)
    cudaStream_t ts, ns;
    if (h->lanes > 0 && !single) { ts = lane[tp.seq % h->lanes]; ns = ts; }
    else {
        ts = (single || (tp.seq & 1)) ? h->s_trk : h->s_trk2;
        // NMS launches of consecutive steps are independent.  With tracker launches on one stream they alternate
        // between two streams (148 streams: 83 us per batch against 112); with two overlapping tracker grids a second
        // NMS grid in flight would starve them of SMs (64 streams: 39.4 us against 37.2), so one stream then.
        ns = (single && (tp.seq & 1)) ? h->s_trk2 : h->s_nms;
    }
    PB_CUDA(cudaStreamWaitEvent(ns, sl.ev_gather, 0));
    if (sl.used) PB_CUDA(cudaStreamWaitEvent(ns, sl.ev_trk, 0));              // kept detections still being read
    sl.post.dbg_slot = tp.seq & 63;
    if (h->pipe_tier.ok)
        PB_CUDA(launch_nms_tier(h->pipe_tier, d_heads, c.num_anchors, nms_sweep_mode(h), c.num_streams, c.max_candidates, c.max_keep, nms, h->dplan,
                                sl.cand, sl.post, sl.spill, ns));
    else
        PB_CUDA(launch_nms(d_heads, c.num_anchors, nms_sweep_mode(h), c.num_streams, c.max_candidates, c.max_keep, nms, h->dplan, sl.cand, sl.post, ns));
    PB_CUDA(cudaEventRecord(sl.ev_nms, ns));
    // the lazy NMS sweep fetches keypoints from the borrowed head tensor: later work on the caller's
    // stream (e.g. the engine writing the next batch into the same buffer) is ordered after it then.
    // With complete records the decode+gather kernel is the only reader of the head and nothing is needed.
    if (h->lazy_keypoints) PB_CUDA(cudaStreamWaitEvent(stream, sl.ev_nms, 0));
    if ((h->rb_tracks || h->last_was_readback) && h->last_trk_stream && h->last_trk_stream != ts)
        PB_CUDA(cudaStreamWaitEvent(ts, h->ring[h->cur].ev_trk, 0));   // the previous step's records are still being copied out
    if (ns != ts) PB_CUDA(cudaStreamWaitEvent(ts, sl.ev_nms, 0));
    DetSource src{sl.post.det_poses, sl.post.det_scores, sl.post.num_keep, c.max_keep};
    TrackParams tpp = tp;
    if (h->plan.pre_slices > 0) { PB_CUDA(launch_tracker_pre(h->trk, tpp, src, h->plan, ts)); tpp.precomputed = 1; }
    PB_CUDA(launch_tracker(h->trk, tpp, src, h->plan, ts));
    h->trk_seq = tp.seq;
    PB_TRY(enqueue_readback(h, ts));
    PB_CUDA(cudaEventRecord(sl.ev_trk, ts));
    h->last_trk_stream = ts;
    h->last_was_readback = h->rb_tracks != nullptr;
    sl.used = true;
    h->cur = pos; h->post = sl.post; h->cand = sl.cand;
    h->inflight = true;
    h->frames++;
    return PB_OK;
}

int pb_step(pb_handle_t h, const float* d_heads, float conf, float nms, int frame_id, pb_stream_t stream) {
    NvtxRange nvtx_range("pb_step");
    if (!h || !d_heads) { pb_set_error("pb_step: null argument"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    if (h->fplan.ok && !h->profiling) return step_fused(h, d_heads, conf, nms, frame_id, (cudaStream_t)stream);
    if (h->cfg.pipeline_depth > 1 && !h->profiling) return step_pipelined(h, d_heads, conf, nms, frame_id, (cudaStream_t)stream);
    PB_TRY(pb_postprocess(h, d_heads, conf, nms, stream));
    PB_TRY(pb_tracker_update(h, nullptr, nullptr, nullptr, 0, frame_id, stream));
    return enqueue_readback(h, (cudaStream_t)stream);
}

// ---- resident-tracker path of pb_step_seq ------------------------------------------------------------------------------
// A sequence of steps is cut into chunks of at most seq_chunk steps.  Per chunk: ONE tracker launch (one CTA per video stream,
// resident for the whole chunk, the stream's state in shared memory from the first frame to the last) on its own stream,
// enqueued FIRST, then per step the decode+gather and NMS kernels back to back on internal lane (step mod lanes).  The tracker
// CTA of a stream runs frame i as soon as the NMS kernel of step i has published that stream's kept detections in the step's
// slot (a per-stream flag, release / acquire): no launch, no hand-over between CTAs and no event between the frames of a
// stream, and no stream waits for another.  The decode and NMS kernels depend on nothing the tracker kernel does (their slots
// belong to the half of the ring the previous chunk did not use), so its waits cannot deadlock as long as its CTAs leave SMs
// free (seq_configure); the kernels it waits for are loaded before the first launch (preload_post_kernels).  Chunks alternate
// between the two halves of the ring; a half is reused when the tracker launch that read it has completed (event).  The
// borrowed head tensors are read by the decode kernels only: the caller's stream is ordered behind the last of them.
static int alloc_slot(pb_handle_st* h, PipeSlot& sl, unsigned long long* post_ns) {
    const pb_config& c = h->cfg;
    const size_t B = c.num_streams, Dm = c.max_detections, K = c.max_keep;
    const size_t nsc = B * (size_t)h->dplan.nseg * (size_t)h->dplan.segcap;
    PB_TRY(dev_alloc(h, &sl.post.det_poses, B * K * POSE_F));
    PB_TRY(dev_alloc(h, &sl.post.det_bboxes, B * K * 4));
    PB_TRY(dev_alloc(h, &sl.post.det_scores, B * K));
    PB_TRY(dev_alloc(h, &sl.post.keep_slots, B * K));
    PB_TRY(dev_alloc(h, &sl.post.keep_anchors, B * K));
    PB_TRY(dev_alloc(h, &sl.post.num_keep, B));
    PB_TRY(dev_alloc(h, &sl.post.num_cand, B));
    sl.post.stage_ns = post_ns;
    PB_TRY(dev_alloc(h, &sl.cand.records, nsc * HEAD_ROWS));
    PB_TRY(dev_alloc(h, &sl.cand.anchors, nsc));
    PB_TRY(dev_alloc(h, &sl.cand.counts, B * (size_t)h->dplan.nseg));
    unsigned char* outp = nullptr;
    PB_TRY(dev_alloc(h, &outp, B * Dm * 228));
    sl.outputs = outp;
    PB_TRY(dev_alloc(h, &sl.num_outputs, B));
    PB_TRY(dev_alloc(h, &sl.ready, B));
    // (the spill scratch of the tiered NMS kernel belongs to the lane, not to the slot: the NMS launches of a lane run one after the other)
    PB_CUDA(cudaEventCreateWithFlags(&sl.ev_gather, cudaEventDisableTiming));
    PB_CUDA(cudaEventCreateWithFlags(&sl.ev_nms, cudaEventDisableTiming));
    PB_CUDA(cudaEventCreateWithFlags(&sl.ev_trk, cudaEventDisableTiming));
    return PB_OK;
}

// Can this handle use the resident path, and with which tracker plan?  Small tables only (the auction runs in one warp, the
// frame is a chain of short stages: exactly what a resident CTA is good at; large tables keep the row-sliced pre-kernel), and
// the tracker CTAs must leave at least half of the SMs to the decode and NMS kernels they wait for.
// smallest shared-memory configuration (KB) of an sm_100 SM that holds a CTA with `bytes` of dynamic shared memory (+ 1 KB reserved)
static int smem_config_kb(size_t bytes) {
    static const int cls[] = {8, 16, 32, 64, 100, 132, 164, 196, 228};
    for (int k : cls) if (bytes + 1024 <= (size_t)k * 1024) return k;
    return 228;
}

// Tiered NMS kernel variants by number: 1 = 2 x 512 threads per SM, 2 = 3 x 384, 3 = 3 x 256, 4 = 4 x 256 (0: none)
static NmsTierPlan tier_plan_by_id(const pb_config& c, int id) {
    // id >= 32: 1024 threads, one CTA per SM, id KB of shared memory (the working set of fewer candidates than the cap; more spill)
    static const int threads[5] = {0, 512, 384, 256, 256}, per_sm[5] = {0, 2, 3, 3, 4};
    // 1000 + kb: two 512-thread CTAs per SM with kb KB of shared memory each (81: the pair fills the 164 KB configuration)
    if (id >= 1032 && id <= 1113) {
        int dev2 = 0, optin2 = 0;
        cudaGetDevice(&dev2);
        cudaDeviceGetAttribute(&optin2, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev2);
        return nms_tier_plan(c.max_candidates, c.max_keep, (size_t)optin2, 512, 2, id - 1000);
    }
    if (id < 1 || (id > 4 && id < 32) || id > 227) return NmsTierPlan{};
    int dev = 0, optin = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (id >= 32) return nms_tier_plan(c.max_candidates, c.max_keep, (size_t)optin, 1024, 1, id);
    return nms_tier_plan(c.max_candidates, c.max_keep, (size_t)optin, threads[id], per_sm[id]);
}

static void seq_configure(pb_handle_st* h, int sm_count) {
    const pb_config& c = h->cfg;
    h->seq_chunk = 0;
    if (c.pipeline_depth < 2 || h->plan.pre_slices > 0) return;
    if (const char* e = getenv("PB_SEQ")) { if (atoi(e) == 0) return; }
    bool resident = true, compact = false;
    if (const char* e = getenv("PB_SEQ_RESIDENT_STATE")) resident = atoi(e) != 0;
    if (const char* e = getenv("PB_SEQ_COMPACT")) compact = atoi(e) != 0;
    TrackerPlan p = tracker_plan(c.max_tracks, c.max_detections, compact, resident);
    if (compact) p.threads = 512;
    if (const char* e = getenv("PB_SEQ_THREADS")) { const int t = atoi(e); if (t == 256 || t == 512 || t == 1024) p.threads = t; }
    if (!(p.cost_in_smem && p.det_in_smem && p.pred_in_smem)) return;
    if ((long)c.max_tracks * c.max_detections > 16384) return;
    // One tracker CTA per SM (1024 threads), resident for a whole chunk: the decode and NMS kernels that feed them get the other
    // SMs.  What those SMs must deliver is one step's NMS results per tracker frame (18-21 us): with one 1024-thread NMS CTA per
    // SM (33-35 us each) the 84 SMs left beside 64 streams cannot (26.7-39 us per step; the per-step path: 26.8 us).  The steps of
    // this path therefore launch the NMS kernel as pairs of 512-thread CTAs with 81 KB of shared memory each — the working set of
    // 256 candidates, the spill path beyond — so that a pair fills the SM's 164 KB configuration: no SM re-splits L1 / shared
    // memory between the kernels that alternate on it (pairs of 113 KB, which need the 228 KB configuration: 31-32 us per step).
    // Measured (tools/seq_probe.py, 640x640 heads, 20 persons, four lanes): 64 streams 23.3 us per step against 26.8 us for the
    // per-step path (20-step bursts 26.2-26.7 against 30.2), 72 streams 25.8 against 29.0, 48 streams 22.5 against 25.5, 32
    // streams 22.0 against 25.1; the tracker CTAs wait 0.2-0.5 us per frame.  Three or five lanes: 24.4 us at 64 streams.
    // The path needs half of the SMs free for the kernels the tracker CTAs wait for (no deadlock): 2 * streams <= SM count.
    if (2 * c.num_streams > sm_count) return;
    h->seq_plan = p;
    h->seq_nms = tier_plan_by_id(c, 1081);
    h->seq_lanes = 4;
    if (const char* e = getenv("PB_SEQ_NMS_TIER")) h->seq_nms = tier_plan_by_id(c, atoi(e));
    if (const char* e = getenv("PB_SEQ_LANES")) { const int v = atoi(e); if (v >= 1 && v <= pb_handle_st::SEQ_MAX_LANES) h->seq_lanes = v; }
    int chunk = PB_SEQ_MAX;
    if (const char* e = getenv("PB_SEQ_CHUNK")) { const int v = atoi(e); if (v >= 1 && v <= PB_SEQ_MAX) chunk = v; }
    h->seq_chunk = chunk;
}

static int seq_prepare(pb_handle_st* h) {
    if (!h->seq_ring.empty()) return PB_OK;
    PB_CUDA(preload_post_kernels(h->cfg.max_candidates, h->cfg.max_keep, &h->seq_nms));
    std::vector<PipeSlot> ring((size_t)2 * h->seq_chunk);
    for (PipeSlot& sl : ring) PB_TRY(alloc_slot(h, sl, h->ring[0].post.stage_ns));
    for (int i = 0; i < 2; ++i) PB_CUDA(cudaEventCreateWithFlags(&h->ev_seq_trk[i], cudaEventDisableTiming));
    for (int i = 0; i < h->seq_lanes; ++i) {
        if (h->seq_nms.ok && h->seq_nms.spill_stride)
            PB_TRY(dev_alloc(h, &h->seq_spill[i], (size_t)h->cfg.num_streams * h->seq_nms.spill_stride));
        PB_CUDA(cudaStreamCreateWithFlags(&h->s_seq_nms[i], cudaStreamNonBlocking));
        PB_CUDA(cudaEventCreateWithFlags(&h->ev_seq_nms[i], cudaEventDisableTiming));
        PB_CUDA(cudaEventCreateWithFlags(&h->ev_seq_dec[i], cudaEventDisableTiming));
        PB_CUDA(cudaEventRecord(h->ev_seq_nms[i], h->s_seq_nms[i]));
    }
    PB_CUDA(cudaStreamCreateWithFlags(&h->s_seq_trk, cudaStreamNonBlocking));
    PB_CUDA(cudaEventCreateWithFlags(&h->ev_seq_start, cudaEventDisableTiming));
    h->seq_ring.swap(ring);
    return PB_OK;
}

static int step_seq_resident(pb_handle_st* h, const float* d_heads, size_t step_stride, int period, int first, int n_steps,
                             float conf, float nms, int frame0, cudaStream_t stream) {
    const pb_config& c = h->cfg;
    PB_TRY(seq_prepare(h));
    if (h->inflight) {
        // steps of the other paths may still run on their internal streams: everything of this call is ordered behind them
        for (PipeSlot& sl : h->ring)
            if (sl.used) { PB_CUDA(cudaStreamWaitEvent(stream, sl.ev_nms, 0)); PB_CUDA(cudaStreamWaitEvent(stream, sl.ev_trk, 0)); }
        h->inflight = false;
    }
    // the internal streams follow whatever the caller's stream holds now (and, through it, the steps joined above)
    PB_CUDA(cudaEventRecord(h->ev_seq_start, stream));
    PB_CUDA(cudaStreamWaitEvent(h->s_seq_trk, h->ev_seq_start, 0));
    const int L = h->seq_lanes;
    for (int l = 0; l < L; ++l) PB_CUDA(cudaStreamWaitEvent(h->s_seq_nms[l], h->ev_seq_start, 0));
    const int M = h->seq_chunk;
    for (int s0 = 0; s0 < n_steps; s0 += M) {
        const int n = (n_steps - s0 < M) ? (n_steps - s0) : M;
        const int half = h->seq_half;
        TrackParams tp = track_params(h, frame0 + s0);
        SeqTable q{};
        q.n = n; q.stride = c.max_keep;
        for (int i = 0; i < n; ++i) {
            PipeSlot& sl = h->seq_ring[(size_t)half * M + i];
            q.det_poses[i] = sl.post.det_poses; q.det_scores[i] = sl.post.det_scores; q.num_keep[i] = sl.post.num_keep;
            q.ready[i] = sl.ready; q.outputs[i] = sl.outputs; q.num_outputs[i] = sl.num_outputs;
        }
        // the slots of this half are free once the tracker launch that read them has completed: every lane waits for it
        // before its first kernel of this chunk
        const bool wait_half = h->seq_half_used[half];
        if (wait_half)
            for (int l = 0; l < L && l < n; ++l) PB_CUDA(cudaStreamWaitEvent(h->s_seq_nms[(tp.seq + l) % L], h->ev_seq_trk[half], 0));
        // The tracker launch first: its CTAs are resident (and waiting for frame 0) while the host enqueues the steps.  Where
        // kernels cannot overlap — a profiler that replays kernels one at a time (ncu), CUDA_LAUNCH_BLOCKING=1 — a kernel must
        // not wait for one launched after it: PB_SEQ_TRACKER_LAST=1 (or CUDA_LAUNCH_BLOCKING) enqueues it behind the chunk's steps.
        // A tool injected into the process (Nsight Compute sets CUDA_INJECTION64_PATH / NV_COMPUTE_PROFILER_* for its target) is
        // taken for a serialising one: the order costs some overlap and nothing else.
        static const bool tracker_last = (getenv("PB_SEQ_TRACKER_LAST") ? atoi(getenv("PB_SEQ_TRACKER_LAST")) != 0
                                                                         : (getenv("CUDA_INJECTION64_PATH") != nullptr || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") != nullptr ||
                                                                            getenv("NV_NSIGHT_COMPUTE_INJECTION") != nullptr)) ||
                                         (getenv("CUDA_LAUNCH_BLOCKING") && atoi(getenv("CUDA_LAUNCH_BLOCKING")) != 0);
        if (!tracker_last) {
            PB_CUDA(launch_tracker_seq(h->trk, tp, q, h->seq_plan, h->s_seq_trk));
            PB_CUDA(cudaEventRecord(h->ev_seq_trk[half], h->s_seq_trk));
        }
        h->seq_inflight = true;
        for (int i = 0; i < n; ++i) {
            PipeSlot& sl = h->seq_ring[(size_t)half * M + i];
            const int seq = tp.seq + i;
            const int lane = seq % L;
            cudaStream_t ns = h->s_seq_nms[lane];
            const float* heads = d_heads + (size_t)((first + s0 + i) % period) * step_stride;
            PB_CUDA(launch_decode_gather(heads, c.num_streams, c.num_anchors, conf, false, h->dplan, sl.cand, ns));
            if (s0 + i >= n_steps - L) PB_CUDA(cudaEventRecord(h->ev_seq_dec[lane], ns));
            sl.post.ready = sl.ready; sl.post.ready_seq = seq;
            sl.post.dbg_slot = seq & 63;
            if (h->seq_nms.ok)
                PB_CUDA(launch_nms_tier(h->seq_nms, heads, c.num_anchors, 0, c.num_streams, c.max_candidates, c.max_keep, nms, h->dplan, sl.cand,
                                        sl.post, h->seq_spill[lane], ns));
            else
                PB_CUDA(launch_nms(heads, c.num_anchors, 0, c.num_streams, c.max_candidates, c.max_keep, nms, h->dplan, sl.cand, sl.post, ns));
            if (s0 + i >= n_steps - L) PB_CUDA(cudaEventRecord(h->ev_seq_nms[lane], ns));
            if (i == n - 1) {
                h->post = sl.post; h->cand = sl.cand;
                h->trk.outputs = sl.outputs; h->trk.num_outputs = sl.num_outputs;
            }
        }
        if (tracker_last) {
            // (serialising tools run the steps' kernels to completion first; otherwise the lanes' work is simply ahead)
            for (int l = 0; l < L && l < n; ++l) {
                PB_CUDA(cudaEventRecord(h->ev_seq_nms[(tp.seq + n - 1 - l) % L], h->s_seq_nms[(tp.seq + n - 1 - l) % L]));
                PB_CUDA(cudaStreamWaitEvent(h->s_seq_trk, h->ev_seq_nms[(tp.seq + n - 1 - l) % L], 0));
            }
            PB_CUDA(launch_tracker_seq(h->trk, tp, q, h->seq_plan, h->s_seq_trk));
            PB_CUDA(cudaEventRecord(h->ev_seq_trk[half], h->s_seq_trk));
        }
        h->seq_half_used[half] = true;
        h->trk_seq = tp.seq + n - 1;
        h->frames += n;
        h->seq_half ^= 1;
    }
    // the borrowed head tensors: later work on the caller's stream is ordered behind their last readers (the decode kernels)
    for (int l = 0; l < L && l < n_steps; ++l) PB_CUDA(cudaStreamWaitEvent(stream, h->ev_seq_dec[(h->trk_seq - l) % L], 0));
    return PB_OK;
}

int pb_step_seq(pb_handle_t h, const float* d_heads, size_t step_stride, int period, int first, int n_steps,
                float conf, float nms, int frame0, pb_stream_t stream) {
    NvtxRange nvtx_range("pb_step_seq");
    if (!h || !d_heads || period < 1 || first < 0 || n_steps < 0) { pb_set_error("pb_step_seq: bad argument"); return PB_ERR_INVALID; }
    if (h->seq_chunk > 0 && n_steps >= 2 && !h->profiling && !h->lazy_keypoints && !h->rb_tracks) {
        DevGuard dev_guard(h->cfg.device);
        return step_seq_resident(h, d_heads, step_stride, period, first, n_steps, conf, nms, frame0, (cudaStream_t)stream);
    }
    for (int i = 0; i < n_steps; ++i)
        PB_TRY(pb_step(h, d_heads + (size_t)((first + i) % period) * step_stride, conf, nms, frame0 + i, stream));
    return PB_OK;
}

int pb_step_path(pb_handle_t h, int* per_step, int* seq_chunk) {
    if (!h) { pb_set_error("pb_step_path: null handle"); return PB_ERR_INVALID; }
    if (per_step) *per_step = h->fplan.ok ? 2 : (h->cfg.pipeline_depth > 1 ? 1 : 0);
    if (seq_chunk) *seq_chunk = h->seq_chunk;
    return PB_OK;
}

int pb_nms_plan(pb_handle_t h, int* threads, int* ctas_per_sm, int* smem_bytes, int* tier_candidates) {
    if (!h) { pb_set_error("pb_nms_plan: null handle"); return PB_ERR_INVALID; }
    const bool tier = h->pipe_tier.ok && !h->fplan.ok;
    if (threads) *threads = h->fplan.ok ? h->fplan.threads : (tier ? h->pipe_tier.threads : 1024);
    if (ctas_per_sm) *ctas_per_sm = h->fplan.ok ? (h->fplan.threads <= 512 ? 2 : 1) : (tier ? h->pipe_tier.per_sm : 1);
    if (smem_bytes) *smem_bytes = (int)(h->fplan.ok ? h->fplan.smem_bytes : (tier ? h->pipe_tier.smem_bytes : decode_nms_smem_bytes(h->cfg.max_candidates, h->cfg.max_keep)));
    if (tier_candidates) *tier_candidates = h->fplan.ok ? h->fplan.CT : (tier ? h->pipe_tier.CT : h->cfg.max_candidates);
    return PB_OK;
}

int pb_step_host(pb_handle_t h, const float* h_heads, float conf, float nms, int frame_id,
                 void* h_tracks, int* h_counts) {
    NvtxRange nvtx_range("pb_step_host");
    if (!h || !h_heads || !h_tracks || !h_counts) { pb_set_error("pb_step_host: null argument"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    const pb_config& c = h->cfg;
    const size_t B = c.num_streams, Dm = c.max_detections;
    const size_t head_bytes = B * HEAD_ROWS * (size_t)c.num_anchors * sizeof(float);
    cudaStream_t s = h->own_stream;
    // Page-locked (cudaHostAlloc / cudaHostRegister) input is read IN PLACE by the decode kernel
    // over PCIe: the confidence rows and the 32-byte sectors at the candidate anchors — about a
    // seventh of the tensor at 130 candidates per stream — instead of staging all of it in HBM.
    // Pageable input cannot be mapped and is staged with one copy.
    const float* src = nullptr;
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, h_heads) == cudaSuccess && at.devicePointer != nullptr &&
        (at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged)) {
        src = static_cast<const float*>(at.devicePointer);
    } else {
        (void)cudaGetLastError();
        if (!h->d_stage) PB_CUDA(cudaMalloc(&h->d_stage, head_bytes));
        PB_CUDA(cudaMemcpyAsync(h->d_stage, h_heads, head_bytes, cudaMemcpyHostToDevice, s));
        src = h->d_stage;
    }
    const bool lazy_before = h->lazy_keypoints;
    if (c.keypoint_fetch == 0) h->lazy_keypoints = (src != h->d_stage) && at.type == cudaMemoryTypeHost;
    const int step_rc = pb_step(h, src, conf, nms, frame_id, (pb_stream_t)s);
    h->lazy_keypoints = lazy_before;
    PB_TRY(step_rc);
    PB_TRY(join_on(h, s));
    // results: straight into the caller's buffers when they are page-locked, else via pinned staging
    cudaPointerAttributes ot{}, oc{};
    const bool direct = cudaPointerGetAttributes(&ot, h_tracks) == cudaSuccess && ot.type == cudaMemoryTypeHost &&
                        cudaPointerGetAttributes(&oc, h_counts) == cudaSuccess && oc.type == cudaMemoryTypeHost;
    (void)cudaGetLastError();
    if (direct) {
        PB_CUDA(cudaMemcpyAsync(h_counts, h->trk.num_outputs, B * sizeof(int), cudaMemcpyDeviceToHost, s));
        PB_CUDA(cudaMemcpyAsync(h_tracks, h->trk.outputs, B * Dm * 228, cudaMemcpyDeviceToHost, s));
        PB_CUDA(cudaStreamSynchronize(s));
    } else {
        if (!h->h_out_pinned) PB_CUDA(cudaMallocHost(&h->h_out_pinned, B * Dm * 228));
        if (!h->h_cnt_pinned) PB_CUDA(cudaMallocHost(&h->h_cnt_pinned, B * sizeof(int)));
        PB_CUDA(cudaMemcpyAsync(h->h_cnt_pinned, h->trk.num_outputs, B * sizeof(int), cudaMemcpyDeviceToHost, s));
        PB_CUDA(cudaMemcpyAsync(h->h_out_pinned, h->trk.outputs, B * Dm * 228, cudaMemcpyDeviceToHost, s));
        PB_CUDA(cudaStreamSynchronize(s));
        memcpy(h_counts, h->h_cnt_pinned, B * sizeof(int));
        memcpy(h_tracks, h->h_out_pinned, B * Dm * 228);
    }
    return check_order_flag(h);
}

// ---- SURVEY.md §8f rows f2 / f4 ---------------------------------------------------------------

int pb_set_output_transform(pb_handle_t h, const float* h_xform) {
    if (!h) { pb_set_error("pb_set_output_transform: null handle"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    PB_CUDA(cudaDeviceSynchronize());
    if (!h_xform) { h->trk.out_xform = nullptr; return PB_OK; }
    const size_t n = (size_t)h->cfg.num_streams * 4;
    if (!h->d_xform) PB_TRY(dev_alloc(h, &h->d_xform, n));
    PB_CUDA(cudaMemcpy(h->d_xform, h_xform, n * sizeof(float), cudaMemcpyHostToDevice));
    h->trk.out_xform = h->d_xform;
    return PB_OK;
}

int pb_state_size(pb_handle_t h, size_t* bytes) {
    if (!h || !bytes) { pb_set_error("pb_state_size: bad argument"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    size_t total = sizeof(SnapHeader);
    PB_TRY(for_each_state_slab(h, [&](void*, size_t n) -> int { total += n; return PB_OK; }));
    *bytes = total;
    return PB_OK;
}

int pb_state_save(pb_handle_t h, void* h_blob, size_t capacity) {
    size_t need = 0;
    PB_TRY(pb_state_size(h, &need));
    DevGuard dev_guard(h->cfg.device);
    if (!h_blob || capacity < need) { pb_set_error("pb_state_save: blob too small (%zu < %zu)", capacity, need); return PB_ERR_INVALID; }
    PB_CUDA(cudaDeviceSynchronize());
    PB_TRY(check_order_flag(h));
    SnapHeader hd{kSnapMagic, 1u, h->cfg.num_streams, h->cfg.max_tracks, h->cfg.max_detections, h->frames};
    unsigned char* p = static_cast<unsigned char*>(h_blob);
    memcpy(p, &hd, sizeof(hd)); p += sizeof(hd);
    return for_each_state_slab(h, [&](void* d, size_t n) -> int {
        PB_CUDA(cudaMemcpy(p, d, n, cudaMemcpyDeviceToHost));
        p += n;
        return PB_OK;
    });
}

int pb_state_load(pb_handle_t h, const void* h_blob, size_t bytes) {
    size_t need = 0;
    PB_TRY(pb_state_size(h, &need));
    DevGuard dev_guard(h->cfg.device);
    if (!h_blob || bytes < need) { pb_set_error("pb_state_load: blob too small"); return PB_ERR_INVALID; }
    SnapHeader hd{};
    memcpy(&hd, h_blob, sizeof(hd));
    if (hd.magic != kSnapMagic || hd.version != 1u || hd.B != h->cfg.num_streams || hd.T != h->cfg.max_tracks || hd.Dm != h->cfg.max_detections) {
        pb_set_error("pb_state_load: snapshot is for %d streams x %d tracks x %d detections, handle has %d x %d x %d", hd.B, hd.T, hd.Dm,
                     h->cfg.num_streams, h->cfg.max_tracks, h->cfg.max_detections);
        return PB_ERR_INVALID;
    }
    PB_CUDA(cudaDeviceSynchronize());
    const unsigned char* p = static_cast<const unsigned char*>(h_blob) + sizeof(hd);
    h->frames = hd.frames;
    return for_each_state_slab(h, [&](void* d, size_t n) -> int {
        PB_CUDA(cudaMemcpy(d, p, n, cudaMemcpyHostToDevice));
        p += n;
        return PB_OK;
    });
}

int pb_submit_host(pb_handle_t h, const float* h_heads, float conf, float nms, int frame_id,
                   void* h_tracks, int* h_counts) {
    NvtxRange nvtx_range("pb_submit_host");
    if (!h || !h_heads || !h_tracks || !h_counts) { pb_set_error("pb_submit_host: null argument"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    const pb_config& c = h->cfg;
    cudaPointerAttributes at{}, ot{}, oc{};
    const bool ok = cudaPointerGetAttributes(&at, h_heads) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer &&
                    cudaPointerGetAttributes(&ot, h_tracks) == cudaSuccess && ot.type == cudaMemoryTypeHost &&
                    cudaPointerGetAttributes(&oc, h_counts) == cudaSuccess && oc.type == cudaMemoryTypeHost;
    (void)cudaGetLastError();
    if (!ok) { pb_set_error("pb_submit_host: heads, tracks and counts must be page-locked host memory (cudaHostAlloc / cudaHostRegister)"); return PB_ERR_INVALID; }
    const bool lazy_before = h->lazy_keypoints;
    if (c.keypoint_fetch == 0) h->lazy_keypoints = true;
    h->rb_tracks = h_tracks; h->rb_counts = h_counts;
    h->borrow_until_wait = true;
    const int rc = pb_step(h, static_cast<const float*>(at.devicePointer), conf, nms, frame_id, (pb_stream_t)h->own_stream);
    h->borrow_until_wait = false;
    h->rb_tracks = nullptr; h->rb_counts = nullptr;
    h->lazy_keypoints = lazy_before;
    return rc;
}

// After a synchronisation: did a tracker CTA give up waiting for its predecessor (tracker.cu)?
static int check_order_flag(pb_handle_st* h) {
    int f = 0;
    PB_CUDA(cudaMemcpy(&f, h->trk.error_flag, sizeof(int), cudaMemcpyDeviceToHost));
    if (f) { pb_set_error("tracker: a stream's previous frame did not complete within the time-out; results are invalid"); return PB_ERR_CUDA; }
    return PB_OK;
}

int pb_wait(pb_handle_t h) {
    NvtxRange nvtx_range("pb_wait");
    if (!h) { pb_set_error("pb_wait: null handle"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    PB_TRY(join_on(h, h->own_stream));
    PB_CUDA(cudaStreamSynchronize(h->own_stream));
    return check_order_flag(h);
}

int pb_get_tracks(pb_handle_t h, int b, void* out, int cap, int* n_out) {
    if (!h || b < 0 || b >= h->cfg.num_streams || !n_out) { pb_set_error("pb_get_tracks: bad argument"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    PB_CUDA(cudaDeviceSynchronize());
    PB_TRY(check_order_flag(h));
    int n = 0;
    PB_CUDA(cudaMemcpy(&n, h->trk.num_outputs + b, sizeof(int), cudaMemcpyDeviceToHost));
    if (n > cap) n = cap;
    if (n > 0 && out)
        PB_CUDA(cudaMemcpy(out, static_cast<unsigned char*>(h->trk.outputs) + (size_t)b * h->cfg.max_detections * 228,
                           (size_t)n * 228, cudaMemcpyDeviceToHost));
    *n_out = n;
    return PB_OK;
}

int pb_get_tracks_all(pb_handle_t h, void* out, int* counts) {
    if (!h || !out || !counts) { pb_set_error("pb_get_tracks_all: bad argument"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    PB_CUDA(cudaDeviceSynchronize());
    PB_TRY(check_order_flag(h));
    const size_t B = h->cfg.num_streams;
    PB_CUDA(cudaMemcpy(counts, h->trk.num_outputs, B * sizeof(int), cudaMemcpyDeviceToHost));
    PB_CUDA(cudaMemcpy(out, h->trk.outputs, B * h->cfg.max_detections * 228, cudaMemcpyDeviceToHost));
    return PB_OK;
}

int pb_get_num_active(pb_handle_t h, int* out) {
    if (!h || !out) { pb_set_error("pb_get_num_active: bad argument"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    PB_CUDA(cudaDeviceSynchronize());
    PB_TRY(check_order_flag(h));
    std::vector<int> sc((size_t)h->cfg.num_streams * 4);
    PB_CUDA(cudaMemcpy(sc.data(), h->trk.scalars, sc.size() * sizeof(int), cudaMemcpyDeviceToHost));
    for (int b = 0; b < h->cfg.num_streams; ++b) out[b] = sc[(size_t)b * 4 + 3];
    return PB_OK;
}

#define PB_D2H(dst, src, count)                                                                     \
    do { if (dst) PB_CUDA(cudaMemcpy(dst, src, (size_t)(count) * sizeof(*(dst)), cudaMemcpyDeviceToHost)); } while (0)

int pb_get_kept(pb_handle_t h, int b, float* poses, float* bboxes, float* scores, int* keep_slots,
                int* keep_anchors, int cap, int* num_keep, int* num_cand) {
    if (!h || b < 0 || b >= h->cfg.num_streams) { pb_set_error("pb_get_kept: bad argument"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    PB_CUDA(cudaDeviceSynchronize());
    const size_t K = h->cfg.max_keep;
    int nk = 0, nc = 0;
    PB_CUDA(cudaMemcpy(&nk, h->post.num_keep + b, sizeof(int), cudaMemcpyDeviceToHost));
    PB_CUDA(cudaMemcpy(&nc, h->post.num_cand + b, sizeof(int), cudaMemcpyDeviceToHost));
    if (num_keep) *num_keep = nk;
    if (num_cand) *num_cand = nc;
    int n = nk < cap ? nk : cap;
    if (n > 0) {
        PB_D2H(poses, h->post.det_poses + b * K * POSE_F, (size_t)n * POSE_F);
        PB_D2H(bboxes, h->post.det_bboxes + b * K * 4, (size_t)n * 4);
        PB_D2H(scores, h->post.det_scores + b * K, n);
        PB_D2H(keep_slots, h->post.keep_slots + b * K, n);
        PB_D2H(keep_anchors, h->post.keep_anchors + b * K, n);
    }
    return PB_OK;
}

int pb_get_state(pb_handle_t h, int b, float* poses, float* vel, float* scores, int* states, int* ids,
                 int* hits, int* ages, int* last_frame, int* active, int* row_assign, int* col_assign,
                 float* cost, float* predicted, float* centers, int* scalars) {
    if (!h || b < 0 || b >= h->cfg.num_streams) { pb_set_error("pb_get_state: bad argument"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    PB_CUDA(cudaDeviceSynchronize());
    PB_TRY(check_order_flag(h));
    const size_t T = h->cfg.max_tracks, Dm = h->cfg.max_detections;
    const TrackBuffers& t = h->trk;
    PB_D2H(poses, t.poses + b * T * POSE_F, T * POSE_F);
    PB_D2H(vel, t.vel + b * T * 34, T * 34);
    PB_D2H(scores, t.scores + b * T, T);
    PB_D2H(states, t.states + b * T, T);
    PB_D2H(ids, t.ids + b * T, T);
    PB_D2H(hits, t.hits + b * T, T);
    PB_D2H(ages, t.ages + b * T, T);
    PB_D2H(last_frame, t.last_frame + b * T, T);
    PB_D2H(active, t.active + b * T, T);
    PB_D2H(row_assign, t.row_assign + b * T, T);
    PB_D2H(col_assign, t.col_assign + b * Dm, Dm);
    PB_D2H(cost, t.cost + b * T * Dm, T * Dm);
    PB_D2H(predicted, t.predicted + b * T * POSE_F, T * POSE_F);
    PB_D2H(centers, t.tcent + b * T * 4, T * 4);
    PB_D2H(scalars, t.scalars + b * 4, 4);
    return PB_OK;
}

int pb_get_device_views(pb_handle_t h, pb_device_views* v) {
    if (!h || !v) { pb_set_error("pb_get_device_views: bad argument"); return PB_ERR_INVALID; }
    v->det_poses = h->post.det_poses; v->det_bboxes = h->post.det_bboxes; v->det_scores = h->post.det_scores;
    v->num_keep = h->post.num_keep; v->num_cand = h->post.num_cand;
    v->keep_slots = h->post.keep_slots; v->keep_anchors = h->post.keep_anchors;
    v->track_poses = h->trk.poses; v->track_scores = h->trk.scores; v->track_states = h->trk.states;
    v->track_ids = h->trk.ids; v->track_outputs = h->trk.outputs; v->num_outputs = h->trk.num_outputs;
    v->num_active = h->trk.scalars;
    return PB_OK;
}

int pb_get_post_stage_us(pb_handle_t h, double* out5) {
    if (!h || !out5) { pb_set_error("pb_get_post_stage_us: bad argument"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    PB_CUDA(cudaDeviceSynchronize());
    const int B = h->cfg.num_streams;
    std::vector<unsigned long long> ns((size_t)B * 16);
    PB_CUDA(cudaMemcpy(ns.data(), h->post.stage_ns, ns.size() * 8, cudaMemcpyDeviceToHost));
    double acc[16] = {0};
    for (int b = 0; b < B; ++b) for (int i = 0; i < 16; ++i) acc[i] += (double)ns[(size_t)b * 16 + i];
    for (int i = 0; i < 5; ++i) out5[i] = acc[7] > 0 ? acc[i] / acc[7] / 1e3 : 0.0;
    return PB_OK;
}

int pb_get_nms_path_counts(pb_handle_t h, long long* fast_path, long long* complete_path, long long* keypoint_fetches) {
    if (!h) { pb_set_error("pb_get_nms_path_counts: null handle"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    PB_CUDA(cudaDeviceSynchronize());
    const int B = h->cfg.num_streams;
    std::vector<unsigned long long> ns((size_t)B * 16);
    PB_CUDA(cudaMemcpy(ns.data(), h->post.stage_ns, ns.size() * 8, cudaMemcpyDeviceToHost));
    unsigned long long fast = 0, full = 0, fetched = 0;
    for (int b = 0; b < B; ++b) { fast += ns[(size_t)b * 16 + 7]; full += ns[(size_t)b * 16 + 8]; fetched += ns[(size_t)b * 16 + 9]; }
    if (fast_path) *fast_path = (long long)fast;
    if (complete_path) *complete_path = (long long)full;
    if (keypoint_fetches) *keypoint_fetches = (long long)fetched;
    return PB_OK;
}

int pb_set_profiling(pb_handle_t h, int enabled) {
    if (!h) { pb_set_error("pb_set_profiling: null handle"); return PB_ERR_INVALID; }
    h->profiling = enabled != 0;
    return PB_OK;
}

int pb_get_kernel_ms(pb_handle_t h, double* post_ms, int* post_n, double* track_ms, int* track_n) {
    if (!h) { pb_set_error("pb_get_kernel_ms: null handle"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    PB_CUDA(cudaDeviceSynchronize());
    auto sum = [&](std::vector<std::pair<int, int>>& v, double* ms, int* n) {
        double acc = 0;
        for (auto& pr : v) { float t = 0; if (cudaEventElapsedTime(&t, h->ev_pool[pr.first], h->ev_pool[pr.second]) == cudaSuccess) acc += t; }
        if (ms) *ms = acc;
        if (n) *n = (int)v.size();
        v.clear();
    };
    sum(h->ev_post, post_ms, post_n);
    sum(h->ev_track, track_ms, track_n);
    h->ev_gather.clear();
    h->ev_used = 0;
    return PB_OK;
}

int pb_get_kernel_us(pb_handle_t h, double* gather_us, double* nms_us, double* track_us, int* launches) {
    if (!h) { pb_set_error("pb_get_kernel_us: null handle"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    PB_CUDA(cudaDeviceSynchronize());
    auto total = [&](std::vector<std::pair<int, int>>& v) {
        double acc = 0;
        for (auto& pr : v) { float t = 0; if (cudaEventElapsedTime(&t, h->ev_pool[pr.first], h->ev_pool[pr.second]) == cudaSuccess) acc += t; }
        return acc * 1e3;
    };
    const double g = total(h->ev_gather), p = total(h->ev_post), t = total(h->ev_track);
    const int n = (int)h->ev_post.size(), nt = (int)h->ev_track.size();
    if (gather_us) *gather_us = n ? g / n : 0.0;
    if (nms_us) *nms_us = n ? (p - g) / n : 0.0;
    if (track_us) *track_us = nt ? t / nt : 0.0;
    if (launches) *launches = n;
    h->ev_gather.clear(); h->ev_post.clear(); h->ev_track.clear();
    h->ev_used = 0;
    return PB_OK;
}

int pb_get_stream_stage_ns(pb_handle_t h, unsigned long long* out) {
    if (!h || !out) { pb_set_error("pb_get_stream_stage_ns: bad argument"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    PB_CUDA(cudaDeviceSynchronize());
    PB_CUDA(cudaMemcpy(out, h->trk.stage_ns, (size_t)h->cfg.num_streams * 20 * 8, cudaMemcpyDeviceToHost));
    return PB_OK;
}

int pb_debug_timeline(pb_handle_t h, unsigned long long* out) {
    if (!h || !out) { pb_set_error("pb_debug_timeline: bad argument"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    if (!h->trk.dbg) { pb_set_error("pb_debug_timeline: create the handle with PB_TIMELINE=1 in the environment"); return PB_ERR_UNSUPPORTED; }
    PB_CUDA(cudaDeviceSynchronize());
    PB_CUDA(cudaMemcpy(out, h->trk.dbg, (size_t)64 * h->cfg.num_streams * 6 * 8, cudaMemcpyDeviceToHost));
    return PB_OK;
}

int pb_get_timing(pb_handle_t h, pb_timing* out) {
    if (!h || !out) { pb_set_error("pb_get_timing: bad argument"); return PB_ERR_INVALID; }
    DevGuard dev_guard(h->cfg.device);
    PB_CUDA(cudaDeviceSynchronize());
    const int B = h->cfg.num_streams;
    std::vector<unsigned long long> ns((size_t)B * 20);
    PB_CUDA(cudaMemcpy(ns.data(), h->trk.stage_ns, ns.size() * 8, cudaMemcpyDeviceToHost));
    unsigned long long acc[20] = {0};
    for (int b = 0; b < B; ++b) for (int i = 0; i < 20; ++i) acc[i] += ns[(size_t)b * 20 + i];
    if (getenv("PB_DEBUG_STAGES")) {
        const char* names[20] = {"prologue", "predict", "gate", "t1rest", "tier2", "tier3", "update", "age", "new", "dedup",
                                 "total", "frames", "t1cost", "t1auction", "t1lock", "", "", "", "", ""};
        for (int i = 0; i < 15; ++i) {
            if (i == 11) continue;
            double mn = 1e30, mx = 0, sm = 0;
            for (int b = 0; b < B; ++b) {
                const double f = (double)ns[(size_t)b * 20 + 11];
                const double v = f > 0 ? ns[(size_t)b * 20 + i] / 1e3 / f : 0;
                mn = v < mn ? v : mn; mx = v > mx ? v : mx; sm += v;
            }
            fprintf(stderr, "[pb] %-10s us/frame over streams: min %7.2f mean %7.2f max %7.2f\n", names[i], mn, sm / B, mx);
        }
    }
    if (getenv("PB_DEBUG_STAGES"))
        fprintf(stderr, "[pb] tier1 us/frame: cost %.2f auction %.2f lock %.2f | prologue %.2f\n",
                acc[12] / 1e3 / B / (double)(acc[11] / B), acc[13] / 1e3 / B / (double)(acc[11] / B),
                acc[14] / 1e3 / B / (double)(acc[11] / B), acc[0] / 1e3 / B / (double)(acc[11] / B));
    auto us = [&](int i) { return (long long)(acc[i] / 1000ull / (unsigned long long)B); };   // mean over streams
    out->predict_us = us(1); out->gate_us = us(2); out->high_assoc_us = us(3); out->low_assoc_us = us(4);
    out->lost_assoc_us = us(5); out->update_us = us(6); out->age_us = us(7); out->new_track_us = us(8);
    out->dedup_us = us(9); out->total_us = us(10);
    out->frame_count = (int)(acc[11] / (unsigned long long)B);
    return PB_OK;
}

}  // extern "C"
