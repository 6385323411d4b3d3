/* include/posebyte_b200.h — C ABI of the B200-native PoseBYTE post-inference path.
 *
 * This is the drop-in boundary.  The reference (naveedprojects/yolo-pose-cpp) has no
 * FFI layer: its boundary is a set of C++ classes over raw device pointers
 * (include/cuda/gpu_postprocess.h:11-72, gpu_tracker.h:61-110, nms.h:11-60,
 * kalman_filter.h:14-138, hungarian.h:11-160).  Every entry point below names the
 * reference interface it replaces; include/cuda/ *.h re-creates those classes as
 * header-only shims on top of this ABI, so main.cpp:207-224 compiles unchanged.
 *
 * Conventions: plain pointers and sizes only; device pointers are prefixed d_, host
 * pointers h_; every function returns PB_OK (0) or a negative pb_status and never calls
 * exit() (the reference prints-and-continues or exits, gpu_tracker.cu:9-16,
 * hungarian.cu:11-19); pb_last_error() returns the text of the last failure on the
 * calling thread.  A handle owns B independent streams (one tracker state each) that
 * are processed by one launch; nothing is shared between streams.  All work is enqueued
 * on the caller's CUDA stream; functions with "_sync" or host outputs synchronise it.
 */
#ifndef POSEBYTE_B200_H
#define POSEBYTE_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* pb_stream_t;        /* == cudaStream_t */
typedef struct pb_handle_st* pb_handle_t;

typedef enum pb_status {
    PB_OK = 0,
    PB_ERR_INVALID = -1,     /* bad argument / configuration */
    PB_ERR_CUDA = -2,        /* a CUDA runtime call failed */
    PB_ERR_UNSUPPORTED = -3, /* configuration exceeds a compiled-in limit */
    PB_ERR_NO_DEVICE = -4    /* no sm_100 device: there is no CPU fallback */
} pb_status;

/* GPUTrackerConfig (gpu_tracker.h:16-26) + GPUPostprocess ctor args
 * (gpu_postprocess.h:13) + batching.  Zero-initialise, call pb_default_config, then edit. */
typedef struct pb_config {
    int num_streams;         /* B: independent video streams in this handle (new capability) */
    int num_anchors;         /* N: 8400 @640, 33600 @1280 (GPUPostprocess num_anchors) */
    int max_candidates;      /* GPUPostprocess max_detections, 1024 */
    int max_keep;            /* hard-coded 256 upstream (gpu_postprocess.cu:224) */
    int max_tracks;          /* 128 */
    int max_detections;      /* 64: tracker input is truncated to this (gpu_tracker.cu:1066) */
    float match_threshold;   /* 0.5 - kept for API parity, inert upstream (hungarian.cu:358-405) */
    float high_thresh;       /* 0.30 - inert upstream (outputs never read, gpu_tracker.cu:595) */
    float low_thresh;        /* 0.15 - inert upstream */
    float new_track_thresh;  /* 0.30 */
    int max_age;             /* 10 */
    int min_hits;            /* 3 */
    int use_cuda_graph;      /* kept for API parity, unused upstream (gpu_tracker.cu:1660) */
    int gating_enabled;      /* extension: 0 replaces the spatial gate by all-ones (config 5) */
    int device;              /* CUDA device ordinal */
    int pipeline_depth;      /* 1: pb_step runs entirely on the caller's stream.  2..12: pb_step
                              * overlaps consecutive steps on internal streams (see pb_join) */
    int keypoint_fetch;      /* 0 (default): the NMS kernel fetches keypoints lazily (IoU rule first, keypoints
                              * only for the ranks whose OKS tests are unavoidable) iff the head is read in
                              * place from host memory (pb_step_host); 1: always lazy; 2: never (complete
                              * sweep); 3: complete records, but the lazy sweep's evaluation order (OKS tests
                              * deferred until a rank's own tile; ~12 % faster on dense 1280x1280 crowds) */
    int fuse_stages;         /* NMS and tracker update of a stream-frame in ONE CTA (fused.cu: kept detections handed over in
                              * shared memory, shared memory sized by the candidates present so that two stream-CTAs share an
                              * SM, frames of a video stream handed from CTA to CTA through a chain word): 1 wherever the tables
                              * allow it, 0 never (separate NMS and tracker kernels), 2 (default) where it is the faster plan —
                              * measured on B200: more than 74 streams per handle, where two one-CTA-per-SM tracker grids no
                              * longer fit the device side by side.  Same results either way. */
} pb_config;

/* TrackerTiming (gpu_tracker.h:29-41), filled from device timestamps. */
typedef struct pb_timing {
    long long predict_us, gate_us, high_assoc_us, low_assoc_us, lost_assoc_us;
    long long update_us, age_us, new_track_us, dedup_us, total_us;
    int frame_count;
} pb_timing;

const char* pb_last_error(void);
const char* pb_version(void);
void pb_default_config(pb_config* cfg);

/* GPUPostprocess::GPUPostprocess + GPUTracker::GPUTracker (gpu_postprocess.cu:319-347,
 * gpu_tracker.cu:925-1010).  Fails with PB_ERR_NO_DEVICE when no GPU is present and with
 * PB_ERR_UNSUPPORTED for sizes the kernels do not hold: num_anchors > 65536, max_tracks or
 * max_detections >= 65536, max_tracks * max_detections^2 >= 2^32, candidate or track tables
 * beyond the shared memory of an SM (pb_last_error() names the limit). */
int pb_create(const pb_config* cfg, pb_handle_t* out);
int pb_destroy(pb_handle_t h);
/* Back to the freshly constructed state (all tracks dropped, ids restart at 1). */
int pb_reset(pb_handle_t h, pb_stream_t stream);

/* ---- the path ------------------------------------------------------------------ */

/* GPUPostprocess::process for B streams (gpu_postprocess.cu:366-476): decode + confidence
 * filter + OKS/IoU NMS.  d_heads = [B,56,N] fp32, borrowed.  Asynchronous. */
int pb_postprocess(pb_handle_t h, const float* d_heads, float conf_threshold,
                   float nms_threshold, pb_stream_t stream);

/* GPUTracker::update for B streams (gpu_tracker.cu:1057-1158).
 * d_det_poses [B, det_stride, 17, 3], d_det_scores [B, det_stride], d_num_dets [B]
 * (device).  Passing NULL pointers uses the handle's own postprocess outputs (the chain
 * main.cpp:207-221 builds).  Asynchronous. */
int pb_tracker_update(pb_handle_t h, const float* d_det_poses, const float* d_det_scores,
                      const int* d_num_dets, int det_stride, int frame_id, pb_stream_t stream);

/* detectGPUNative's tail + tracker.update + getActiveTracks assembly for B streams in one
 * call (main.cpp:207-224).  Asynchronous; results stay on the device.  d_heads is borrowed: it is
 * read by work enqueued on `stream` or ordered before later work on `stream` (also with
 * pipeline_depth > 1), so it may be reused by anything enqueued on `stream` after this call. */
int pb_step(pb_handle_t h, const float* d_heads, float conf_threshold, float nms_threshold,
            int frame_id, pb_stream_t stream);

/* n_steps consecutive pb_step calls in one: step i takes the head batch d_heads + ((first + i) % period) * step_stride
 * (floats) and frame id frame0 + i.  For callers that hold several batches in device memory (offline video, benchmarks):
 * one library call instead of n_steps crossings of the language boundary.  Same results as the loop it replaces.
 * Knowing the frames ahead lets the library keep every video stream's tracker state ON ITS SM for the whole sequence
 * (the reference's per-stream frame loop, main.cpp:207-224, as one resident CTA per stream): with pipeline_depth > 1 and
 * at most half of the SMs' worth of streams (74 on a B200; PB_SEQ=0 in the environment turns it off), chunks of up to 32 steps run
 * with ONE tracker launch each, whose CTAs take every frame's kept detections from the NMS kernel of its step as soon as
 * they are published.  pb_step_path tells which path a handle takes.  The head batches are borrowed until the work
 * enqueued on `stream` by this call has run (as for pb_step). */
int pb_step_seq(pb_handle_t h, const float* d_heads, size_t step_stride, int period, int first, int n_steps,
                float conf_threshold, float nms_threshold, int frame0, pb_stream_t stream);

/* How this handle runs its steps: *per_step 0 serial (three launches on the caller's stream), 1 pipelined three-kernel
 * step (lanes), 2 fused per-stream kernel; *seq_chunk steps per resident tracker launch of pb_step_seq (0: pb_step_seq
 * is a loop of pb_step). */
int pb_step_path(pb_handle_t h, int* per_step, int* seq_chunk);

/* Launch plan of the NMS kernel of the per-step paths: threads per CTA, CTAs that share an SM, dynamic shared memory (bytes)
 * and the number of candidates whose working set that holds (*tier_candidates == max_candidates: the plain kernel; fewer:
 * the tiered kernel, sized to the shared-memory configuration of the tracker kernel — streams with more candidates keep
 * their per-candidate arrays in a global scratch, same results).  Any pointer may be null. */
int pb_nms_plan(pb_handle_t h, int* threads, int* ctas_per_sm, int* smem_bytes, int* tier_candidates);

/* With pipeline_depth > 1 pb_step returns with NMS and tracker work still running on internal
 * streams.  pb_join makes `stream` wait for all of it (asynchronous, no host blocking); every
 * pb_get_* function and the stage-level entry points join implicitly. */
int pb_join(pb_handle_t h, pb_stream_t stream);

/* Same with HOST buffers: h_heads [B,56,N]; the TrackOutput records are returned in h_tracks
 * [B, max_detections] (228-byte records) with h_counts [B].  Synchronous.  Page-locked input
 * (cudaHostAlloc / cudaHostRegister) is read in place over PCIe by the decode kernel — only the
 * confidence rows and the sectors at candidate anchors cross the bus; pageable input is staged
 * with one full copy.  Page-locked output buffers receive the device-to-host copies directly. */
int pb_step_host(pb_handle_t h, const float* h_heads, float conf_threshold,
                 float nms_threshold, int frame_id, void* h_tracks, int* h_counts);

/* Asynchronous form of pb_step_host for throughput: all three buffers must be page-locked.  The
 * step (in-place read of h_heads over PCIe, kernels, copies of the records into h_tracks /
 * h_counts) is enqueued on the handle's own streams and overlaps earlier submissions when
 * pipeline_depth > 1.  The buffers of a submission belong to the library until pb_wait returns.
 * pb_wait blocks until every submitted step has completed and its results are in place. */
int pb_submit_host(pb_handle_t h, const float* h_heads, float conf_threshold, float nms_threshold,
                   int frame_id, void* h_tracks, int* h_counts);
int pb_wait(pb_handle_t h);

/* ---- results --------------------------------------------------------------------- */

/* GPUTracker::getActiveTracks (gpu_tracker.cu:1559-1639) for one stream; synchronises.
 * out = TrackOutput[cap]. */
int pb_get_tracks(pb_handle_t h, int stream_idx, void* out, int cap, int* n_out);
/* All streams at once: out [B, max_detections] records, counts [B]. */
int pb_get_tracks_all(pb_handle_t h, void* out, int* counts);
/* update()'s return value per stream (active slots incl. lost/tentative). */
int pb_get_num_active(pb_handle_t h, int* h_num_active /*[B]*/);

/* Post-NMS detections of one stream (getDetectionPoses/Bboxes/Scores + d_keep_indices_):
 * any pointer may be NULL.  poses [Kp,51], bboxes [Kp,4], scores [Kp], keep_slots [Kp]
 * (candidate slots, = d_keep_indices_), keep_anchors [Kp].  Synchronises. */
int pb_get_kept(pb_handle_t h, int stream_idx, float* poses, float* bboxes, float* scores,
                int* keep_slots, int* keep_anchors, int cap, int* num_keep, int* num_cand);

/* Raw tracker state of one stream (the "Kalman states" parity target); pointers may be
 * NULL.  Sizes as in oracle/posebyte_oracle.h:orc_tracker_get_state.  Synchronises. */
int pb_get_state(pb_handle_t h, int stream_idx, float* poses, float* vel, float* scores,
                 int* states, int* ids, int* hits, int* ages, int* last_frame, int* active,
                 int* row_assign, int* col_assign, float* cost, float* predicted,
                 float* centers, int* scalars);

/* Device pointers for chaining (getDetectionPoses / getTrackPosesDevice ...). */
typedef struct pb_device_views {
    float* det_poses;   /* [B, max_keep, 51] */
    float* det_bboxes;  /* [B, max_keep, 4]  */
    float* det_scores;  /* [B, max_keep]     */
    int* num_keep;      /* [B] */
    int* num_cand;      /* [B] */
    int* keep_slots;    /* [B, max_keep] */
    int* keep_anchors;  /* [B, max_keep] */
    float* track_poses; /* [B, T, 51] */
    float* track_scores;/* [B, T] */
    int* track_states;  /* [B, T] */
    int* track_ids;     /* [B, T] */
    void* track_outputs;/* [B, max_detections] TrackOutput */
    int* num_outputs;   /* [B] */
    int* num_active;    /* [B] */
} pb_device_views;
int pb_get_device_views(pb_handle_t h, pb_device_views* out);
int pb_get_timing(pb_handle_t h, pb_timing* out);
/* Raw per-stream accumulators behind pb_get_timing: out [B,20] nanosecond sums since creation:
 * 0 prologue, 1 predict, 2 gate, 3 tier-1 rest, 4 tier 2, 5 tier 3, 6 update, 7 age, 8 new,
 * 9 dedup, 10 total, 11 frame count, 12 tier-1 cost, 13 tier-1 auction, 14 tier-1 lock. */
int pb_get_stream_stage_ns(pb_handle_t h, unsigned long long* out);
/* Development aid (handle created with PB_TIMELINE=1 in the environment): absolute globaltimer stamps of the last
 * 64 steps, out [64][num_streams][6] = tracker CTA begin, state acquired, end, NMS CTA begin, end, unused;
 * slot = tracker launch sequence number mod 64. */
int pb_debug_timeline(pb_handle_t h, unsigned long long* out);
/* Number of kernels this library has launched since process start (bench bookkeeping). */
long long pb_launch_count(void);
/* Per-kernel device timing for benchmarks: when enabled, every pb_postprocess /
 * pb_tracker_update launch is bracketed by CUDA events on its stream.  pb_get_kernel_ms
 * synchronises, returns the summed milliseconds and launch counts since the last call and
 * clears the record.  Off by default (TrackerTiming's per-stage numbers come from in-kernel
 * timestamps and are always on). */
int pb_set_profiling(pb_handle_t h, int enabled);
/* Mean microseconds per launch and stream of the decode+NMS kernel's stages since creation
 * (in-kernel timestamps): scan+compaction, ranking, gather, suppression, output. */
int pb_get_post_stage_us(pb_handle_t h, double* out5);
/* Lazy-sweep telemetry since creation: stream-frames processed by the NMS kernel, and how many
 * 64-rank tiles needed their second round (keypoints fetched for every live rank of the tile
 * because a speculative survivor fell to an OKS rule). */
int pb_get_nms_path_counts(pb_handle_t h, long long* stream_frames, long long* second_rounds,
                           long long* keypoint_fetches /* candidates whose 51 keypoint values were fetched */);
int pb_get_kernel_ms(pb_handle_t h, double* post_ms, int* post_launches, double* track_ms, int* track_launches);
/* Same record as mean microseconds per launch of each of the three kernels (decode+gather, NMS,
 * tracker); clears the record. */
int pb_get_kernel_us(pb_handle_t h, double* gather_us, double* nms_us, double* track_us, int* launches);

/* ---- rows next to the path ------------------------------------------------- */

/* scaleTrackOutputs (src/main.cpp:48-68) fused into the output stage: every TrackOutput of stream b
 * leaves as (value - pad) * scale for the box and the keypoints.  h_xform = [B,4] host floats
 * (scale_x, scale_y, pad_x, pad_y) or NULL to switch it off (default).  Synchronises. */
int pb_set_output_transform(pb_handle_t h, const float* h_xform);

/* Tracker state snapshot / restore (all streams): a flat host blob of pb_state_size bytes.
 * A handle restored from a snapshot continues exactly as the saved one would have. */
int pb_state_size(pb_handle_t h, size_t* bytes);
int pb_state_save(pb_handle_t h, void* h_blob, size_t capacity);
int pb_state_load(pb_handle_t h, const void* h_blob, size_t bytes);

/* OKSDistanceCUDA device entry points (oks_distance.cu:26-261, 479-537) for `batch` independent
 * problems in one launch: d_tracks [batch, num_tracks, 17*3], d_dets [batch, num_dets, 17*3],
 * d_out_costs [batch, num_tracks, num_dets].  mode 0: computeOKSDistanceDeviceAsync (1 - OKS,
 * ungated, with the 0.05-confidence fallback); 1: computeIoUDistanceDeviceAsync (keypoint boxes
 * with a 10 px margin); 2: computeCombinedDistance (alpha * OKS cost + (1 - alpha) * IoU cost). */
int pb_pose_distance(const float* d_tracks, const float* d_dets, int batch, int num_tracks, int num_dets,
                     int mode, float alpha, float* d_out_costs, pb_stream_t stream);

/* GreedyMatcherCUDA (hungarian.cu:407-543) for `batch` problems: d_row_matched [batch, num_rows]
 * = matched column or -1.  Deterministic rule = the class's own host path (:441-467): cells below
 * the threshold in ascending (cost, row, col) order, taken when row and column are free.  (The
 * reference's device kernel, :126-157, races on the columns.) */
int pb_greedy_match(const float* d_cost, int batch, int num_rows, int num_cols, float threshold,
                    int* d_row_matched, pb_stream_t stream);

/* LinearAssignmentCUDA::solve (hungarian.cu:235-339), the legacy host-threshold entry point, for `batch`
 * problems on the device: fewer than 100 cells -> greedyAssign (:198-233); otherwise the auction over all
 * rows with at most 3*rows iterations (:283) and the convergence exit of :317, then assignments with
 * cost > threshold are cleared (:328-336).  d_count [batch] (may be NULL) = solve()'s return value. */
int pb_assign_legacy(const float* d_cost, int batch, int num_rows, int num_cols, float threshold,
                     int* d_row_assign, int* d_col_assign, int* d_count, pb_stream_t stream);

/* PreprocessorCUDA::preprocess (preprocess.cu:19-153) for a batch of frames in one launch:
 * d_frames = BGR u8 HWC images, frame f at d_frames + f*frame_stride_bytes with size d_sizes[2f] x
 * d_sizes[2f+1] (width, height); d_out [batch, 3, target_height, target_width] fp32 RGB in [0,1],
 * gray 114/255 letterbox bars; d_xform [batch, 4] (may be NULL) = {scale_x, scale_y, pad_x, pad_y} as
 * preprocess() returns them — the layout pb_set_output_transform takes to undo the letterbox. */
int pb_letterbox_batch(const unsigned char* d_frames, size_t frame_stride_bytes, const int* d_sizes, int batch,
                       int target_width, int target_height, float* d_out, float* d_xform, pb_stream_t stream);

/* ---- stage-level entry points ------------------------------------------------------ */

/* nms.h:48-60 declares this symbol and never defines it.  Device pointers; keep[i] in
 * {0,1}.  Rule: pairwise OKS as nms.cu:25-117, stable score order, suppress at
 * OKS > oks_threshold. */
void launchPoseNMS(const float* poses, const float* scores, const float* sigmas, int* keep,
                   int num_detections, int num_keypoints, float oks_threshold,
                   float score_threshold, pb_stream_t stream);

/* NMSCuda::apply / applyBatch rule set (nms.cu:142-330) on the device.
 * d_dets = PoseDetection[total] (224 B), d_offsets [num_images+1]; d_keep [total]
 * receives original indices (image-local) in score order, d_num_keep [num_images]. */
int pb_nms_legacy(const void* d_dets, const int* d_offsets, int num_images, int max_per_image,
                  float oks_threshold, float score_threshold, int* d_keep, int* d_num_keep,
                  pb_stream_t stream);

/* LinearAssignmentCUDA::solveDeviceAsyncWithActive (hungarian.cu:358-405) for `batch`
 * independent problems: d_cost [batch, rows, cols], d_row_active [batch, rows] or NULL. */
int pb_auction_solve(const float* d_cost, int batch, int num_rows, int num_cols,
                     int* d_row_assign, int* d_col_assign, const int* d_row_active,
                     pb_stream_t stream);

/* KalmanFilterCUDA (kalman_filter.cu).  State: d_means [T,136], d_diag [T,136] (the
 * covariance diagonal; off-diagonal entries are identically zero upstream). */
int pb_kf3_initiate(float* d_means, float* d_diag, const float* d_dets, const int* d_slots,
                    int num_new, pb_stream_t stream);
int pb_kf3_predict(float* d_means, float* d_diag, int num_tracks, float accel_memory,
                   float jerk_memory, pb_stream_t stream);
int pb_kf3_update(float* d_means, float* d_diag, const float* d_dets, const int* d_matches,
                  int num_matches, pb_stream_t stream);
int pb_kf3_extract(const float* d_means, float* d_out_poses, const int* d_slots, int num_tracks,
                   pb_stream_t stream);
/* getState's full 136x136 matrix from the diagonal. */
int pb_kf3_materialize_cov(const float* d_diag, int track, float* d_cov136x136, pb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
