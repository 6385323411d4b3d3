/* oracle/posebyte_oracle.h — C interface of the CPU checker.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may
 * load this library.  It is a scalar restatement of the reference's CUDA kernels
 * (naveedprojects/yolo-pose-cpp, src/cuda/ *.cu) with the reference's data races
 * resolved by the rules R1-R6 written in DESIGN.md.
 *
 * Pinning status: the reference ships no tests, golden vectors or known-answer
 * fixtures for this path.  The restatement is pinned against the
 * reference's own code instead: (1) NMSCuda::apply compiled from the reference's
 * src/cuda/nms.cu and run on the host (oracle/_ref, tests/test_oracle_vs_ref_host.py);
 * (2) the reference's GPU kernels compiled unchanged for sm_100a and run on a B200
 * (oracle/_ref/libposebyte_ref.so, tests/test_gpu_ref_crosscheck.py; fixtures produced
 * by that run are committed under tests/golden/).
 */
#ifndef POSEBYTE_ORACLE_H
#define POSEBYTE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_tracker_config {
    int max_tracks;
    int max_detections;
    float match_threshold;   /* inert in the reference (hungarian.cu:358-405) */
    float high_thresh;       /* inert: masks computed, never read (gpu_tracker.cu:595) */
    float low_thresh;        /* inert */
    float new_track_thresh;
    int max_age;
    int min_hits;
    int gating_enabled;      /* extension: 0 = all-ones spatial gate */
} orc_tracker_config;

/* pb_expf of include/pb_math.h over an array (accuracy tests). */
void orc_expf_array(const float* x, float* y, int n);
/* A2 literal: the full C x C overlap matrix, one byte per bit (gpu_postprocess.cu:88-172). */
void orc_nms_mask(const float* poses, const float* bboxes, int C, float thr, unsigned char* mask);

/* A1  gpu_postprocess.cu:30-81 (R1: ascending anchor order, first max_cand kept). */
int orc_decode(const float* raw, int num_anchors, float conf_thr, int max_cand,
               float* poses, float* bboxes, float* scores, int* anchors);

/* A2+A3  gpu_postprocess.cu:88-313,366-476.  In-place on the candidate arrays, as the
 * reference does; keep_slots receives the candidate slots in score order. */
int orc_nms_native(float* poses, float* bboxes, float* scores, int num_cand,
                   float nms_thr, int max_keep, int* keep_slots);

/* A1+A2+A3 for one stream.  Returns Kp; *num_cand receives C. */
int orc_postprocess(const float* raw, int num_anchors, float conf_thr, float nms_thr,
                    int max_cand, int max_keep,
                    float* poses, float* bboxes, float* scores,
                    int* keep_slots, int* keep_anchors, int* num_cand);

/* A4  nms.cu:142-306 (host-legacy rule set).  dets = PoseDetection[n] (224 B each). */
int orc_nms_legacy(const void* dets, int n, float oks_thr, float score_thr, int* keep);

/* launchPoseNMS (nms.h:48-60, undefined upstream): OKS of nms.cu:25-117 + stable greedy. */
void orc_pose_nms(const float* poses, const float* scores, const float* sigmas, int* keep,
                  int n, int num_keypoints, float oks_thr, float score_thr);

/* A11  hungarian.cu:27-123,358-405.  row_active may be NULL. */
void orc_auction(const float* cost, int num_rows, int num_cols,
                 int* row_assign, int* col_assign, const int* row_active);

/* SURVEY.md 8f row f1.  OKSDistanceCUDA: oks_distance.cu:26-261 (mode 0 OKS cost, 1 keypoint-box IoU
 * cost, 2 alpha-combined); tracks [nt,51], dets [nd,51], out [nt,nd]. */
void orc_pose_distance(const float* tracks, const float* dets, int nt, int nd, int mode, float alpha, float* out);
/* GreedyMatcherCUDA::match host rule, hungarian.cu:441-467; row_matched [R] = column or -1. */
void orc_greedy_match(const float* cost, int R, int C, float threshold, int* row_matched);
/* LinearAssignmentCUDA::solve (hungarian.cu:235-339): greedy below 100 cells, else auction with 3*rows
 * iterations and the host threshold filter.  Returns the number of assignments kept. */
int orc_assign_legacy(const float* cost, int R, int C, float threshold, int* row, int* col);
/* PreprocessorCUDA::preprocess (preprocess.cu:19-153): bgr [h,w,3] u8 -> out [3,th,tw] fp32,
 * xform4 = {scale_x, scale_y, pad_x, pad_y}. */
void orc_letterbox(const unsigned char* bgr, int w, int h, int tw, int th, float* out, float* xform4);

/* A5-A16  gpu_tracker.cu:102-919,1057-1639. */
void* orc_tracker_create(const orc_tracker_config* cfg);
void orc_tracker_destroy(void* t);
int orc_tracker_update(void* t, const float* det_poses, const float* det_scores,
                       int num_dets, int frame_id);
/* Replay mode for pinning against a recorded run of the reference: for the NEXT update only,
 * new tracks take slots[d] / ids[d] (d = detection index, n entries, slot -1 = none) instead
 * of rules R3/R4 — i.e. the outcome of the reference's atomics race is supplied, everything
 * else is computed.  orc_tracker_forced_errors counts detections whose supplied slot was
 * missing or occupied. */
void orc_tracker_force_new(void* t, const int* slots, const int* ids, int n);
int orc_tracker_forced_errors(void* t);
int orc_tracker_get_tracks(void* t, void* track_outputs /*TrackOutput[cap]*/, int cap);
/* Raw state readback; any pointer may be NULL.  Sizes: poses T*51, vel T*34, scores T,
 * states/ids/hits/ages/last_frame/active T, row_assign T, col_assign Dmax,
 * cost T*Dmax, predicted T*51, centers T*4, scalars[4] = {next_id, slot_hint, D, num_active}. */
void orc_tracker_get_state(void* t, float* poses, float* vel, float* scores, int* states,
                           int* ids, int* hits, int* ages, int* last_frame, int* active,
                           int* row_assign, int* col_assign, float* cost, float* predicted,
                           float* centers, int* scalars);

/* K1-K4  kalman_filter.cu:24-283,422-491.  State = mean[T,136] + covariance diagonal[T,136]
 * (the reference keeps a 136x136 matrix per track of which only the diagonal is ever
 * non-zero; orc_kf3_get_state materialises it). */
void* orc_kf3_create(int max_tracks);
void orc_kf3_destroy(void* k);
void orc_kf3_initiate(void* k, const float* dets, const int* slots, int n);
void orc_kf3_predict(void* k, int n, float accel_memory, float jerk_memory);
void orc_kf3_update(void* k, const float* dets, const int* matches, int n);
void orc_kf3_extract(void* k, float* out_poses, const int* slots, int n);
void orc_kf3_get_state(void* k, int track, float* mean136, float* cov136x136_or_null);
void orc_kf3_get_diag(void* k, float* means, float* diag); /* [T,136] each */

/* Whole path over many independent streams (CPU baseline).  heads = [B][F][56][N] when
 * frame_major==0, [F][B][56][N] otherwise.  Streams are spread over n_threads.
 * out_hash[B] receives a 64-bit FNV hash of every frame's TrackOutput records and kept
 * anchors; stage_seconds[3] = summed decode / nms / track time over all threads.
 * Returns wall seconds of the parallel region. */
double orc_run_streams(const float* heads, int B, int F, int num_anchors, int frame_major,
                       float conf_thr, float nms_thr, int max_cand, int max_keep,
                       const orc_tracker_config* cfg, int n_threads,
                       unsigned long long* out_hash, long long* out_tracks_total,
                       double* stage_seconds);

#ifdef __cplusplus
}
#endif
#endif
