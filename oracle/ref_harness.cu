// oracle/ref_harness.cu — C entry points over the UNMODIFIED reference classes.
//
// TEST INFRASTRUCTURE ONLY.  Compiled (oracle/Makefile, target `ref`) together with the
// reference's own src/cuda/*.cu, taken where they lie under /root/reference, into
// oracle/_ref/libposebyte_ref.so.  No reference source is copied into this repository;
// this file only calls the reference's public classes (and reads a few private device
// pointers of GPUTracker for state comparison).  Uses:
//   * pinning the CPU restatement (oracle/posebyte_oracle.cpp) against the reference's real
//     code: NMSCuda::apply on the host, GPUPostprocess/GPUTracker/KalmanFilterCUDA/
//     LinearAssignmentCUDA on a B200;
//   * the "reference on the same box" timing that bench.py reports beside its own number.
#include <cstdio>
#include <cstring>
#include <chrono>
#include <memory>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#include "types.h"
#include "cuda/kalman_filter.h"
#include "cuda/oks_distance.h"
#include "cuda/hungarian.h"
#include "cuda/gpu_postprocess.h"
#include "cuda/nms.h"
#include "cuda/preprocess.h"
#define private public
#include "cuda/gpu_tracker.h"
#undef private

using namespace posebyte;
using namespace posebyte::cuda;

extern "C" {

// ---- GPUPostprocess -------------------------------------------------------------------
void* ref_post_create(int max_detections, int num_anchors) { return new GPUPostprocess(max_detections, num_anchors); }
void ref_post_destroy(void* p) { delete static_cast<GPUPostprocess*>(p); }
int ref_post_process(void* p, const float* d_raw, float conf, float nms) {
    return static_cast<GPUPostprocess*>(p)->process(d_raw, conf, nms, 0);
}
void ref_post_get(void* p, int n, float* poses, float* bboxes, float* scores) {
    GPUPostprocess* g = static_cast<GPUPostprocess*>(p);
    if (n <= 0) return;
    if (poses) cudaMemcpy(poses, g->getDetectionPoses(), (size_t)n * 51 * 4, cudaMemcpyDeviceToHost);
    if (bboxes) cudaMemcpy(bboxes, g->getDetectionBboxes(), (size_t)n * 16, cudaMemcpyDeviceToHost);
    if (scores) cudaMemcpy(scores, g->getDetectionScores(), (size_t)n * 4, cudaMemcpyDeviceToHost);
}
void* ref_post_poses_dev(void* p) { return static_cast<GPUPostprocess*>(p)->getDetectionPoses(); }
void* ref_post_scores_dev(void* p) { return static_cast<GPUPostprocess*>(p)->getDetectionScores(); }

// ---- GPUTracker -----------------------------------------------------------------------
void* ref_tracker_create(int max_tracks, int max_detections, float match_threshold, float high_thresh,
                         float low_thresh, float new_track_thresh, int max_age, int min_hits) {
    GPUTrackerConfig c;
    c.max_tracks = max_tracks; c.max_detections = max_detections; c.match_threshold = match_threshold;
    c.high_thresh = high_thresh; c.low_thresh = low_thresh; c.new_track_thresh = new_track_thresh;
    c.max_age = max_age; c.min_hits = min_hits;
    GPUTracker* t = new GPUTracker(c);
    // A fresh cudaMalloc is not guaranteed to be zero; the restatement models it as zero.
    // Zero the persistent buffers the reference never initialises so both start equal.
    cudaMemset(t->d_cost_matrix_, 0, (size_t)max_tracks * max_detections * 4);
    cudaMemset(t->d_predicted_poses_, 0, (size_t)max_tracks * 51 * 4);
    cudaMemset(t->d_track_centers_, 0, (size_t)max_tracks * 16);
    cudaMemset(t->d_det_centers_, 0, (size_t)max_detections * 16);
    cudaMemset(t->d_track_poses_, 0, (size_t)max_tracks * 51 * 4);
    cudaMemset(t->d_track_scores_, 0, (size_t)max_tracks * 4);
    cudaMemset(t->d_track_ids_, 0, (size_t)max_tracks * 4);
    cudaMemset(t->d_track_hits_, 0, (size_t)max_tracks * 4);
    cudaMemset(t->d_track_ages_, 0, (size_t)max_tracks * 4);
    cudaMemset(t->d_track_last_frame_, 0, (size_t)max_tracks * 4);
    cudaMemset(t->d_gate_mask_, 0, (size_t)max_tracks * max_detections * 4);
    cudaMemset(t->d_lost_gate_mask_, 0, (size_t)max_tracks * max_detections * 4);
    cudaDeviceSynchronize();
    return t;
}
void ref_tracker_destroy(void* t) { delete static_cast<GPUTracker*>(t); }
int ref_tracker_update(void* t, const float* d_poses, const float* d_scores, int n, int frame) {
    return static_cast<GPUTracker*>(t)->update(d_poses, d_scores, n, frame);
}
int ref_tracker_get_tracks(void* t, void* out, int cap) {
    std::vector<TrackOutput> v = static_cast<GPUTracker*>(t)->getActiveTracks();
    int n = std::min((int)v.size(), cap);
    if (n > 0) memcpy(out, v.data(), (size_t)n * sizeof(TrackOutput));
    return n;
}
void ref_tracker_get_state(void* tp, float* poses, float* vel, float* scores, int* states, int* ids,
                           int* hits, int* ages, int* last_frame, int* active, int* row_assign,
                           int* col_assign, float* cost, float* predicted, float* centers, int* scalars) {
    GPUTracker* t = static_cast<GPUTracker*>(tp);
    const size_t T = t->config_.max_tracks, Dm = t->config_.max_detections;
    cudaDeviceSynchronize();
    auto get = [](void* dst, const void* src, size_t bytes) { if (dst) cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost); };
    get(poses, t->d_track_poses_, T * 51 * 4); get(vel, t->d_track_velocities_, T * 34 * 4);
    get(scores, t->d_track_scores_, T * 4); get(states, t->d_track_states_, T * 4);
    get(ids, t->d_track_ids_, T * 4); get(hits, t->d_track_hits_, T * 4); get(ages, t->d_track_ages_, T * 4);
    get(last_frame, t->d_track_last_frame_, T * 4); get(active, t->d_track_active_, T * 4);
    get(row_assign, t->d_row_assignments_, T * 4); get(col_assign, t->d_col_assignments_, Dm * 4);
    get(cost, t->d_cost_matrix_, T * Dm * 4); get(predicted, t->d_predicted_poses_, T * 51 * 4);
    get(centers, t->d_track_centers_, T * 16);
    if (scalars) {
        int nid = 0, hint = 0;
        cudaMemcpy(&nid, t->d_next_track_id_, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&hint, t->d_next_slot_hint_, 4, cudaMemcpyDeviceToHost);
        scalars[0] = nid; scalars[1] = hint; scalars[2] = t->current_num_detections_; scalars[3] = t->num_active_tracks_;
    }
}

// ---- NMSCuda::apply (host code; runs without a GPU) ---------------------------------------
int ref_nms_apply(const void* dets, int n, float oks_thr, float score_thr, int* keep) {
    static NMSCuda* nms = nullptr;
    if (!nms) nms = new NMSCuda(16);
    std::vector<int> k = nms->apply(static_cast<const PoseDetection*>(dets), n, oks_thr, score_thr);
    for (size_t i = 0; i < k.size(); ++i) keep[i] = k[i];
    return (int)k.size();
}

// ---- LinearAssignmentCUDA -------------------------------------------------------------
void ref_auction(const float* h_cost, int R, int C, int* h_row, int* h_col, const int* h_active) {
    LinearAssignmentCUDA la(std::max(R, C));
    float* d_cost; int *d_row, *d_col, *d_act = nullptr;
    cudaMalloc(&d_cost, (size_t)R * C * 4); cudaMalloc(&d_row, R * 4); cudaMalloc(&d_col, C * 4);
    cudaMemcpy(d_cost, h_cost, (size_t)R * C * 4, cudaMemcpyHostToDevice);
    if (h_active) { cudaMalloc(&d_act, R * 4); cudaMemcpy(d_act, h_active, R * 4, cudaMemcpyHostToDevice); }
    la.solveDeviceAsyncWithActive(d_cost, R, C, d_row, d_col, d_act, 0.5f, 0);
    cudaDeviceSynchronize();
    cudaMemcpy(h_row, d_row, R * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(h_col, d_col, C * 4, cudaMemcpyDeviceToHost);
    cudaFree(d_cost); cudaFree(d_row); cudaFree(d_col); if (d_act) cudaFree(d_act);
}

// LinearAssignmentCUDA::solve (legacy host entry point, hungarian.cu:235-339); returns the count.
int ref_assign_solve(const float* h_cost, int R, int C, float threshold, int* h_row, int* h_col) {
    LinearAssignmentCUDA la(std::max(std::max(R, C), 1));
    return la.solve(h_cost, R, C, h_row, h_col, threshold);
}

// ---- PreprocessorCUDA (preprocess.cu) -----------------------------------------------------------
// h_bgr [h,w,3] u8 -> h_out [3,th,tw] fp32, xform4 = {scale_x, scale_y, pad_x, pad_y}
void ref_preprocess(const unsigned char* h_bgr, int w, int h, int tw, int th, float* h_out, float* xform4) {
    PreprocessorCUDA pp(w, h, tw, th);
    float sx = 0, sy = 0; int px = 0, py = 0;
    pp.preprocess(h_bgr, w, h, pp.getDeviceOutput(), sx, sy, px, py);
    cudaMemcpy(h_out, pp.getDeviceOutput(), (size_t)3 * tw * th * 4, cudaMemcpyDeviceToHost);
    xform4[0] = sx; xform4[1] = sy; xform4[2] = (float)px; xform4[3] = (float)py;
}

// ---- OKSDistanceCUDA / GreedyMatcherCUDA -------------------------------------------------------
// tracks / dets: PoseDetection arrays (224 B); mode 0 OKS, 1 IoU, 2 combined.
void ref_pose_distance(const void* tracks, const void* dets, int nt, int nd, int mode, float alpha, float* out) {
    OKSDistanceCUDA od(std::max(nt, 1), std::max(nd, 1));
    const PoseDetection* t = static_cast<const PoseDetection*>(tracks);
    const PoseDetection* d = static_cast<const PoseDetection*>(dets);
    if (mode == 0) od.computeOKSDistance(t, d, out, nt, nd);
    else if (mode == 1) od.computeIoUDistance(t, d, out, nt, nd);
    else od.computeCombinedDistance(t, d, out, nt, nd, alpha);
}
// GreedyMatcherCUDA::match: deterministic host path below 200 cells, racy device kernel above.
int ref_greedy_match(const float* cost, int R, int C, float threshold, int* row_matched) {
    GreedyMatcherCUDA gm(std::max(std::max(R, C), 1));
    std::vector<std::pair<int, int>> m = gm.match(cost, R, C, threshold);
    for (int r = 0; r < R; ++r) row_matched[r] = -1;
    for (auto& pr : m) row_matched[pr.first] = pr.second;
    return (int)m.size();
}

// ---- KalmanFilterCUDA -----------------------------------------------------------------
void* ref_kf3_create(int max_tracks) { return new KalmanFilterCUDA(max_tracks); }
void ref_kf3_destroy(void* k) { delete static_cast<KalmanFilterCUDA*>(k); }
void ref_kf3_initiate(void* kp, const float* h_dets, const int* h_slots, int n) {
    KalmanFilterCUDA* k = static_cast<KalmanFilterCUDA*>(kp);
    float* d; int* s;
    cudaMalloc(&d, (size_t)n * 51 * 4); cudaMalloc(&s, n * 4);
    cudaMemcpy(d, h_dets, (size_t)n * 51 * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(s, h_slots, n * 4, cudaMemcpyHostToDevice);
    k->initiateBatchAsync(d, s, n, 0); k->sync(0); cudaDeviceSynchronize();
    cudaFree(d); cudaFree(s);
}
void ref_kf3_predict(void* kp, int n, float am, float jm) { static_cast<KalmanFilterCUDA*>(kp)->predict(n, am, jm); }
void ref_kf3_update(void* kp, const float* h_dets, int ndets, const int* h_matches, int n) {
    KalmanFilterCUDA* k = static_cast<KalmanFilterCUDA*>(kp);
    float* d; int* m;
    cudaMalloc(&d, (size_t)ndets * 51 * 4); cudaMalloc(&m, n * 8);
    cudaMemcpy(d, h_dets, (size_t)ndets * 51 * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(m, h_matches, n * 8, cudaMemcpyHostToDevice);
    k->updateBatchAsync(d, m, n, 0); k->sync(0); cudaDeviceSynchronize();
    cudaFree(d); cudaFree(m);
}
// means [T,136] and the covariance diagonal [T,136]; returns the largest |off-diagonal| seen.
float ref_kf3_get(void* kp, int T, float* means, float* diag) {
    KalmanFilterCUDA* k = static_cast<KalmanFilterCUDA*>(kp);
    cudaDeviceSynchronize();
    cudaMemcpy(means, k->getMeansDevice(), (size_t)T * 136 * 4, cudaMemcpyDeviceToHost);
    std::vector<float> cov((size_t)136 * 136);
    float off = 0.f;
    for (int t = 0; t < T; ++t) {
        cudaMemcpy(cov.data(), k->getCovariancesDevice() + (size_t)t * 136 * 136, cov.size() * 4, cudaMemcpyDeviceToHost);
        for (int i = 0; i < 136; ++i)
            for (int j = 0; j < 136; ++j) {
                if (i == j) diag[(size_t)t * 136 + i] = cov[i * 136 + j];
                else off = std::max(off, std::abs(cov[i * 136 + j]));
            }
    }
    return off;
}

// ---- timing: the reference's own frame loop (main.cpp:207-224) on one stream ----------------
// d_heads: [F,56,N] on the device.  Returns mean milliseconds per frame over `frames`
// frames (after `warm` untimed ones), wall clock around process()+update()+getActiveTracks().
double ref_time_frames(const float* d_heads, int F, int N, int warm, int frames, float conf, float nms,
                       int max_tracks, int max_dets, int max_age, int with_readback) {
    GPUPostprocess post(1024, N);
    GPUTrackerConfig c;
    c.max_tracks = max_tracks; c.max_detections = max_dets; c.match_threshold = 0.5f;
    c.high_thresh = conf; c.low_thresh = conf * 0.5f; c.new_track_thresh = conf; c.min_hits = 3; c.max_age = max_age;
    GPUTracker trk(c);
    const size_t slab = (size_t)56 * N;
    auto run = [&](int f) {
        int n = post.process(d_heads + (size_t)(f % F) * slab, conf, nms, 0);
        trk.update(post.getDetectionPoses(), post.getDetectionScores(), n, f);
        if (with_readback) { auto v = trk.getActiveTracks(); (void)v; }
    };
    for (int f = 0; f < warm; ++f) run(f);
    cudaDeviceSynchronize();
    auto t0 = std::chrono::steady_clock::now();
    for (int f = warm; f < warm + frames; ++f) run(f);
    cudaDeviceSynchronize();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::milli>(t1 - t0).count() / frames;
}

}  // extern "C"
