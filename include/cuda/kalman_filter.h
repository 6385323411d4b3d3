// include/cuda/kalman_filter.h — posebyte::cuda::KalmanFilterCUDA over the B200 C ABI.
//
// The reference's 3rd-order per-keypoint filter (reference include/cuda/kalman_filter.h:14-138,
// src/cuda/kalman_filter.cu): state per keypoint [x y vx vy ax ay jx jy], 136 values per track.
// The reference stores a 136x136 covariance per track (74 KB) of which only the diagonal is
// ever non-zero; this view keeps the 136 diagonal entries (544 B) and materialises the full
// matrix on request (getState / getCovariancesDevice).
#pragma once

#include <vector>

#include "pb_shim_common.h"

namespace posebyte {
namespace cuda {

class KalmanFilterCUDA {
public:
    explicit KalmanFilterCUDA(int max_tracks = 256) : max_tracks_(max_tracks) {
        const size_t n = (size_t)max_tracks * TOTAL_STATE_DIM;
        detail::cu_check(cudaStreamCreate(&stream_), "cudaStreamCreate");
        detail::cu_check(cudaMalloc(&d_means_, n * sizeof(float)), "cudaMalloc");
        detail::cu_check(cudaMalloc(&d_diag_, n * sizeof(float)), "cudaMalloc");
        detail::cu_check(cudaMemset(d_means_, 0, n * sizeof(float)), "cudaMemset");
        detail::cu_check(cudaMemset(d_diag_, 0, n * sizeof(float)), "cudaMemset");
        detail::cu_check(cudaMalloc(&d_det_, (size_t)max_tracks * 51 * sizeof(float)), "cudaMalloc");
        detail::cu_check(cudaMalloc(&d_idx_, (size_t)max_tracks * 2 * sizeof(int)), "cudaMalloc");
    }
    ~KalmanFilterCUDA() {
        cudaFree(d_means_); cudaFree(d_diag_); cudaFree(d_det_); cudaFree(d_idx_);
        if (d_cov_) cudaFree(d_cov_);
        cudaStreamDestroy(stream_);
    }
    KalmanFilterCUDA(const KalmanFilterCUDA&) = delete;
    KalmanFilterCUDA& operator=(const KalmanFilterCUDA&) = delete;

    // ---- device-pointer operations (kalman_filter.cu:422-491) ----
    void predictAsync(int num_active_tracks, cudaStream_t stream = 0) {
        detail::pb_check(pb_kf3_predict(d_means_, d_diag_, clamp(num_active_tracks), accel_memory_, jerk_memory_, detail::as_pb(stream)), "pb_kf3_predict");
    }
    void updateBatchAsync(const float* d_detections, const int* d_matches, int num_matches, cudaStream_t stream = 0) {
        detail::pb_check(pb_kf3_update(d_means_, d_diag_, d_detections, d_matches, num_matches, detail::as_pb(stream)), "pb_kf3_update");
    }
    void initiateBatchAsync(const float* d_detections, const int* d_track_slots, int num_new, cudaStream_t stream = 0) {
        detail::pb_check(pb_kf3_initiate(d_means_, d_diag_, d_detections, d_track_slots, num_new, detail::as_pb(stream)), "pb_kf3_initiate");
    }
    void extractPosesToDeviceAsync(float* d_out_poses, const int* d_track_slots, int num_tracks, cudaStream_t stream = 0) {
        detail::pb_check(pb_kf3_extract(d_means_, d_out_poses, d_track_slots, num_tracks, detail::as_pb(stream)), "pb_kf3_extract");
    }
    void sync(cudaStream_t stream = 0) { detail::cu_check(cudaStreamSynchronize(stream), "cudaStreamSynchronize"); }

    // ---- host-pointer wrappers (kalman_filter.cu:502-647), blocking ----
    void initiate(int track_idx, const PoseDetection& detection) {
        if (track_idx < 0 || track_idx >= max_tracks_) { std::fprintf(stderr, "KalmanFilterCUDA::initiate: bad track index %d\n", track_idx); return; }
        float det[51];
        flatten(detection, det);
        detail::cu_check(cudaMemcpyAsync(d_det_, det, sizeof(det), cudaMemcpyHostToDevice, stream_), "upload");
        detail::cu_check(cudaMemcpyAsync(d_idx_, &track_idx, sizeof(int), cudaMemcpyHostToDevice, stream_), "upload");
        initiateBatchAsync(d_det_, d_idx_, 1, stream_);
        sync(stream_);
    }
    void predict(int num_tracks, float accel_memory = 0.9f, float jerk_memory = 0.9f) {
        setMotionParams(accel_memory, jerk_memory);
        predictAsync(num_tracks, stream_);
        sync(stream_);
    }
    // matches: [num_matches, 2] = (track slot, detection index)
    void update(const PoseDetection* detections, const int* matches, int num_matches) {
        if (num_matches <= 0) return;
        int max_det = 0;
        for (int i = 0; i < num_matches; ++i) if (matches[2 * i + 1] > max_det) max_det = matches[2 * i + 1];
        if (max_det >= max_tracks_ || num_matches > max_tracks_) { std::fprintf(stderr, "KalmanFilterCUDA::update: too many detections\n"); return; }
        std::vector<float> flat((size_t)(max_det + 1) * 51);
        for (int d = 0; d <= max_det; ++d) flatten(detections[d], &flat[(size_t)d * 51]);
        detail::cu_check(cudaMemcpyAsync(d_det_, flat.data(), flat.size() * sizeof(float), cudaMemcpyHostToDevice, stream_), "upload");
        detail::cu_check(cudaMemcpyAsync(d_idx_, matches, (size_t)num_matches * 2 * sizeof(int), cudaMemcpyHostToDevice, stream_), "upload");
        updateBatchAsync(d_det_, d_idx_, num_matches, stream_);
        sync(stream_);
    }
    void getPredictedPose(int track_idx, PoseDetection& out_pose) {
        detail::cu_check(cudaMemcpyAsync(d_idx_, &track_idx, sizeof(int), cudaMemcpyHostToDevice, stream_), "upload");
        extractPosesToDeviceAsync(d_det_, d_idx_, 1, stream_);
        float det[51];
        detail::cu_check(cudaMemcpyAsync(det, d_det_, sizeof(det), cudaMemcpyDeviceToHost, stream_), "download");
        sync(stream_);
        unflatten(det, out_pose);
    }
    void getAllPredictedPoses(PoseDetection* out_poses, int num_tracks) {
        num_tracks = clamp(num_tracks);
        if (num_tracks == 0) return;
        std::vector<int> slots((size_t)num_tracks);
        for (int i = 0; i < num_tracks; ++i) slots[i] = i;
        std::vector<float> flat((size_t)num_tracks * 51);
        detail::cu_check(cudaMemcpyAsync(d_idx_, slots.data(), slots.size() * sizeof(int), cudaMemcpyHostToDevice, stream_), "upload");
        extractPosesToDeviceAsync(d_det_, d_idx_, num_tracks, stream_);
        detail::cu_check(cudaMemcpyAsync(flat.data(), d_det_, flat.size() * sizeof(float), cudaMemcpyDeviceToHost, stream_), "download");
        sync(stream_);
        for (int i = 0; i < num_tracks; ++i) unflatten(&flat[(size_t)i * 51], out_poses[i]);
    }
    // mean [136], covariance [136*136] (row-major, zero off the diagonal); either may be null.
    void getState(int track_idx, float* mean, float* covariance) {
        if (track_idx < 0 || track_idx >= max_tracks_) return;
        sync(stream_);
        if (mean) detail::cu_check(cudaMemcpy(mean, d_means_ + (size_t)track_idx * TOTAL_STATE_DIM, TOTAL_STATE_DIM * sizeof(float), cudaMemcpyDeviceToHost), "download");
        if (covariance) {
            float diag[TOTAL_STATE_DIM];
            detail::cu_check(cudaMemcpy(diag, d_diag_ + (size_t)track_idx * TOTAL_STATE_DIM, sizeof(diag), cudaMemcpyDeviceToHost), "download");
            for (int i = 0; i < TOTAL_STATE_DIM * TOTAL_STATE_DIM; ++i) covariance[i] = 0.0f;
            for (int i = 0; i < TOTAL_STATE_DIM; ++i) covariance[i * TOTAL_STATE_DIM + i] = diag[i];
        }
    }
    void resetTrack(int track_idx) {
        if (track_idx < 0 || track_idx >= max_tracks_) return;
        detail::cu_check(cudaMemsetAsync(d_means_ + (size_t)track_idx * TOTAL_STATE_DIM, 0, TOTAL_STATE_DIM * sizeof(float), stream_), "memset");
        detail::cu_check(cudaMemsetAsync(d_diag_ + (size_t)track_idx * TOTAL_STATE_DIM, 0, TOTAL_STATE_DIM * sizeof(float), stream_), "memset");
        sync(stream_);
    }

    float* getMeansDevice() { return d_means_; }                   // [max_tracks, 136]
    float* getCovarianceDiagonalsDevice() { return d_diag_; }      // [max_tracks, 136]
    // Full [max_tracks, 136, 136] matrices, materialised from the diagonals on every call.
    float* getCovariancesDevice() {
        const size_t per = (size_t)TOTAL_STATE_DIM * TOTAL_STATE_DIM;
        if (!d_cov_) detail::cu_check(cudaMalloc(&d_cov_, (size_t)max_tracks_ * per * sizeof(float)), "cudaMalloc");
        for (int t = 0; t < max_tracks_; ++t)
            detail::pb_check(pb_kf3_materialize_cov(d_diag_, t, d_cov_ + (size_t)t * per, detail::as_pb(stream_)), "pb_kf3_materialize_cov");
        sync(stream_);
        return d_cov_;
    }
    int getMaxTracks() const { return max_tracks_; }
    cudaStream_t getStream() const { return stream_; }
    void setMotionParams(float accel_memory, float jerk_memory) { accel_memory_ = accel_memory; jerk_memory_ = jerk_memory; }

private:
    int clamp(int n) const { return n < 0 ? 0 : (n > max_tracks_ ? max_tracks_ : n); }
    static void flatten(const PoseDetection& d, float* out) {
        for (int k = 0; k < NUM_KEYPOINTS; ++k) { out[k * 3] = d.keypoints[k].x; out[k * 3 + 1] = d.keypoints[k].y; out[k * 3 + 2] = d.keypoints[k].confidence; }
    }
    // only the keypoints are written, like upstream (kalman_filter.cu:590-596): confidence 1.0
    static void unflatten(const float* in, PoseDetection& d) {
        for (int k = 0; k < NUM_KEYPOINTS; ++k) d.keypoints[k] = Keypoint{in[k * 3], in[k * 3 + 1], in[k * 3 + 2]};
    }
    int max_tracks_;
    float accel_memory_ = 0.9f, jerk_memory_ = 0.9f;
    float *d_means_ = nullptr, *d_diag_ = nullptr, *d_det_ = nullptr, *d_cov_ = nullptr;
    int* d_idx_ = nullptr;
    cudaStream_t stream_ = nullptr;
};

}  // namespace cuda
}  // namespace posebyte
