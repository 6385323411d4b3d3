// include/cuda/oks_distance.h — posebyte::cuda::OKSDistanceCUDA over the B200 C ABI.
//
// The reference class (reference include/cuda/oks_distance.h:11-115, src/cuda/oks_distance.cu):
// ungated OKS cost (1 - OKS; pose scale = mean of the two keypoint-box areas, floor 1000; both
// confidences > 0.2, fallback > 0.05 when fewer than three keypoints qualify), keypoint-box IoU
// cost (confidence > 0, 10 px margin) and their alpha blend; device-pointer and host-pointer
// variants.  One launch per call here (upstream: up to five and two temporaries).
#pragma once

#include <vector>

#include "pb_shim_common.h"

namespace posebyte {
namespace cuda {

class OKSDistanceCUDA {
public:
    explicit OKSDistanceCUDA(int max_tracks = 256, int max_detections = 256) : max_tracks_(max_tracks), max_detections_(max_detections) {
        detail::cu_check(cudaStreamCreate(&stream_), "cudaStreamCreate");
        detail::cu_check(cudaMalloc(&d_tracks_, (size_t)max_tracks * 51 * sizeof(float)), "cudaMalloc");
        detail::cu_check(cudaMalloc(&d_detections_, (size_t)max_detections * 51 * sizeof(float)), "cudaMalloc");
        detail::cu_check(cudaMalloc(&d_costs_, (size_t)max_tracks * max_detections * sizeof(float)), "cudaMalloc");
        detail::cu_check(cudaMalloc(&d_sigmas_, NUM_KEYPOINTS * sizeof(float)), "cudaMalloc");
        detail::cu_check(cudaMemcpy(d_sigmas_, COCO_SIGMAS, NUM_KEYPOINTS * sizeof(float), cudaMemcpyHostToDevice), "upload");   // oks_distance.cu: ctor
        detail::cu_check(cudaMalloc(&d_track_bboxes_, (size_t)max_tracks * 4 * sizeof(float)), "cudaMalloc");
        detail::cu_check(cudaMalloc(&d_det_bboxes_, (size_t)max_detections * 4 * sizeof(float)), "cudaMalloc");
        detail::cu_check(cudaMemset(d_track_bboxes_, 0, (size_t)max_tracks * 4 * sizeof(float)), "cudaMemset");
        detail::cu_check(cudaMemset(d_det_bboxes_, 0, (size_t)max_detections * 4 * sizeof(float)), "cudaMemset");
    }
    ~OKSDistanceCUDA() {
        cudaFree(d_tracks_); cudaFree(d_detections_); cudaFree(d_costs_); cudaFree(d_sigmas_); cudaFree(d_track_bboxes_); cudaFree(d_det_bboxes_);
        cudaStreamDestroy(stream_);
    }
    OKSDistanceCUDA(const OKSDistanceCUDA&) = delete;
    OKSDistanceCUDA& operator=(const OKSDistanceCUDA&) = delete;

    // ---- device pointers, asynchronous (oks_distance.cu:479-537) ----
    void computeOKSDistanceDeviceAsync(const float* d_tracks, const float* d_detections, float* d_out_costs, int num_tracks,
                                       int num_detections, cudaStream_t stream = 0) {
        run(d_tracks, d_detections, d_out_costs, num_tracks, num_detections, 0, 0.0f, stream ? stream : stream_);
    }
    void computeIoUDistanceDeviceAsync(const float* d_track_poses, const float* d_det_poses, float* d_out_costs, int num_tracks,
                                       int num_detections, cudaStream_t stream = 0) {
        run(d_track_poses, d_det_poses, d_out_costs, num_tracks, num_detections, 1, 0.0f, stream ? stream : stream_);
    }
    // new: the alpha blend without the host round trip computeCombinedDistance makes upstream
    void computeCombinedDistanceDeviceAsync(const float* d_tracks, const float* d_detections, float* d_out_costs, int num_tracks,
                                            int num_detections, float alpha = 0.7f, cudaStream_t stream = 0) {
        run(d_tracks, d_detections, d_out_costs, num_tracks, num_detections, 2, alpha, stream ? stream : stream_);
    }

    // ---- host pointers, blocking (oks_distance.cu:305-477) ----
    void computeOKSDistance(const PoseDetection* tracks, const PoseDetection* detections, float* out_costs, int num_tracks, int num_detections) {
        host_call(tracks, detections, out_costs, num_tracks, num_detections, 0, 0.0f);
    }
    void computeIoUDistance(const PoseDetection* tracks, const PoseDetection* detections, float* out_costs, int num_tracks, int num_detections) {
        host_call(tracks, detections, out_costs, num_tracks, num_detections, 1, 0.0f);
    }
    void computeCombinedDistance(const PoseDetection* tracks, const PoseDetection* detections, float* out_costs, int num_tracks,
                                 int num_detections, float alpha = 0.7f) {
        host_call(tracks, detections, out_costs, num_tracks, num_detections, 2, alpha);
    }

    float* getCostsDevice() { return d_costs_; }
    float* getTracksDevice() { return d_tracks_; }
    float* getDetectionsDevice() { return d_detections_; }
    // reference oks_distance.h:86-88.  The sigma table is the one the kernels use (COCO_SIGMAS, types.h).  Upstream stages
    // keypoint boxes in the two bbox buffers on the way to the IoU distance; here boxes stay in registers inside
    // pb_pose_distance, so the buffers exist (sized as upstream, zero-filled) but are not written.
    float* getSigmasDevice() { return d_sigmas_; }
    float* getTrackBboxesDevice() { return d_track_bboxes_; }
    float* getDetBboxesDevice() { return d_det_bboxes_; }
    int getMaxTracks() const { return max_tracks_; }
    int getMaxDetections() const { return max_detections_; }
    cudaStream_t getStream() const { return stream_; }

private:
    void run(const float* t, const float* d, float* out, int nt, int nd, int mode, float alpha, cudaStream_t s) {
        if (nt == 0 || nd == 0) return;
        detail::pb_check(pb_pose_distance(t, d, 1, nt, nd, mode, alpha, out, detail::as_pb(s)), "pb_pose_distance");
    }
    void host_call(const PoseDetection* tracks, const PoseDetection* dets, float* out, int nt, int nd, int mode, float alpha) {
        if (nt == 0 || nd == 0) return;
        if (nt > max_tracks_ || nd > max_detections_) throw std::runtime_error("OKSDistanceCUDA: more poses than the object was sized for");
        std::vector<float> ft((size_t)nt * 51), fd((size_t)nd * 51);
        for (int i = 0; i < nt; ++i)
            for (int k = 0; k < NUM_KEYPOINTS; ++k) { ft[i * 51 + k * 3] = tracks[i].keypoints[k].x; ft[i * 51 + k * 3 + 1] = tracks[i].keypoints[k].y; ft[i * 51 + k * 3 + 2] = tracks[i].keypoints[k].confidence; }
        for (int i = 0; i < nd; ++i)
            for (int k = 0; k < NUM_KEYPOINTS; ++k) { fd[i * 51 + k * 3] = dets[i].keypoints[k].x; fd[i * 51 + k * 3 + 1] = dets[i].keypoints[k].y; fd[i * 51 + k * 3 + 2] = dets[i].keypoints[k].confidence; }
        detail::cu_check(cudaMemcpyAsync(d_tracks_, ft.data(), ft.size() * sizeof(float), cudaMemcpyHostToDevice, stream_), "upload");
        detail::cu_check(cudaMemcpyAsync(d_detections_, fd.data(), fd.size() * sizeof(float), cudaMemcpyHostToDevice, stream_), "upload");
        run(d_tracks_, d_detections_, d_costs_, nt, nd, mode, alpha, stream_);
        detail::cu_check(cudaMemcpyAsync(out, d_costs_, (size_t)nt * nd * sizeof(float), cudaMemcpyDeviceToHost, stream_), "download");
        detail::cu_check(cudaStreamSynchronize(stream_), "cudaStreamSynchronize");
    }
    int max_tracks_, max_detections_;
    float *d_tracks_ = nullptr, *d_detections_ = nullptr, *d_costs_ = nullptr;
    float *d_sigmas_ = nullptr, *d_track_bboxes_ = nullptr, *d_det_bboxes_ = nullptr;
    cudaStream_t stream_ = nullptr;
};

}  // namespace cuda
}  // namespace posebyte
