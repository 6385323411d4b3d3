// include/cuda/gpu_postprocess.h — posebyte::cuda::GPUPostprocess as a one-stream view of
// the B200 C ABI (include/posebyte_b200.h).
//
// Same public surface and blocking behaviour as the reference class
// (reference include/cuda/gpu_postprocess.h:11-72, src/cuda/gpu_postprocess.cu:319-503), so
// YoloPoseEngine::detectGPUNative (yolo_pose_engine.cpp:610-646) and main.cpp:207-221 compile
// against it unchanged.  Differences, all deliberate: candidates are compacted in ascending
// anchor order (rule R1 of DESIGN.md) instead of atomics order; failures throw.
#pragma once

#include <vector>

#include "pb_shim_common.h"

namespace posebyte {
namespace cuda {

class GPUPostprocess {
public:
    explicit GPUPostprocess(int max_detections = 1024, int num_anchors = 8400) : h_(config(max_detections, num_anchors)) {
        v_ = h_.views();
    }

    // Decode + confidence filter + pose-NMS of one raw head tensor [56, num_anchors] (device
    // memory, borrowed).  Returns the number of kept detections; they sit in score order at the
    // front of the buffers below until the next call.  Synchronises `stream` like the reference.
    int process(const float* d_raw_output, float conf_threshold, float nms_threshold, cudaStream_t stream = 0) {
        detail::pb_check(pb_postprocess(h_.get(), d_raw_output, conf_threshold, nms_threshold, detail::as_pb(stream)), "pb_postprocess");
        detail::cu_check(cudaMemcpyAsync(&num_host_, v_.num_keep, sizeof(int), cudaMemcpyDeviceToHost, stream), "read num_keep");
        detail::cu_check(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
        return num_host_;
    }

    float* getDetectionPoses() { return v_.det_poses; }     // [<=256, 17, 3]
    float* getDetectionBboxes() { return v_.det_bboxes; }   // [<=256, 4]  x1 y1 x2 y2
    float* getDetectionScores() { return v_.det_scores; }   // [<=256]
    int* getNumDetections() { return v_.num_keep; }         // device scalar
    int getNumDetectionsHost() {
        detail::cu_check(cudaMemcpy(&num_host_, v_.num_keep, sizeof(int), cudaMemcpyDeviceToHost), "read num_keep");
        return num_host_;
    }
    // Anchor index / candidate slot of every kept detection (the reference keeps the slots in a
    // private buffer, d_keep_indices_).
    int* getKeepAnchorsDevice() { return v_.keep_anchors; }
    int* getKeepSlotsDevice() { return v_.keep_slots; }

    // Kept detections as host records (track_id = -1), gpu_postprocess.cu:505-543.
    std::vector<TrackOutput> getRawDetections(int num_dets) {
        std::vector<TrackOutput> out;
        if (num_dets <= 0) return out;
        std::vector<float> poses((size_t)num_dets * 51), boxes((size_t)num_dets * 4), scores(num_dets);
        int nk = 0, nc = 0;
        detail::pb_check(pb_get_kept(h_.get(), 0, poses.data(), boxes.data(), scores.data(), nullptr, nullptr, num_dets, &nk, &nc), "pb_get_kept");
        if (nk > num_dets) nk = num_dets;
        out.resize(nk);
        for (int i = 0; i < nk; ++i) {
            out[i].track_id = -1;
            out[i].score = scores[i];
            for (int e = 0; e < 4; ++e) out[i].bbox[e] = boxes[(size_t)i * 4 + e];
            for (int k = 0; k < NUM_KEYPOINTS; ++k)
                out[i].keypoints[k] = Keypoint{poses[(size_t)i * 51 + k * 3], poses[(size_t)i * 51 + k * 3 + 1], poses[(size_t)i * 51 + k * 3 + 2]};
        }
        return out;
    }

    void debugDumpDetections(int num_dets) {
        std::vector<TrackOutput> d = getRawDetections(num_dets < 3 ? num_dets : 3);
        for (size_t i = 0; i < d.size(); ++i)
            std::printf("det %zu: score %.3f box [%.1f %.1f %.1f %.1f] nose (%.1f, %.1f, %.2f)\n", i, d[i].score, d[i].bbox[0],
                        d[i].bbox[1], d[i].bbox[2], d[i].bbox[3], d[i].keypoints[0].x, d[i].keypoints[0].y, d[i].keypoints[0].confidence);
    }

    pb_handle_t handle() const { return h_.get(); }

private:
    static pb_config config(int max_detections, int num_anchors) {
        pb_config c{};
        pb_default_config(&c);
        c.num_streams = 1;
        c.num_anchors = num_anchors;
        c.max_candidates = max_detections;
        if (c.max_keep > max_detections) c.max_keep = max_detections;
        c.max_tracks = 1; c.max_detections = 1;          // tracker half of the handle unused here
        return c;
    }
    detail::Handle h_;
    pb_device_views v_{};
    int num_host_ = 0;
};

}  // namespace cuda
}  // namespace posebyte
