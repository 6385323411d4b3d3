// include/cuda/gpu_tracker.h — posebyte::cuda::GPUTracker as a one-stream view of the B200
// C ABI (include/posebyte_b200.h).
//
// Public surface of the reference class (reference include/cuda/gpu_tracker.h:16-110):
// GPUTrackerConfig with the same fields and defaults, update(d_det_poses, d_det_scores, n,
// frame_id), getActiveTracks(), TrackerTiming, the device accessors and the CUDA-graph stubs.
// main.cpp:132-140,214-224 and benchmark.cpp:231-284 compile against it unchanged.  One call of
// update() is one kernel launch instead of ~490 stream operations; the results follow the
// reference's arithmetic with its races resolved by rules R3-R5 of DESIGN.md.
#pragma once

#include <cstdio>
#include <vector>

#include "pb_shim_common.h"

namespace posebyte {
namespace cuda {

struct GPUTrackerConfig {
    int max_tracks = 128;
    int max_detections = 64;          // update() truncates its input to this many detections
    float match_threshold = 0.5f;     // inert upstream (hungarian.cu:358-405 never reads it); kept for API parity
    float high_thresh = 0.30f;        // inert upstream (masks computed, never read)
    float low_thresh = 0.15f;         // inert upstream
    float new_track_thresh = 0.30f;   // an unmatched detection starts a track at or above this score
    int max_age = 10;                 // confirmed -> lost after this many missed frames; removed 10 frames later
    int min_hits = 3;                 // tentative -> confirmed
    bool use_cuda_graph = false;      // unused upstream as well
};

struct TrackerTiming {                // microseconds summed over frames, from device timestamps
    long long predict_us = 0, gate_us = 0, high_assoc_us = 0, low_assoc_us = 0, lost_assoc_us = 0;
    long long update_us = 0, age_us = 0, new_track_us = 0, dedup_us = 0, total_us = 0;
    int frame_count = 0;
};

struct GPUTrackState { int track_id, state, hits, age, last_frame; };   // state: 0 tentative, 1 confirmed, 2 lost

class GPUTracker {
public:
    explicit GPUTracker(const GPUTrackerConfig& config = GPUTrackerConfig()) : cfg_(config), h_(to_pb(config)) {
        v_ = h_.views();
        detail::cu_check(cudaStreamCreate(&stream_), "cudaStreamCreate");
        detail::cu_check(cudaMalloc(&d_num_, sizeof(int)), "cudaMalloc");
    }
    ~GPUTracker() {
        cudaFree(d_num_);
        cudaStreamDestroy(stream_);
    }
    GPUTracker(const GPUTracker&) = delete;
    GPUTracker& operator=(const GPUTracker&) = delete;

    // d_det_poses [n,17,3] and d_det_scores [n] are device pointers in score order (what
    // GPUPostprocess leaves behind).  Returns the number of live track slots, lost and
    // tentative ones included, like the reference (gpu_tracker.cu:1130-1136).  Blocking.
    int update(const float* d_det_poses, const float* d_det_scores, int num_detections, int frame_id) {
        if (num_detections < 0) num_detections = 0;
        detail::cu_check(cudaMemcpyAsync(d_num_, &num_detections, sizeof(int), cudaMemcpyHostToDevice, stream_), "upload n");
        const int stride = num_detections > 0 ? num_detections : 1;
        detail::pb_check(pb_tracker_update(h_.get(), d_det_poses ? d_det_poses : v_.det_poses, d_det_scores ? d_det_scores : v_.det_scores,
                                           d_num_, stride, frame_id, detail::as_pb(stream_)), "pb_tracker_update");
        detail::cu_check(cudaMemcpyAsync(&num_active_, v_.num_active + 3, sizeof(int), cudaMemcpyDeviceToHost, stream_), "read active");
        detail::cu_check(cudaStreamSynchronize(stream_), "cudaStreamSynchronize");
        return num_active_;
    }

    // Tracks matched in the last update that are confirmed (or tentative with enough hits) and
    // not lost: id, detection score, smoothed keypoints, keypoint box padded by 10 %
    // (gpu_tracker.cu:1559-1639).  Assembled on the device; one copy back.
    std::vector<TrackOutput> getActiveTracks() {
        std::vector<TrackOutput> out((size_t)cfg_.max_detections);
        int n = 0;
        detail::pb_check(pb_get_tracks(h_.get(), 0, out.data(), cfg_.max_detections, &n), "pb_get_tracks");
        out.resize((size_t)n);
        return out;
    }

    int getNumActiveTracks() const { return num_active_; }

    const TrackerTiming& getTiming() {
        pb_timing t{};
        detail::pb_check(pb_get_timing(h_.get(), &t), "pb_get_timing");
        timing_.predict_us = t.predict_us; timing_.gate_us = t.gate_us; timing_.high_assoc_us = t.high_assoc_us;
        timing_.low_assoc_us = t.low_assoc_us; timing_.lost_assoc_us = t.lost_assoc_us; timing_.update_us = t.update_us;
        timing_.age_us = t.age_us; timing_.new_track_us = t.new_track_us; timing_.dedup_us = t.dedup_us;
        timing_.total_us = t.total_us; timing_.frame_count = t.frame_count;
        return timing_;
    }
    void printTimingStats() {
        const TrackerTiming& t = getTiming();
        const double n = t.frame_count > 0 ? t.frame_count : 1;
        std::printf("tracker stages, mean us/frame over %d frames (device timestamps):\n"
                    "  predict %.2f | gate %.2f | tier1 %.2f | tier2 %.2f | tier3 %.2f | update %.2f | age %.2f | new %.2f | dedup %.2f | total %.2f\n",
                    t.frame_count, t.predict_us / n, t.gate_us / n, t.high_assoc_us / n, t.low_assoc_us / n, t.lost_assoc_us / n,
                    t.update_us / n, t.age_us / n, t.new_track_us / n, t.dedup_us / n, t.total_us / n);
    }

    // CUDA-graph hooks: stubs upstream (gpu_tracker.cu:1660-1667); one launch per update here
    // leaves nothing to capture.
    void captureGraph() {}
    void executeGraph() {}
    bool isGraphCaptured() const { return false; }

    float* getTrackPosesDevice() { return v_.track_poses; }     // [max_tracks, 17, 3]
    float* getTrackScoresDevice() { return v_.track_scores; }   // [max_tracks]
    int* getTrackStatesDevice() { return v_.track_states; }     // [max_tracks]
    int* getTrackIdsDevice() { return v_.track_ids; }           // [max_tracks]
    cudaStream_t getStream() const { return stream_; }
    pb_handle_t handle() const { return h_.get(); }

private:
    static pb_config to_pb(const GPUTrackerConfig& g) {
        pb_config c{};
        pb_default_config(&c);
        c.num_streams = 1;
        c.num_anchors = 64; c.max_candidates = 64; c.max_keep = 64;    // decode half of the handle unused here
        c.max_tracks = g.max_tracks; c.max_detections = g.max_detections;
        c.match_threshold = g.match_threshold; c.high_thresh = g.high_thresh; c.low_thresh = g.low_thresh;
        c.new_track_thresh = g.new_track_thresh; c.max_age = g.max_age; c.min_hits = g.min_hits;
        c.use_cuda_graph = g.use_cuda_graph ? 1 : 0;
        return c;
    }
    GPUTrackerConfig cfg_;
    detail::Handle h_;
    pb_device_views v_{};
    cudaStream_t stream_ = nullptr;
    int* d_num_ = nullptr;
    int num_active_ = 0;
    TrackerTiming timing_;
};

}  // namespace cuda
}  // namespace posebyte
