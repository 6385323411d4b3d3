// include/types.h — ABI types of the PoseBYTE post-inference path.
//
// Drop-in for the reference's include/types.h: the same names, field order and byte
// layout (reference include/types.h:17-174), so code written against the reference's
// structs (main.cpp:207-227, benchmark.cpp:19-66) compiles and links unchanged.
// Sizes are pinned by the static_asserts at the bottom (12 / 224 / 228 / 248 bytes).
#pragma once

#include <cstdint>
#include <cmath>
#include <vector>

#if defined(__CUDACC__)
#define PB_TYPES_HD __host__ __device__
#else
#define PB_TYPES_HD
#endif

namespace posebyte {

// COCO-17 keypoint order (reference types.h:17-36).
enum CocoKeypoint {
    NOSE, LEFT_EYE, RIGHT_EYE, LEFT_EAR, RIGHT_EAR,
    LEFT_SHOULDER, RIGHT_SHOULDER, LEFT_ELBOW, RIGHT_ELBOW, LEFT_WRIST, RIGHT_WRIST,
    LEFT_HIP, RIGHT_HIP, LEFT_KNEE, RIGHT_KNEE, LEFT_ANKLE, RIGHT_ANKLE,
    NUM_KEYPOINTS
};
static_assert(NUM_KEYPOINTS == 17, "COCO-17");

// Per-keypoint OKS falloff (reference types.h:40-58).
constexpr float COCO_SIGMAS[NUM_KEYPOINTS] = {
    0.026f, 0.025f, 0.025f, 0.035f, 0.035f, 0.079f, 0.079f, 0.072f, 0.072f,
    0.062f, 0.062f, 0.107f, 0.107f, 0.087f, 0.087f, 0.089f, 0.089f};

struct Keypoint { float x, y, confidence; };   // types.h:61-65

// One decoded detection (types.h:68-106).
struct PoseDetection {
    float bbox[4];                       // x1 y1 x2 y2
    float score;
    Keypoint keypoints[NUM_KEYPOINTS];

    // Area of the box spanned by keypoints with confidence > 0 (0 if fewer than two).
    PB_TYPES_HD float getPoseArea() const {
        float lo_x = 1e9f, lo_y = 1e9f, hi_x = -1e9f, hi_y = -1e9f;
        int n = 0;
        for (const Keypoint& k : keypoints) {
            if (!(k.confidence > 0.0f)) continue;
            lo_x = fminf(lo_x, k.x); hi_x = fmaxf(hi_x, k.x);
            lo_y = fminf(lo_y, k.y); hi_y = fmaxf(hi_y, k.y);
            ++n;
        }
        return n < 2 ? 0.0f : (hi_x - lo_x) * (hi_y - lo_y);
    }
    // Vertical extent of those keypoints.
    PB_TYPES_HD float getPoseHeight() const {
        float lo_y = 1e9f, hi_y = -1e9f;
        for (const Keypoint& k : keypoints)
            if (k.confidence > 0.0f) { lo_y = fminf(lo_y, k.y); hi_y = fmaxf(hi_y, k.y); }
        return hi_y - lo_y;
    }
};

enum class TrackState { New = 0, Tracked = 1, Lost = 2, Removed = 3 };   // types.h:109-114

// Third-order keypoint filter dimensions (types.h:120-123).
constexpr int MOTION_ORDERS = 4;
constexpr int COORDS_PER_KP = 2;
constexpr int STATE_DIM_PER_KP = MOTION_ORDERS * COORDS_PER_KP;
constexpr int TOTAL_STATE_DIM = NUM_KEYPOINTS * STATE_DIM_PER_KP;         // 136

struct KalmanState {                                                      // types.h:126-132
    float mean[TOTAL_STATE_DIM];
    float covariance[TOTAL_STATE_DIM * TOTAL_STATE_DIM];
    void initFromDetection(const PoseDetection& det);  // declared, never defined upstream
};

// Legacy configuration block kept for source compatibility (types.h:135-155); the
// live tracker is configured through cuda::GPUTrackerConfig.
struct TrackerConfig {
    float high_thresh = 0.6f;
    float low_thresh = 0.1f;
    float new_track_thresh = 0.7f;
    int max_time_lost = 30;
    int min_hits = 3;
    float match_thresh = 0.8f;
    float iou_thresh = 0.3f;
    float accel_memory = 0.9f;
    float jerk_memory = 0.9f;
    float nms_thresh = 0.65f;
};

struct Track {                                                            // types.h:158-166
    int id;
    TrackState state;
    PoseDetection pose;
    int age;
    int hits;
    int time_lost;
    float score;
};

// What GPUTracker::getActiveTracks() returns per visible track (types.h:169-174).
struct TrackOutput {
    int track_id;
    float score;
    float bbox[4];
    Keypoint keypoints[NUM_KEYPOINTS];
};

static_assert(sizeof(Keypoint) == 12, "Keypoint layout");
static_assert(sizeof(PoseDetection) == 224, "PoseDetection layout");
static_assert(sizeof(TrackOutput) == 228, "TrackOutput layout");
static_assert(sizeof(Track) == 248, "Track layout");
static_assert(sizeof(TrackerConfig) == 40, "TrackerConfig layout");

}  // namespace posebyte
