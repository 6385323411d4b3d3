// oracle/posebyte_oracle.cpp — scalar CPU restatement of the PoseBYTE post-inference path.
//
// TEST INFRASTRUCTURE ONLY (see posebyte_oracle.h).  Each function follows one reference
// kernel; the citation is file:line in naveedprojects/yolo-pose-cpp.  Arithmetic goes
// through include/pb_math.h so that the CUDA kernels and this file evaluate identical
// IEEE operations (build with -ffp-contract=off).  Where the reference's result depends
// on thread scheduling, the rule applied here is stated (R1..R6).
//
// Pinning: no golden vectors exist upstream.  See the header for how this file is pinned
// against the reference's own compiled code.
#include "posebyte_oracle.h"
#include "../include/pb_math.h"
#include "../include/types.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <tuple>
#include <vector>

namespace {

using posebyte::COCO_SIGMAS;
using posebyte::PoseDetection;
using posebyte::TrackOutput;

constexpr int KP = 17;
constexpr int ST_TENTATIVE = 0, ST_CONFIRMED = 1, ST_LOST = 2;   // gpu_tracker.cu:23-25

// ---------------------------------------------------------------------------------
// A1  kernelDecodeAndFilter  gpu_postprocess.cu:30-81
// R1: the reference orders candidates by atomicAdd arrival; here ascending anchor index,
// and on overflow the first max_cand anchors win.
// ---------------------------------------------------------------------------------
int decode(const float* raw, int N, float thr, int max_cand,
           float* poses, float* bboxes, float* scores, int* anchors) {
    int c = 0;
    const float* conf_row = raw + 4 * (size_t)N;
    for (int a = 0; a < N && c < max_cand; ++a) {
        float conf = conf_row[a];
        if (conf < thr) continue;                                  // :51
        float cx = raw[0 * (size_t)N + a], cy = raw[1 * (size_t)N + a];
        float w = raw[2 * (size_t)N + a], h = raw[3 * (size_t)N + a];
        bboxes[c * 4 + 0] = cx - w * 0.5f;                         // :66-69
        bboxes[c * 4 + 1] = cy - h * 0.5f;
        bboxes[c * 4 + 2] = cx + w * 0.5f;
        bboxes[c * 4 + 3] = cy + h * 0.5f;
        scores[c] = conf;
        anchors[c] = a;
        for (int k = 0; k < KP * 3; ++k)                           // :75-80 verbatim copy
            poses[c * 51 + k] = raw[(size_t)(5 + k) * N + a];
        ++c;
    }
    return c;
}

// ---------------------------------------------------------------------------------
// A2  kernelComputeNMSMask  gpu_postprocess.cu:88-172 — one (i,j) pair
// ---------------------------------------------------------------------------------
bool nms_overlap(const float* poses, const float* bboxes, int i, int j,
                 float iou_thr, float oks_thr) {
    float xi1 = bboxes[i * 4 + 0], yi1 = bboxes[i * 4 + 1], xi2 = bboxes[i * 4 + 2], yi2 = bboxes[i * 4 + 3];
    float xj1 = bboxes[j * 4 + 0], yj1 = bboxes[j * 4 + 1], xj2 = bboxes[j * 4 + 2], yj2 = bboxes[j * 4 + 3];
    float ix1 = pb_max(xi1, xj1), iy1 = pb_max(yi1, yj1);
    float ix2 = pb_min(xi2, xj2), iy2 = pb_min(yi2, yj2);
    float iw = pb_max(0.0f, ix2 - ix1), ih = pb_max(0.0f, iy2 - iy1);
    float inter = iw * ih;
    float area_i = (xi2 - xi1) * (yi2 - yi1);
    float area_j = (xj2 - xj1) * (yj2 - yj1);
    float uni = area_i + area_j - inter;
    float iou = (uni > 0) ? (inter / uni) : 0.0f;
    if (iou > iou_thr) return true;                                // :134

    float scale_sq = pb_max(area_i, area_j);                       // :140-141
    if (scale_sq < 32.0f * 32.0f) scale_sq = 32.0f * 32.0f;
    float sum = 0.0f;
    int cnt = 0;
    for (int k = 0; k < KP; ++k) {
        float ci = poses[i * 51 + k * 3 + 2], cj = poses[j * 51 + k * 3 + 2];
        if (ci > 0.2f && cj > 0.2f) {
            float dx = poses[i * 51 + k * 3 + 0] - poses[j * 51 + k * 3 + 0];
            float dy = poses[i * 51 + k * 3 + 1] - poses[j * 51 + k * 3 + 1];
            float d2 = dx * dx + dy * dy;
            float s = COCO_SIGMAS[k];
            sum += pb_expf(-d2 / (2.0f * scale_sq * 4.0f * s * s));  // :156
            ++cnt;
        }
    }
    if (cnt >= 3) {
        float oks = sum / cnt;                                     // :163 (int -> float)
        if (oks > oks_thr || (oks > 0.4f && iou > 0.2f)) return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------
// A3  kernelSortByScore :178-203 (stable insertion sort, R2) + kernelApplyNMSMask
// :209-242 (greedy sweep, stops at max_keep=256) + kernelCompactDetections/CopyBack
// :248-313.  The mask is evaluated lazily: bit(i,j) is only read for kept i and live j,
// and the mask is symmetric, so evaluating it on demand yields the same sweep.
// ---------------------------------------------------------------------------------
int nms_native(float* poses, float* bboxes, float* scores, int C, float thr, int max_keep,
               int* keep_slots) {
    if (C == 0) return 0;                                          // :399
    std::vector<int> order(C);
    for (int i = 0; i < C; ++i) order[i] = i;
    for (int i = 1; i < C; ++i) {                                  // :192-202
        int key = order[i];
        float ks = scores[key];
        int j = i - 1;
        while (j >= 0 && scores[order[j]] < ks) { order[j + 1] = order[j]; --j; }
        order[j + 1] = key;
    }
    std::vector<char> sup(C, 0);
    int nk = 0;
    for (int r = 0; r < C && nk < max_keep; ++r) {                 // :224
        int i = order[r];
        if (sup[i]) continue;
        keep_slots[nk++] = i;
        for (int j = 0; j < C; ++j)                                // :238-240 (row OR)
            if (j != i && !sup[j] && nms_overlap(poses, bboxes, i, j, thr, thr)) sup[j] = 1;
    }
    if (nk == 0) return 0;
    std::vector<float> tp((size_t)nk * 51), tb((size_t)nk * 4), ts(nk);   // :248-313
    for (int k = 0; k < nk; ++k) {
        int s = keep_slots[k];
        std::memcpy(&tp[(size_t)k * 51], &poses[(size_t)s * 51], 51 * sizeof(float));
        std::memcpy(&tb[(size_t)k * 4], &bboxes[(size_t)s * 4], 4 * sizeof(float));
        ts[k] = scores[s];
    }
    std::memcpy(poses, tp.data(), tp.size() * sizeof(float));
    std::memcpy(bboxes, tb.data(), tb.size() * sizeof(float));
    std::memcpy(scores, ts.data(), ts.size() * sizeof(float));
    return nk;
}

// ---------------------------------------------------------------------------------
// A4  NMSCuda::apply  nms.cu:142-306 (host-legacy rules).  std::sort is unstable
// upstream; ties are broken here by original index (callers avoid exact score ties).
// ---------------------------------------------------------------------------------
float legacy_iou(const float* a, const float* b) {                 // nms.cu:166-181
    float x1 = pb_max(a[0], b[0]), y1 = pb_max(a[1], b[1]);
    float x2 = pb_min(a[2], b[2]), y2 = pb_min(a[3], b[3]);
    float iw = pb_max(0.0f, x2 - x1), ih = pb_max(0.0f, y2 - y1);
    float inter = iw * ih;
    float a1 = (a[2] - a[0]) * (a[3] - a[1]);
    float a2 = (b[2] - b[0]) * (b[3] - b[1]);
    float uni = a1 + a2 - inter;
    return (uni > 0) ? (inter / uni) : 0.0f;
}

float legacy_oks(const PoseDetection& p, const PoseDetection& q) { // nms.cu:184-234
    float lx1 = 1e9f, ly1 = 1e9f, hx1 = -1e9f, hy1 = -1e9f;
    float lx2 = 1e9f, ly2 = 1e9f, hx2 = -1e9f, hy2 = -1e9f;
    int v1 = 0, v2 = 0;
    for (int k = 0; k < KP; ++k) {
        if (p.keypoints[k].confidence > 0.2f) {
            lx1 = pb_min(lx1, p.keypoints[k].x); ly1 = pb_min(ly1, p.keypoints[k].y);
            hx1 = pb_max(hx1, p.keypoints[k].x); hy1 = pb_max(hy1, p.keypoints[k].y);
            ++v1;
        }
        if (q.keypoints[k].confidence > 0.2f) {
            lx2 = pb_min(lx2, q.keypoints[k].x); ly2 = pb_min(ly2, q.keypoints[k].y);
            hx2 = pb_max(hx2, q.keypoints[k].x); hy2 = pb_max(hy2, q.keypoints[k].y);
            ++v2;
        }
    }
    if (v1 < 3 || v2 < 3) return 0.0f;
    float a1 = (hx1 - lx1) * (hy1 - ly1), a2 = (hx2 - lx2) * (hy2 - ly2);
    float scale_sq = pb_max(a1, a2);
    if (scale_sq < 32.0f * 32.0f) scale_sq = 32.0f * 32.0f;
    float sum = 0.0f;
    int cnt = 0;
    for (int k = 0; k < KP; ++k) {
        if (p.keypoints[k].confidence > 0.2f && q.keypoints[k].confidence > 0.2f) {
            float dx = p.keypoints[k].x - q.keypoints[k].x;
            float dy = p.keypoints[k].y - q.keypoints[k].y;
            float d2 = dx * dx + dy * dy;
            float s = COCO_SIGMAS[k];
            sum += pb_expf(-d2 / (2.0f * scale_sq * 4.0f * s * s));
            ++cnt;
        }
    }
    return (cnt >= 3) ? (sum / cnt) : 0.0f;
}

bool legacy_suppresses(const PoseDetection& a, const PoseDetection& b) {   // nms.cu:259-301
    float iou = legacy_iou(a.bbox, b.bbox);
    if (iou > 0.55f) return true;
    float oks = legacy_oks(a, b);
    if (oks > 0.5f) return true;
    if (iou > 0.2f && oks > 0.4f) return true;
    float cx1 = (a.bbox[0] + a.bbox[2]) / 2.0f, cy1 = (a.bbox[1] + a.bbox[3]) / 2.0f;
    float cx2 = (b.bbox[0] + b.bbox[2]) / 2.0f, cy2 = (b.bbox[1] + b.bbox[3]) / 2.0f;
    float w1 = a.bbox[2] - a.bbox[0], h1 = a.bbox[3] - a.bbox[1];
    float scale = pb_max(w1, h1);
    if (scale < 32.0f) scale = 32.0f;
    float ddx = cx1 - cx2, ddy = cy1 - cy2;
    float dist = sqrtf(ddx * ddx + ddy * ddy);
    float nd = dist / scale;
    return nd < 0.3f && oks > 0.15f;
}

int nms_legacy(const PoseDetection* dets, int n, float /*oks_thr ignored :142*/,
               float score_thr, int* keep) {
    std::vector<int> idx;
    for (int i = 0; i < n; ++i)
        if (dets[i].score >= score_thr) idx.push_back(i);          // :153-157
    std::stable_sort(idx.begin(), idx.end(),
                     [&](int a, int b) { return dets[a].score > dets[b].score; });
    int m = (int)idx.size(), nk = 0;
    std::vector<char> sup(m, 0);
    for (int i = 0; i < m; ++i) {                                  // :246-303
        if (sup[i]) continue;
        keep[nk++] = idx[i];
        for (int j = i + 1; j < m; ++j)
            if (!sup[j] && legacy_suppresses(dets[idx[i]], dets[idx[j]])) sup[j] = 1;
    }
    return nk;
}

// ---------------------------------------------------------------------------------
// launchPoseNMS (nms.h:48-60 declares it; no definition upstream).  Semantics chosen:
// pairwise OKS exactly as the otherwise unused kernelComputeOKSMatrix (nms.cu:25-117),
// score filter, stable descending order, greedy suppression at OKS > oks_thr.
// ---------------------------------------------------------------------------------
float pose_oks(const float* poses, const float* sig, int i, int j, int nk) {
    const float* a = poses + (size_t)i * nk * 3;
    const float* b = poses + (size_t)j * nk * 3;
    float lxi = 1e9f, lyi = 1e9f, hxi = -1e9f, hyi = -1e9f;
    float lxj = 1e9f, lyj = 1e9f, hxj = -1e9f, hyj = -1e9f;
    int vi = 0, vj = 0;
    for (int k = 0; k < nk; ++k) {
        if (a[k * 3 + 2] > 0.2f) {
            lxi = pb_min(lxi, a[k * 3]); lyi = pb_min(lyi, a[k * 3 + 1]);
            hxi = pb_max(hxi, a[k * 3]); hyi = pb_max(hyi, a[k * 3 + 1]); ++vi;
        }
        if (b[k * 3 + 2] > 0.2f) {
            lxj = pb_min(lxj, b[k * 3]); lyj = pb_min(lyj, b[k * 3 + 1]);
            hxj = pb_max(hxj, b[k * 3]); hyj = pb_max(hyj, b[k * 3 + 1]); ++vj;
        }
    }
    float ai = (hxi - lxi) * (hyi - lyi), aj = (hxj - lxj) * (hyj - lyj);
    float scale_sq = pb_max(ai, aj);
    if (scale_sq < 32.0f * 32.0f || vi < 3 || vj < 3) return 0.0f;  // nms.cu:81-85
    float sum = 0.0f;
    int cnt = 0;
    for (int k = 0; k < nk; ++k) {
        if (a[k * 3 + 2] > 0.2f && b[k * 3 + 2] > 0.2f) {
            float dx = a[k * 3] - b[k * 3], dy = a[k * 3 + 1] - b[k * 3 + 1];
            float d2 = dx * dx + dy * dy;
            float s = sig[k];
            sum += pb_expf(-d2 / (2.0f * scale_sq * 4.0f * s * s));
            ++cnt;
        }
    }
    return (cnt >= 3) ? (sum / (float)cnt) : 0.0f;
}

void pose_nms(const float* poses, const float* scores, const float* sig, int* keep,
              int n, int nk, float oks_thr, float score_thr) {
    std::vector<int> idx;
    for (int i = 0; i < n; ++i) {
        keep[i] = 0;
        if (scores[i] >= score_thr) idx.push_back(i);
    }
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return scores[a] > scores[b]; });
    int m = (int)idx.size();
    std::vector<char> sup(m, 0);
    for (int i = 0; i < m; ++i) {
        if (sup[i]) continue;
        keep[idx[i]] = 1;
        for (int j = i + 1; j < m; ++j) {
            int lo = std::min(idx[i], idx[j]), hi = std::max(idx[i], idx[j]);
            if (!sup[j] && pose_oks(poses, sig, lo, hi, nk) > oks_thr) sup[j] = 1;
        }
    }
}

// ---------------------------------------------------------------------------------
// A11  LinearAssignmentCUDA::solveDeviceAsyncWithActive  hungarian.cu:358-405 with
// kernelAuctionBidding :27-75 and kernelAuctionAssignment :78-123.
// R6 (tie-break, the reference's own): lowest column on equal value (strict '>' at :63),
// lowest row on equal bid (strict '>' at :100).
// An iteration in which no row bids changes neither prices nor assignments, so every
// later iteration is identical to it; the loop may stop there (epsilon is the only thing
// that still changes and nothing reads it afterwards).  ORC_LITERAL_AUCTION=1 disables
// the early stop so that tests can show both give the same result.
// ---------------------------------------------------------------------------------
bool literal_auction() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("ORC_LITERAL_AUCTION"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}

void auction(const float* cost, int R, int C, int* row, int* col, const int* row_active,
             std::vector<float>& price, std::vector<float>& bestv, std::vector<float>& secv,
             std::vector<int>& bestc, int max_iters = -1) {
    if (R == 0 || C == 0) return;                                  // :368
    for (int r = 0; r < R; ++r) row[r] = -1;                       // :373-375
    for (int c = 0; c < C; ++c) col[c] = -1;
    price.assign(C, 0.0f);
    bestv.resize(R); secv.resize(R); bestc.resize(R);
    float eps = 1.0f / (R + 1);                                    // :378
    int iters = max_iters >= 0 ? max_iters : std::min(R * 3, 50);  // :379 (legacy solve: 3R, :283)
    const bool literal = literal_auction();
    for (int it = 0; it < iters; ++it) {
        bool any = false;
        for (int r = 0; r < R; ++r) {                              // bidding :27-75
            if (row[r] >= 0) { bestc[r] = -1; continue; }
            if (row_active && row_active[r] == 0) { bestc[r] = -1; continue; }
            float bv = -1e9f, sv = -1e9f;
            int bc = -1;
            const float* cr = cost + (size_t)r * C;
            for (int c = 0; c < C; ++c) {
                float v = -cr[c] - price[c];                       // :61
                if (v > bv) { sv = bv; bv = v; bc = c; }
                else if (v > sv) { sv = v; }
            }
            bestv[r] = bv; bestc[r] = bc; secv[r] = sv;
            any |= (bc >= 0);
        }
        if (any) {
            for (int c = 0; c < C; ++c) {                          // assignment :78-123
                float hb = -1e9f;
                int hr = -1;
                for (int r = 0; r < R; ++r) {
                    if (bestc[r] == c) {
                        float bid = bestv[r] - secv[r] + eps;      // :99
                        if (bid > hb) { hb = bid; hr = r; }
                    }
                }
                if (hr >= 0) {
                    int prev = col[c];
                    if (prev >= 0) row[prev] = -1;
                    col[c] = hr;
                    row[hr] = c;
                    price[c] += hb;
                }
            }
        } else if (!literal) {
            break;
        }
        eps *= 0.9f;                                               // :402
    }
}

// ---------------------------------------------------------------------------------
// A5-A16  GPUTracker  gpu_tracker.cu
// Every device buffer of the reference is one vector here with the same flat indexing
// (stride = this frame's D for cost / gate buffers, gpu_tracker.cu:257,348), living
// across frames.  Initial contents: zero (a fresh cudaMalloc is modelled as zeros).
// ---------------------------------------------------------------------------------
struct Tracker {
    orc_tracker_config cfg;
    int T, Dm;
    std::vector<float> poses, vel, scores, predicted, tcent, dcent, cost, det_poses, det_scores;
    std::vector<int> states, ids, hits, ages, last_frame, active;
    std::vector<int> row, col, rowb, colb, gate, lgate, slot_for_det;
    int next_id = 1, hint = 0, D = 0, num_active = 0, frame = 0;   // :987-989
    std::vector<float> a_price, a_bv, a_sv;
    std::vector<int> a_bc;

    explicit Tracker(const orc_tracker_config& c) : cfg(c), T(c.max_tracks), Dm(c.max_detections) {
        poses.assign((size_t)T * 51, 0.f); vel.assign((size_t)T * 34, 0.f); scores.assign(T, 0.f);
        predicted.assign((size_t)T * 51, 0.f); tcent.assign((size_t)T * 4, 0.f);
        dcent.assign((size_t)Dm * 4, 0.f); cost.assign((size_t)T * Dm, 0.f);
        det_poses.assign((size_t)Dm * 51, 0.f); det_scores.assign(Dm, 0.f);
        states.assign(T, 0); ids.assign(T, 0); hits.assign(T, 0); ages.assign(T, 0);
        last_frame.assign(T, 0); active.assign(T, 0);
        row.assign(T, -1); col.assign(Dm, -1); rowb.assign(T, -1); colb.assign(Dm, -1);
        gate.assign((size_t)T * Dm, 0); lgate.assign((size_t)T * Dm, 0); slot_for_det.assign(Dm, -1);
    }

    int count_active() const { int n = 0; for (int t = 0; t < T; ++t) n += (active[t] == 1); return n; }

    // A6 kernelKalmanPredict :102-138 (dt = 1.0f, :1173)
    void predict() {
        const float dt = 1.0f;
        for (int t = 0; t < T; ++t) {
            if (active[t] == 0) continue;
            for (int k = 0; k < KP; ++k) {
                int po = t * 51 + k * 3, vo = t * 34 + k * 2;
                predicted[po + 0] = poses[po + 0] + vel[vo + 0] * dt;
                predicted[po + 1] = poses[po + 1] + vel[vo + 1] * dt;
                predicted[po + 2] = poses[po + 2];
                if (states[t] == ST_LOST) { vel[vo + 0] *= 0.95f; vel[vo + 1] *= 0.95f; }
            }
        }
    }

    // A7 kernelComputeBboxCenters :196-237
    static void centers(const float* p, float* out, int n) {
        for (int i = 0; i < n; ++i) {
            float lx = 1e9f, ly = 1e9f, hx = -1e9f, hy = -1e9f;
            int valid = 0;
            for (int k = 0; k < KP; ++k) {
                if (p[i * 51 + k * 3 + 2] > 0.1f) {
                    float x = p[i * 51 + k * 3], y = p[i * 51 + k * 3 + 1];
                    lx = pb_min(lx, x); ly = pb_min(ly, y); hx = pb_max(hx, x); hy = pb_max(hy, y);
                    ++valid;
                }
            }
            if (valid < 2) { out[i * 4] = out[i * 4 + 1] = out[i * 4 + 2] = out[i * 4 + 3] = 0.f; continue; }
            float w = hx - lx, h = hy - ly;
            out[i * 4 + 0] = (lx + hx) * 0.5f;
            out[i * 4 + 1] = (ly + hy) * 0.5f;
            out[i * 4 + 2] = w;
            out[i * 4 + 3] = h;
        }
    }

    // A8 kernelSpatialGate :241-317
    void spatial_gate(std::vector<int>& g, float base) {
        for (int t = 0; t < T; ++t) {
            for (int d = 0; d < D; ++d) {
                int idx = t * D + d;
                if (active[t] == 0) { g[idx] = 0; continue; }
                float tw = tcent[t * 4 + 2], th = tcent[t * 4 + 3];
                float dw = dcent[d * 4 + 2], dh = dcent[d * 4 + 3];
                if (tw < 1.0f || th < 1.0f || dw < 1.0f || dh < 1.0f) { g[idx] = 1; continue; }
                if (!cfg.gating_enabled) { g[idx] = 1; continue; }   // extension (config 5 "gating off")
                float dx = tcent[t * 4 + 0] - dcent[d * 4 + 0];
                float dy = tcent[t * 4 + 1] - dcent[d * 4 + 1];
                float dist = sqrtf(dx * dx + dy * dy);
                static const int torso[4] = {5, 6, 11, 12};
                float av = 0.0f;
                for (int i = 0; i < 4; ++i) {
                    float vx = vel[t * 34 + torso[i] * 2], vy = vel[t * 34 + torso[i] * 2 + 1];
                    av += sqrtf(vx * vx + vy * vy);
                }
                av *= 0.25f;
                float size = (tw + th + dw + dh) * 0.25f;
                float ratio = dist / (size + 1e-6f);
                float vf = 1.0f + pb_min(av / (size + 1e-6f), 2.0f);
                float thr = base * vf;
                if (states[t] == ST_LOST) thr *= 2.0f;
                g[idx] = (ratio < thr) ? 1 : 0;
            }
        }
    }

    // kernelMaskTracksByState :498-515
    void mask_state(std::vector<int>& g, int st) {
        for (int t = 0; t < T; ++t)
            if (active[t] == 1 && states[t] == st)
                for (int d = 0; d < D; ++d) g[t * D + d] = 0;
    }

    // A9 kernelOKSWithGating :333-425
    void oks_gated(const std::vector<int>& g, float vis) {
        for (int t = 0; t < T; ++t) {
            for (int d = 0; d < D; ++d) {
                int idx = t * D + d;
                if (active[t] == 0) { cost[idx] = 1.0f; continue; }
                if (g[idx] == 0) continue;                         // cell keeps its old value (Q1)
                const float* tp = &predicted[(size_t)t * 51];
                const float* dp = &det_poses[(size_t)d * 51];
                float dlx = 1e9f, dly = 1e9f, dhx = -1e9f, dhy = -1e9f;
                float tlx = 1e9f, tly = 1e9f, thx = -1e9f, thy = -1e9f;
                for (int k = 0; k < KP; ++k) {
                    if (dp[k * 3 + 2] > 0.1f) {
                        dlx = pb_min(dlx, dp[k * 3]); dly = pb_min(dly, dp[k * 3 + 1]);
                        dhx = pb_max(dhx, dp[k * 3]); dhy = pb_max(dhy, dp[k * 3 + 1]);
                    }
                    if (tp[k * 3 + 2] > 0.1f) {
                        tlx = pb_min(tlx, tp[k * 3]); tly = pb_min(tly, tp[k * 3 + 1]);
                        thx = pb_max(thx, tp[k * 3]); thy = pb_max(thy, tp[k * 3 + 1]);
                    }
                }
                float da = (dhx - dlx) * (dhy - dly), ta = (thx - tlx) * (thy - tly);
                float scale_sq = pb_max((da + ta) * 0.5f, 1000.0f);
                float sum = 0.0f;
                int cnt = 0;
                for (int k = 0; k < KP; ++k) {
                    if (dp[k * 3 + 2] > vis && tp[k * 3 + 2] > vis) {
                        float dx = dp[k * 3] - tp[k * 3], dy = dp[k * 3 + 1] - tp[k * 3 + 1];
                        float d2 = dx * dx + dy * dy;
                        float s = COCO_SIGMAS[k] * 2.0f;
                        float s2 = s * s;
                        sum += pb_expf(-d2 / (2.0f * scale_sq * s2));   // :417
                        ++cnt;
                    }
                }
                float oks = (cnt >= 3) ? (sum / cnt) : 0.0f;
                cost[idx] = 1.0f - oks;
            }
        }
    }

    // A10 kernelTorsoOKS :429-490
    void oks_torso(const std::vector<int>& g) {
        static const int torso[4] = {5, 6, 11, 12};
        for (int t = 0; t < T; ++t) {
            for (int d = 0; d < D; ++d) {
                int idx = t * D + d;
                if (active[t] == 0) { cost[idx] = 1.0f; continue; }
                if (g[idx] == 0) continue;
                const float* tp = &predicted[(size_t)t * 51];
                const float* dp = &det_poses[(size_t)d * 51];
                const float scale_sq = 10000.0f;
                float sum = 0.0f;
                int cnt = 0;
                for (int i = 0; i < 4; ++i) {
                    int k = torso[i];
                    if (dp[k * 3 + 2] > 0.1f && tp[k * 3 + 2] > 0.1f) {
                        float dx = dp[k * 3] - tp[k * 3], dy = dp[k * 3 + 1] - tp[k * 3 + 1];
                        float d2 = dx * dx + dy * dy;
                        float s = COCO_SIGMAS[k] * 3.0f;
                        sum += pb_expf(-d2 / (2.0f * scale_sq * s * s));   // :482
                        ++cnt;
                    }
                }
                float oks = (cnt >= 2) ? (sum / cnt) : 0.0f;
                cost[idx] = 1.0f - oks;
            }
        }
    }

    // kernelLockMatchedPairs :540-567
    void lock(std::vector<int>& g) {
        for (int t = 0; t < T; ++t)
            for (int d = 0; d < D; ++d)
                if (row[t] >= 0 || col[d] >= 0) { cost[t * D + d] = 1e9f; g[t * D + d] = 0; }
    }

    void solve() { auction(cost.data(), T, D, row.data(), col.data(), active.data(), a_price, a_bv, a_sv, a_bc); }
    void backup() { rowb = row; std::copy(col.begin(), col.begin() + D, colb.begin()); }
    void merge() {                                                 // kernelMergeAssignments :575-588
        for (int t = 0; t < T; ++t) if (rowb[t] >= 0) row[t] = rowb[t];
        for (int d = 0; d < D; ++d) if (colb[d] >= 0) col[d] = colb[d];
    }

    // A13 kernelKalmanUpdate :141-189 + kernelUpdateMatchedTracks :612-648
    void update_matched() {
        const float process_noise = 0.1f, measurement_noise = 0.3f;   // :1452-1453
        const float K = measurement_noise / (measurement_noise + process_noise);
        const float alpha = 0.3f;
        for (int t = 0; t < T; ++t) {
            if (active[t] == 0) continue;
            int d = row[t];
            if (d < 0) continue;
            for (int k = 0; k < KP; ++k) {
                int to = t * 51 + k * 3, dofs = d * 51 + k * 3, vo = t * 34 + k * 2;
                float ox = poses[to], oy = poses[to + 1];
                float zx = det_poses[dofs], zy = det_poses[dofs + 1], zc = det_poses[dofs + 2];
                float nx = ox + K * (zx - ox);
                float ny = oy + K * (zy - oy);
                float dx = zx - ox, dy = zy - oy;
                vel[vo + 0] = alpha * dx + (1 - alpha) * vel[vo + 0];
                vel[vo + 1] = alpha * dy + (1 - alpha) * vel[vo + 1];
                poses[to] = nx; poses[to + 1] = ny; poses[to + 2] = zc;
            }
        }
        for (int t = 0; t < T; ++t) {
            if (active[t] == 0) continue;
            int d = row[t];
            if (d < 0) continue;
            scores[t] = det_scores[d];
            hits[t]++;
            ages[t] = 0;
            last_frame[t] = frame;
            int st = states[t];
            if (st == ST_TENTATIVE && hits[t] >= cfg.min_hits) states[t] = ST_CONFIRMED;
            else if (st == ST_LOST) states[t] = ST_CONFIRMED;
        }
    }

    // A14 kernelAgeUnmatchedTracks :651-688 (LOST_WINDOW = 10, gpu_tracker.h:119)
    void age_unmatched() {
        const int lost_window = 10;
        for (int t = 0; t < T; ++t) {
            if (active[t] == 0 || row[t] >= 0) continue;
            int age = ++ages[t];
            int st = states[t];
            if (st == ST_TENTATIVE) { if (age > 2) active[t] = 0; }
            else if (st == ST_CONFIRMED) { if (age > cfg.max_age) states[t] = ST_LOST; }
            else if (st == ST_LOST) { if (age > cfg.max_age + lost_window) active[t] = 0; }
        }
    }

    // A15 kernelAllocateNewTrackSlots :695-724 + kernelInitNewTracks :727-780
    // R3: detections take slots in ascending detection order; R4: ids in the same order.
    // Replay mode (orc_tracker_force_new): the slot and id each new detection receives are
    // GIVEN — the outcome the reference's atomics produced in a recorded run — so that the rest
    // of the restatement can be compared with that run although the race fell differently.
    std::vector<int> forced_slot, forced_id;
    int forced_errors = 0;
    void new_tracks() {
        const bool forced = !forced_slot.empty();
        for (int d = 0; d < D; ++d) slot_for_det[d] = -1;
        for (int d = 0; d < D; ++d) {
            if (col[d] >= 0) continue;
            if (det_scores[d] < cfg.new_track_thresh) continue;
            int start = (hint++) % T;
            if (forced) {
                int s = d < (int)forced_slot.size() ? forced_slot[d] : -1;
                if (s >= 0 && s < T && active[s] == 0) { active[s] = 1; slot_for_det[d] = s; }
                else ++forced_errors;
                continue;
            }
            for (int i = 0; i < T; ++i) {
                int s = (start + i) % T;
                if (active[s] == 0) { active[s] = 1; slot_for_det[d] = s; break; }
            }
        }
        for (int d = 0; d < D; ++d) {
            if (col[d] >= 0) continue;
            if (det_scores[d] < cfg.new_track_thresh) continue;
            int s = slot_for_det[d];
            if (s < 0) continue;
            ids[s] = forced ? forced_id[d] : next_id;
            ++next_id;
            scores[s] = det_scores[d];
            hits[s] = 1; ages[s] = 0; states[s] = ST_TENTATIVE; last_frame[s] = frame;
            col[d] = s;
            for (int k = 0; k < 51; ++k) poses[s * 51 + k] = det_poses[d * 51 + k];
            for (int k = 0; k < 34; ++k) vel[s * 34 + k] = 0.0f;
        }
    }

    // A16 kernelTrackIoU :788-857 (upper triangle) + kernelRemoveDuplicates :861-895.
    // R5: sequential semantics, t1 ascending, t2 ascending, in-place.
    void dedup() {
        const float thr = 0.7f;                                    // gpu_tracker.h:122
        std::vector<char> elig(T);
        for (int t = 0; t < T; ++t)
            elig[t] = active[t] == 1 && states[t] != ST_LOST && hits[t] >= cfg.min_hits;
        auto iou = [&](int a, int b) -> float {
            if (!elig[a] || !elig[b]) return 0.0f;
            float cx1 = tcent[a * 4], cy1 = tcent[a * 4 + 1], w1 = tcent[a * 4 + 2], h1 = tcent[a * 4 + 3];
            float cx2 = tcent[b * 4], cy2 = tcent[b * 4 + 1], w2 = tcent[b * 4 + 2], h2 = tcent[b * 4 + 3];
            float x1a = cx1 - w1 * 0.5f, x1b = cx1 + w1 * 0.5f, y1a = cy1 - h1 * 0.5f, y1b = cy1 + h1 * 0.5f;
            float x2a = cx2 - w2 * 0.5f, x2b = cx2 + w2 * 0.5f, y2a = cy2 - h2 * 0.5f, y2b = cy2 + h2 * 0.5f;
            float ix1 = pb_max(x1a, x2a), iy1 = pb_max(y1a, y2a);
            float ix2 = pb_min(x1b, x2b), iy2 = pb_min(y1b, y2b);
            float iw = pb_max(0.0f, ix2 - ix1), ih = pb_max(0.0f, iy2 - iy1);
            float inter = iw * ih;
            float a1 = w1 * h1, a2 = w2 * h2;
            float uni = a1 + a2 - inter;
            return (uni > 0) ? (inter / uni) : 0.0f;
        };
        // IoU values are those of the state before any removal (the matrix is complete
        // before kernelRemoveDuplicates starts); only `active` is read live.
        std::vector<std::pair<int, int>> dup;
        for (int a = 0; a < T; ++a)
            for (int b = a + 1; b < T; ++b)
                if (iou(a, b) > thr) dup.push_back({a, b});
        for (auto& pr : dup) {
            int t1 = pr.first, t2 = pr.second;
            if (active[t1] == 0 || states[t1] == ST_LOST) continue;
            if (active[t2] == 0 || states[t2] == ST_LOST) continue;
            if (hits[t1] < hits[t2] || (hits[t1] == hits[t2] && ids[t1] > ids[t2])) active[t1] = 0;
            else active[t2] = 0;
        }
    }

    // GPUTracker::update :1057-1158
    int update(const float* dp, const float* ds, int n, int frame_id) {
        frame = frame_id;
        D = std::min(n, Dm);
        if (D > 0) {
            std::memcpy(det_poses.data(), dp, (size_t)D * 51 * sizeof(float));
            std::memcpy(det_scores.data(), ds, (size_t)D * sizeof(float));
        }
        for (int t = 0; t < T; ++t) row[t] = -1;
        for (int d = 0; d < D; ++d) col[d] = -1;
        num_active = count_active();

        if (num_active > 0) predict();                             // :1160-1175
        if (num_active > 0 && D > 0) {                             // :1177-1208
            centers(predicted.data(), tcent.data(), T);
            centers(det_poses.data(), dcent.data(), D);
            spatial_gate(gate, 3.0f);
        }
        if (num_active > 0 && D > 0) {                             // tier 1 :1210-1274
            mask_state(gate, ST_LOST);
            oks_gated(gate, 0.2f);
            solve();
            lock(gate);
        }
        if (num_active > 0 && D > 0) {                             // tier 2 :1276-1335
            backup();
            oks_torso(gate);
            solve();
            merge();
            lock(gate);
        }
        if (D > 0) {                                               // tier 3 :1337-1436
            backup();
            spatial_gate(lgate, 3.0f * 1.3f);
            mask_state(lgate, ST_CONFIRMED);
            mask_state(lgate, ST_TENTATIVE);
            lock(lgate);
            oks_gated(lgate, 0.2f);
            solve();
            merge();
        }
        if (D > 0) update_matched();                               // :1438-1472
        age_unmatched();                                           // :1474-1487
        if (D > 0) new_tracks();                                   // :1489-1526
        forced_slot.clear(); forced_id.clear();
        dedup();                                                   // :1528-1557
        num_active = count_active();
        return num_active;
    }

    // GPUTracker::getActiveTracks :1559-1639
    int get_tracks(TrackOutput* out, int cap) const {
        int n = 0;
        for (int d = 0; d < D; ++d) {
            int s = col[d];
            if (s < 0) continue;
            if (states[s] == ST_TENTATIVE && hits[s] < cfg.min_hits) continue;
            if (states[s] == ST_LOST) continue;
            if (n >= cap) break;
            TrackOutput& o = out[n++];
            o.track_id = ids[s];
            o.score = det_scores[d];
            float lx = 1e9f, ly = 1e9f, hx = -1e9f, hy = -1e9f;
            for (int k = 0; k < KP; ++k) {
                o.keypoints[k].x = poses[s * 51 + k * 3];
                o.keypoints[k].y = poses[s * 51 + k * 3 + 1];
                o.keypoints[k].confidence = poses[s * 51 + k * 3 + 2];
                if (o.keypoints[k].confidence > 0.2f) {
                    lx = pb_min(lx, o.keypoints[k].x); ly = pb_min(ly, o.keypoints[k].y);
                    hx = pb_max(hx, o.keypoints[k].x); hy = pb_max(hy, o.keypoints[k].y);
                }
            }
            float px = (hx - lx) * 0.1f, py = (hy - ly) * 0.1f;
            o.bbox[0] = lx - px; o.bbox[1] = ly - py; o.bbox[2] = hx + px; o.bbox[3] = hy + py;
        }
        return n;
    }
};

// ---------------------------------------------------------------------------------
// K1-K4  KalmanFilterCUDA  kalman_filter.cu
// ---------------------------------------------------------------------------------
struct KF3 {
    int T;
    std::vector<float> mean, diag;
    explicit KF3(int t) : T(t), mean((size_t)t * 136, 0.f), diag((size_t)t * 136, 0.f) {}

    void initiate(const float* dets, const int* slots, int n) {    // :24-82
        for (int i = 0; i < n; ++i) {
            int t = slots[i];
            for (int k = 0; k < KP; ++k) {
                float x = dets[i * 51 + k * 3], y = dets[i * 51 + k * 3 + 1], c = dets[i * 51 + k * 3 + 2];
                float* m = &mean[(size_t)t * 136 + k * 8];
                float* p = &diag[(size_t)t * 136 + k * 8];
                m[0] = x; m[1] = y;
                for (int j = 2; j < 8; ++j) m[j] = 0.0f;
                float pv = (c > 0.0f) ? 10.0f : 1000.0f;
                p[0] = pv; p[1] = pv;
                for (int j = 2; j < 8; ++j) p[j] = 100.0f;
            }
        }
    }
    void predict(int n, float am, float jm) {                      // :86-167, slots [0,n)
        for (int t = 0; t < n; ++t) {
            for (int k = 0; k < KP; ++k) {
                float* m = &mean[(size_t)t * 136 + k * 8];
                float px = m[0], py = m[1], vx = m[2], vy = m[3], ax = m[4], ay = m[5], jx = m[6], jy = m[7];
                m[0] = px + vx + 0.5f * ax + (1.0f / 6.0f) * jx;
                m[1] = py + vy + 0.5f * ay + (1.0f / 6.0f) * jy;
                m[2] = vx + ax + 0.5f * jx;
                m[3] = vy + ay + 0.5f * jy;
                m[4] = ax * am; m[5] = ay * am;
                m[6] = jx * jm; m[7] = jy * jm;
            }
            for (int i = 0; i < 136; ++i) {
                int ty = i % 8;
                float noise = ty < 2 ? 1.0f : ty < 4 ? 0.5f : ty < 6 ? 0.1f : 0.05f;
                diag[(size_t)t * 136 + i] += noise * noise;
            }
        }
    }
    void update(const float* dets, const int* matches, int n) {    // :171-237
        for (int i = 0; i < n; ++i) {
            int t = matches[i * 2], d = matches[i * 2 + 1];
            for (int k = 0; k < KP; ++k) {
                float zx = dets[d * 51 + k * 3], zy = dets[d * 51 + k * 3 + 1], c = dets[d * 51 + k * 3 + 2];
                if (c < 0.1f) continue;
                float* m = &mean[(size_t)t * 136 + k * 8];
                float* p = &diag[(size_t)t * 136 + k * 8];
                float yx = zx - m[0], yy = zy - m[1];
                float Pxx = p[0], Pyy = p[1];
                float R = 5.0f / (c + 0.1f);
                float Sxx = Pxx + R, Syy = Pyy + R;
                float Kx = Pxx / Sxx, Ky = Pyy / Syy;
                m[0] += Kx * yx;
                m[1] += Ky * yy;
                float Kv = 0.5f * Kx;
                m[2] += Kv * yx;
                m[3] += Kv * yy;
                p[0] = (1.0f - Kx) * Pxx;
                p[1] = (1.0f - Ky) * Pyy;
            }
        }
    }
    void extract(float* out, const int* slots, int n) const {      // :241-264
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < KP; ++k) {
                out[i * 51 + k * 3] = mean[(size_t)slots[i] * 136 + k * 8];
                out[i * 51 + k * 3 + 1] = mean[(size_t)slots[i] * 136 + k * 8 + 1];
                out[i * 51 + k * 3 + 2] = 1.0f;
            }
    }
};

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

void orc_expf_array(const float* x, float* y, int n) {
    for (int i = 0; i < n; ++i) y[i] = pb_expf(x[i]);
}

// Full symmetric NMS overlap matrix (the reference's kernelComputeNMSMask output, one byte
// per bit) for tests that want to check the lazy sweep against the literal one.
void orc_nms_mask(const float* poses, const float* bboxes, int C, float thr, unsigned char* mask) {
    for (int i = 0; i < C; ++i)
        for (int j = 0; j < C; ++j)
            mask[(size_t)i * C + j] = (i != j && nms_overlap(poses, bboxes, i, j, thr, thr)) ? 1 : 0;
}

int orc_decode(const float* raw, int N, float thr, int max_cand, float* poses, float* bboxes,
               float* scores, int* anchors) {
    return decode(raw, N, thr, max_cand, poses, bboxes, scores, anchors);
}

int orc_nms_native(float* poses, float* bboxes, float* scores, int C, float thr, int max_keep,
                   int* keep_slots) {
    return nms_native(poses, bboxes, scores, C, thr, max_keep, keep_slots);
}

int orc_postprocess(const float* raw, int N, float conf_thr, float nms_thr, int max_cand,
                    int max_keep, float* poses, float* bboxes, float* scores, int* keep_slots,
                    int* keep_anchors, int* num_cand) {
    std::vector<float> cp((size_t)max_cand * 51), cb((size_t)max_cand * 4), cs(max_cand);
    std::vector<int> ca(max_cand), ks(max_cand);
    int C = decode(raw, N, conf_thr, max_cand, cp.data(), cb.data(), cs.data(), ca.data());
    if (num_cand) *num_cand = C;
    int K = nms_native(cp.data(), cb.data(), cs.data(), C, nms_thr, max_keep, ks.data());
    std::memcpy(poses, cp.data(), (size_t)K * 51 * sizeof(float));
    std::memcpy(bboxes, cb.data(), (size_t)K * 4 * sizeof(float));
    std::memcpy(scores, cs.data(), (size_t)K * sizeof(float));
    for (int k = 0; k < K; ++k) {
        if (keep_slots) keep_slots[k] = ks[k];
        if (keep_anchors) keep_anchors[k] = ca[ks[k]];
    }
    return K;
}

int orc_nms_legacy(const void* dets, int n, float oks_thr, float score_thr, int* keep) {
    return nms_legacy(static_cast<const PoseDetection*>(dets), n, oks_thr, score_thr, keep);
}

void orc_pose_nms(const float* poses, const float* scores, const float* sigmas, int* keep, int n,
                  int nk, float oks_thr, float score_thr) {
    pose_nms(poses, scores, sigmas, keep, n, nk, oks_thr, score_thr);
}

// ---- OKSDistanceCUDA (oks_distance.cu) and GreedyMatcherCUDA (hungarian.cu) ----

// kernelOKSDistance :26-164 for one (track, detection) pair.
static float pose_oks_cost(const float* tp, const float* dp) {
    float dlx = 1e9f, dly = 1e9f, dhx = -1e9f, dhy = -1e9f;
    int dvalid = 0;
    for (int k = 0; k < KP; ++k)
        if (dp[k * 3 + 2] > 0.1f) {
            dlx = pb_min(dlx, dp[k * 3]); dly = pb_min(dly, dp[k * 3 + 1]);
            dhx = pb_max(dhx, dp[k * 3]); dhy = pb_max(dhy, dp[k * 3 + 1]); ++dvalid;
        }
    float tlx = 1e9f, tly = 1e9f, thx = -1e9f, thy = -1e9f;
    for (int k = 0; k < KP; ++k)
        if (tp[k * 3 + 2] > 0.1f) {
            tlx = pb_min(tlx, tp[k * 3]); tly = pb_min(tly, tp[k * 3 + 1]);
            thx = pb_max(thx, tp[k * 3]); thy = pb_max(thy, tp[k * 3 + 1]);
        }
    float det_scale = (dhx - dlx) * (dhy - dly), trk_scale = (thx - tlx) * (thy - tly);
    float scale_sq = (det_scale + trk_scale) * 0.5f;                     // :78-86
    if (scale_sq < 1000.0f) scale_sq = 1000.0f;
    if (dvalid < 2) return 1.0f;                                         // :89-92
    auto pass = [&](float thr, float& sum, int& cnt) {
        sum = 0.0f; cnt = 0;
        for (int k = 0; k < KP; ++k)
            if (dp[k * 3 + 2] > thr && tp[k * 3 + 2] > thr) {
                float dx = dp[k * 3] - tp[k * 3], dy = dp[k * 3 + 1] - tp[k * 3 + 1];
                float d2 = dx * dx + dy * dy;
                float sg = COCO_SIGMAS[k] * 2.0f;
                sum += pb_expf(-d2 / (2.0f * scale_sq * (sg * sg)));
                ++cnt;
            }
    };
    float sum; int cnt;
    pass(0.2f, sum, cnt);                                                // :101-127
    float oks;
    if (cnt >= 3) oks = sum / (float)cnt;
    else { pass(0.05f, sum, cnt); oks = cnt > 0 ? sum / (float)cnt : 0.0f; }   // :134-160
    return 1.0f - oks;
}

// kernelExtractBboxes :213-245 + kernelIoUDistance :167-210.
static float pose_iou_cost(const float* tp, const float* dp) {
    auto box = [](const float* p, float* b) {
        float lx = 1e9f, ly = 1e9f, hx = -1e9f, hy = -1e9f;
        for (int k = 0; k < KP; ++k)
            if (p[k * 3 + 2] > 0.0f) {
                lx = pb_min(lx, p[k * 3]); ly = pb_min(ly, p[k * 3 + 1]);
                hx = pb_max(hx, p[k * 3]); hy = pb_max(hy, p[k * 3 + 1]);
            }
        b[0] = lx - 10.0f; b[1] = ly - 10.0f; b[2] = hx + 10.0f; b[3] = hy + 10.0f;
    };
    float t[4], d[4];
    box(tp, t); box(dp, d);
    float ix1 = pb_max(t[0], d[0]), iy1 = pb_max(t[1], d[1]), ix2 = pb_min(t[2], d[2]), iy2 = pb_min(t[3], d[3]);
    float iw = pb_max(0.0f, ix2 - ix1), ih = pb_max(0.0f, iy2 - iy1);
    float inter = iw * ih;
    float ta = (t[2] - t[0]) * (t[3] - t[1]), da = (d[2] - d[0]) * (d[3] - d[1]);
    float uni = ta + da - inter;
    float iou = (uni > 0.0f) ? (inter / uni) : 0.0f;
    return 1.0f - iou;
}

// mode 0: OKS cost, 1: IoU cost, 2: alpha * OKS + (1 - alpha) * IoU (kernelCombineCosts :248-261).
void orc_pose_distance(const float* tracks, const float* dets, int nt, int nd, int mode, float alpha, float* out) {
    for (int t = 0; t < nt; ++t)
        for (int d = 0; d < nd; ++d) {
            const float* tp = tracks + (size_t)t * 51;
            const float* dp = dets + (size_t)d * 51;
            float c;
            if (mode == 0) c = pose_oks_cost(tp, dp);
            else if (mode == 1) c = pose_iou_cost(tp, dp);
            else { float o = pose_oks_cost(tp, dp), u = pose_iou_cost(tp, dp); c = alpha * o + (1.0f - alpha) * u; }
            out[(size_t)t * nd + d] = c;
        }
}

// GreedyMatcherCUDA::match host path, hungarian.cu:441-467: cells below the threshold sorted by
// (cost, row, col), taken greedily.  (The device kernel :126-157 races; this is the stated rule.)
void orc_greedy_match(const float* cost, int R, int C, float threshold, int* row_matched) {
    std::vector<std::tuple<float, int, int>> cells;
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c)
            if (cost[(size_t)r * C + c] < threshold) cells.push_back({cost[(size_t)r * C + c], r, c});
    std::sort(cells.begin(), cells.end());
    std::vector<char> ru(R, 0), cu(C, 0);
    for (int r = 0; r < R; ++r) row_matched[r] = -1;
    for (auto& [v, r, c] : cells)
        if (!ru[r] && !cu[c]) { ru[r] = 1; cu[c] = 1; row_matched[r] = c; }
}

void orc_auction(const float* cost, int R, int C, int* row, int* col, const int* row_active) {
    std::vector<float> p, bv, sv;
    std::vector<int> bc;
    auction(cost, R, C, row, col, row_active, p, bv, sv, bc);
}

// LinearAssignmentCUDA::solve, hungarian.cu:235-339 (the host-threshold legacy entry point).
// Fewer than 100 cells: greedyAssign (:198-233) — rows in order, each takes its cheapest unused column
// below the threshold (strict '<': lowest column on ties).  Otherwise the auction with ALL rows active,
// 3*rows iterations at most (:283), the convergence exit of :317 (an iteration without a bid — the same
// fixed point the early stop above uses), then assignments whose cost exceeds the threshold are
// cleared (:328-336).  Returns the number of assignments kept.
int orc_assign_legacy(const float* cost, int R, int C, float threshold, int* row, int* col) {
    if (R == 0 || C == 0) return 0;                                    // :243
    int count = 0;
    if (R * C < 100) {                                                 // :246
        std::vector<char> used(C, 0);
        for (int r = 0; r < R; ++r) row[r] = -1;
        for (int c = 0; c < C; ++c) col[c] = -1;
        for (int r = 0; r < R; ++r) {
            float best = threshold;
            int bc = -1;
            for (int c = 0; c < C; ++c)
                if (!used[c]) { const float v = cost[(size_t)r * C + c]; if (v < best) { best = v; bc = c; } }
            if (bc >= 0) { row[r] = bc; col[bc] = r; used[bc] = 1; ++count; }
        }
        return count;
    }
    std::vector<float> p, bv, sv;
    std::vector<int> bc;
    auction(cost, R, C, row, col, nullptr, p, bv, sv, bc, R * 3);
    for (int r = 0; r < R; ++r) {                                      // :328-336
        const int c = row[r];
        if (c >= 0) {
            if (cost[(size_t)r * C + c] <= threshold) ++count;
            else { col[c] = -1; row[r] = -1; }
        }
    }
    return count;
}

// PreprocessorCUDA::preprocess + kernelPreprocess, preprocess.cu:19-153: letterbox resize (bilinear),
// BGR -> RGB, /255, HWC u8 -> CHW fp32, gray 114/255 padding.  xform4 = {scale_x, scale_y, pad_x, pad_y}
// as returned upstream (scale_x = scale_y = 1/scale; the pads are integers) — the transform
// scaleTrackOutputs (main.cpp:48-68) later undoes.
void orc_letterbox(const unsigned char* bgr, int w, int h, int tw, int th, float* out, float* xform4) {
    const float scale = std::min(static_cast<float>(tw) / w, static_cast<float>(th) / h);   // :109-112
    const int new_w = static_cast<int>(w * scale), new_h = static_cast<int>(h * scale);      // :114-115
    const int pad_x = (tw - new_w) / 2, pad_y = (th - new_h) / 2;                            // :117-118
    if (xform4) { xform4[0] = 1.0f / scale; xform4[1] = 1.0f / scale; xform4[2] = (float)pad_x; xform4[3] = (float)pad_y; }
    const size_t plane = (size_t)tw * th;
    for (int ty = 0; ty < th; ++ty)
        for (int tx = 0; tx < tw; ++tx) {
            float* o = out + (size_t)ty * tw + tx;
            if (tx < pad_x || tx >= pad_x + new_w || ty < pad_y || ty >= pad_y + new_h) {    // :39-47
                const float gray = 114.0f / 255.0f;
                o[0] = gray; o[plane] = gray; o[2 * plane] = gray;
                continue;
            }
            float sx = (tx - pad_x) / scale, sy = (ty - pad_y) / scale;                     // :50-51
            sx = std::fmin(std::fmax(sx, 0.0f), w - 1.001f);                                 // :54-55
            sy = std::fmin(std::fmax(sy, 0.0f), h - 1.001f);
            const int x0 = (int)sx, y0 = (int)sy;
            const int x1 = std::min(x0 + 1, w - 1), y1 = std::min(y0 + 1, h - 1);
            const float wx = sx - x0, wy = sy - y0;
            for (int c = 0; c < 3; ++c) {                                                    // :66-80
                const float v00 = bgr[((size_t)y0 * w + x0) * 3 + c], v01 = bgr[((size_t)y0 * w + x1) * 3 + c];
                const float v10 = bgr[((size_t)y1 * w + x0) * 3 + c], v11 = bgr[((size_t)y1 * w + x1) * 3 + c];
                const float v = (1 - wx) * (1 - wy) * v00 + wx * (1 - wy) * v01 + (1 - wx) * wy * v10 + wx * wy * v11;
                const int oc = (c == 0) ? 2 : (c == 2) ? 0 : c;
                o[oc * plane] = v / 255.0f;
            }
        }
}

void* orc_tracker_create(const orc_tracker_config* cfg) { return new Tracker(*cfg); }
void orc_tracker_destroy(void* t) { delete static_cast<Tracker*>(t); }
int orc_tracker_update(void* t, const float* dp, const float* ds, int n, int frame) {
    return static_cast<Tracker*>(t)->update(dp, ds, n, frame);
}
void orc_tracker_force_new(void* tp, const int* slots, const int* ids, int n) {
    Tracker* t = static_cast<Tracker*>(tp);
    t->forced_slot.assign(slots, slots + n);
    t->forced_id.assign(ids, ids + n);
}
int orc_tracker_forced_errors(void* tp) { return static_cast<Tracker*>(tp)->forced_errors; }
int orc_tracker_get_tracks(void* t, void* out, int cap) {
    return static_cast<Tracker*>(t)->get_tracks(static_cast<TrackOutput*>(out), cap);
}

#define ORC_COPY(dst, vec) do { if (dst) std::memcpy(dst, (vec).data(), (vec).size() * sizeof((vec)[0])); } while (0)
void orc_tracker_get_state(void* tp, float* poses, float* vel, float* scores, int* states, int* ids,
                           int* hits, int* ages, int* last_frame, int* active, int* row_assign,
                           int* col_assign, float* cost, float* predicted, float* centers,
                           int* scalars) {
    Tracker* t = static_cast<Tracker*>(tp);
    ORC_COPY(poses, t->poses); ORC_COPY(vel, t->vel); ORC_COPY(scores, t->scores);
    ORC_COPY(states, t->states); ORC_COPY(ids, t->ids); ORC_COPY(hits, t->hits);
    ORC_COPY(ages, t->ages); ORC_COPY(last_frame, t->last_frame); ORC_COPY(active, t->active);
    ORC_COPY(row_assign, t->row); ORC_COPY(col_assign, t->col); ORC_COPY(cost, t->cost);
    ORC_COPY(predicted, t->predicted); ORC_COPY(centers, t->tcent);
    if (scalars) { scalars[0] = t->next_id; scalars[1] = t->hint; scalars[2] = t->D; scalars[3] = t->num_active; }
}

void* orc_kf3_create(int T) { return new KF3(T); }
void orc_kf3_destroy(void* k) { delete static_cast<KF3*>(k); }
void orc_kf3_initiate(void* k, const float* d, const int* s, int n) { static_cast<KF3*>(k)->initiate(d, s, n); }
void orc_kf3_predict(void* k, int n, float am, float jm) { static_cast<KF3*>(k)->predict(n, am, jm); }
void orc_kf3_update(void* k, const float* d, const int* m, int n) { static_cast<KF3*>(k)->update(d, m, n); }
void orc_kf3_extract(void* k, float* o, const int* s, int n) { static_cast<KF3*>(k)->extract(o, s, n); }
void orc_kf3_get_state(void* kp, int track, float* mean, float* cov) {
    KF3* k = static_cast<KF3*>(kp);
    std::memcpy(mean, &k->mean[(size_t)track * 136], 136 * sizeof(float));
    if (cov) {
        std::memset(cov, 0, 136 * 136 * sizeof(float));
        for (int i = 0; i < 136; ++i) cov[i * 136 + i] = k->diag[(size_t)track * 136 + i];
    }
}
void orc_kf3_get_diag(void* kp, float* means, float* diag) {
    KF3* k = static_cast<KF3*>(kp);
    ORC_COPY(means, k->mean); ORC_COPY(diag, k->diag);
}

// Position-weighted 64-bit checksum of 32-bit words (order sensitive, numpy-reproducible):
//   h += (word + 0x9E37) * ((2*i + 1) * 0x9E3779B97F4A7C15)   (mod 2^64), i = running index.
static inline void mix_words(unsigned long long& h, unsigned long long& pos, const void* p, size_t nwords) {
    const uint32_t* w = static_cast<const uint32_t*>(p);
    for (size_t i = 0; i < nwords; ++i, ++pos)
        h += ((unsigned long long)w[i] + 0x9E37ULL) * ((2ULL * pos + 1ULL) * 0x9E3779B97F4A7C15ULL);
}

double orc_run_streams(const float* heads, int B, int F, int N, int frame_major, float conf_thr,
                       float nms_thr, int max_cand, int max_keep, const orc_tracker_config* cfg,
                       int n_threads, unsigned long long* out_hash, long long* out_tracks_total,
                       double* stage_seconds) {
    if (n_threads < 1) n_threads = 1;
    std::vector<double> st((size_t)n_threads * 3, 0.0);
    std::vector<long long> tot(n_threads, 0);
    const size_t slab = (size_t)56 * N;
    auto work = [&](int tid) {
        std::vector<float> cp((size_t)max_cand * 51), cb((size_t)max_cand * 4), cs(max_cand);
        std::vector<int> ca(max_cand), ks(max_cand), ka(max_cand);
        std::vector<TrackOutput> outs(cfg->max_detections);
        for (int b = tid; b < B; b += n_threads) {
            Tracker trk(*cfg);
            unsigned long long h = 0, pos = 0;
            for (int f = 0; f < F; ++f) {
                const float* raw = heads + (frame_major ? ((size_t)f * B + b) : ((size_t)b * F + f)) * slab;
                double t0 = now_s();
                int C = decode(raw, N, conf_thr, max_cand, cp.data(), cb.data(), cs.data(), ca.data());
                double t1 = now_s();
                int K = nms_native(cp.data(), cb.data(), cs.data(), C, nms_thr, max_keep, ks.data());
                double t2 = now_s();
                trk.update(cp.data(), cs.data(), K, f);
                int n = trk.get_tracks(outs.data(), (int)outs.size());
                double t3 = now_s();
                st[tid * 3 + 0] += t1 - t0; st[tid * 3 + 1] += t2 - t1; st[tid * 3 + 2] += t3 - t2;
                for (int k = 0; k < K; ++k) ka[k] = ca[ks[k]];
                uint32_t hdr[2] = {(uint32_t)K, (uint32_t)n};
                mix_words(h, pos, hdr, 2);
                mix_words(h, pos, ka.data(), K);
                mix_words(h, pos, outs.data(), (size_t)n * sizeof(TrackOutput) / 4);
                tot[tid] += n;
            }
            if (out_hash) out_hash[b] = h;
        }
    };
    double t0 = now_s();
    if (n_threads == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < n_threads; ++i) th.emplace_back(work, i);
        for (auto& t : th) t.join();
    }
    double wall = now_s() - t0;
    if (stage_seconds) {
        stage_seconds[0] = stage_seconds[1] = stage_seconds[2] = 0;
        for (int i = 0; i < n_threads; ++i)
            for (int s = 0; s < 3; ++s) stage_seconds[s] += st[i * 3 + s];
    }
    if (out_tracks_total) { *out_tracks_total = 0; for (auto v : tot) *out_tracks_total += v; }
    return wall;
}

}  // extern "C"
