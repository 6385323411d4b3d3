// aux_kernels.cu — stage-level entry points that are part of the reference's API surface
// but outside GPUTracker::update: the 3rd-order per-keypoint Kalman filter
// (src/cuda/kalman_filter.cu), the stand-alone auction solve (src/cuda/hungarian.cu), the
// NMS entry point declared in include/cuda/nms.h:48-60 and the host-legacy NMS rule set
// (src/cuda/nms.cu:142-306) on the device.
#include "pb_common.cuh"
#include "auction.cuh"

namespace pb {

// =======================================================================================
// KF3.  State per track: mean[136] + covariance diagonal[136] (the reference keeps a
// 136x136 matrix per track, 74 KB, of which only the diagonal is ever non-zero:
// kalman_filter.cu:59-81, 138-167, 211-236).  One thread per (track, keypoint) owns the
// keypoint's 8 state entries: two 128-bit loads + stores each for mean and diagonal.
// =======================================================================================
__global__ void kf3_initiate_kernel(float* means, float* diag, const float* dets, const int* slots, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;           // kalman_filter.cu:24-82
    if (i >= n * KP) return;
    const int e = i / KP, k = i - e * KP;
    const int t = slots[e];
    const float x = dets[e * POSE_F + k * 3], y = dets[e * POSE_F + k * 3 + 1], cf = dets[e * POSE_F + k * 3 + 2];
    float4* m = reinterpret_cast<float4*>(means + (size_t)t * 136 + k * 8);
    float4* p = reinterpret_cast<float4*>(diag + (size_t)t * 136 + k * 8);
    m[0] = make_float4(x, y, 0.f, 0.f);
    m[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float pv = (cf > 0.0f) ? 10.0f : 1000.0f;
    p[0] = make_float4(pv, pv, 100.0f, 100.0f);
    p[1] = make_float4(100.0f, 100.0f, 100.0f, 100.0f);
}

__global__ void kf3_predict_kernel(float* means, float* diag, int n, float am, float jm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;           // :86-167 (slots [0,n))
    if (i >= n * KP) return;
    float4* m = reinterpret_cast<float4*>(means + (size_t)i * 8);
    float4* p = reinterpret_cast<float4*>(diag + (size_t)i * 8);
    const float4 a = m[0], c = m[1];                               // px py vx vy | ax ay jx jy
    float4 na, nc;
    na.x = a.x + a.z + 0.5f * c.x + (1.0f / 6.0f) * c.z;
    na.y = a.y + a.w + 0.5f * c.y + (1.0f / 6.0f) * c.w;
    na.z = a.z + c.x + 0.5f * c.z;
    na.w = a.w + c.y + 0.5f * c.w;
    nc.x = c.x * am; nc.y = c.y * am; nc.z = c.z * jm; nc.w = c.w * jm;
    m[0] = na; m[1] = nc;
    float4 d0 = p[0], d1 = p[1];
    d0.x += 1.0f * 1.0f; d0.y += 1.0f * 1.0f; d0.z += 0.5f * 0.5f; d0.w += 0.5f * 0.5f;
    d1.x += 0.1f * 0.1f; d1.y += 0.1f * 0.1f; d1.z += 0.05f * 0.05f; d1.w += 0.05f * 0.05f;
    p[0] = d0; p[1] = d1;
}

__global__ void kf3_update_kernel(float* means, float* diag, const float* dets, const int* matches, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;           // :171-237
    if (i >= n * KP) return;
    const int e = i / KP, k = i - e * KP;
    const int t = matches[e * 2], d = matches[e * 2 + 1];
    const float zx = dets[d * POSE_F + k * 3], zy = dets[d * POSE_F + k * 3 + 1], cf = dets[d * POSE_F + k * 3 + 2];
    if (cf < 0.1f) return;
    float4* m = reinterpret_cast<float4*>(means + (size_t)t * 136 + k * 8);
    float2* p = reinterpret_cast<float2*>(diag + (size_t)t * 136 + k * 8);
    float4 a = m[0];
    const float2 P = p[0];
    const float yx = zx - a.x, yy = zy - a.y;
    const float R = 5.0f / (cf + 0.1f);
    const float Sxx = P.x + R, Syy = P.y + R;
    const float Kx = P.x / Sxx, Ky = P.y / Syy;
    a.x += Kx * yx;
    a.y += Ky * yy;
    const float Kv = 0.5f * Kx;
    a.z += Kv * yx;
    a.w += Kv * yy;
    m[0] = a;
    p[0] = make_float2((1.0f - Kx) * P.x, (1.0f - Ky) * P.y);
}

__global__ void kf3_extract_kernel(const float* means, float* out, const int* slots, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;           // :241-264
    if (i >= n * KP) return;
    const int e = i / KP, k = i - e * KP;
    const int t = slots[e];
    out[e * POSE_F + k * 3] = means[(size_t)t * 136 + k * 8];
    out[e * POSE_F + k * 3 + 1] = means[(size_t)t * 136 + k * 8 + 1];
    out[e * POSE_F + k * 3 + 2] = 1.0f;
}

__global__ void kf3_cov_kernel(const float* diag, int track, float* cov) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 136 * 136) return;
    const int r = i / 136, c = i - r * 136;
    cov[i] = (r == c) ? diag[(size_t)track * 136 + r] : 0.0f;
}

// =======================================================================================
// Stand-alone auction: one CTA per problem, cost rows streamed from global/L2.
// =======================================================================================
__global__ void __launch_bounds__(256)
auction_batch_kernel(const float* cost, int R, int C, int* row_out, int* col_out, const int* active, int stage_cost) {
    extern __shared__ __align__(16) unsigned char sm[];
    unsigned long long* colbid = reinterpret_cast<unsigned long long*>(sm);
    float* price = reinterpret_cast<float*>(colbid + C);
    int* col = reinterpret_cast<int*>(price + C);
    int* row = col + C;
    int* act = row + R;
    int* flags = act + R;          // [4]: 0-1 iteration flags, 2 active count
    int* owner = flags + 4;        // [C]   (stage_cost only)
    int* act_list = owner + C;     // [R]
    float* cost_s = reinterpret_cast<float*>(act_list + R);   // [R*C]
    const int b = blockIdx.x;
    const float* cb = cost + (size_t)b * R * C;
    for (int t = threadIdx.x; t < R; t += blockDim.x) act[t] = active ? active[(size_t)b * R + t] : 1;
    if (stage_cost) for (int i = threadIdx.x; i < R * C; i += blockDim.x) cost_s[i] = cb[i];
    __syncthreads();
    int na = 33;
    if (stage_cost) {
        // ordered list of the active rows; the single-warp solve takes up to 32 of them
        if (threadIdx.x == 0) {
            int n = 0;
            for (int t = 0; t < R; ++t) if (act[t] != 0) act_list[n++] = t;
            flags[2] = n;
        }
        __syncthreads();
        na = flags[2];
    }
    if (na <= 32 && C <= 64) {
        float* cc = cost_s + (size_t)R * C;      // [32*C] compacted active rows
        for (int i = threadIdx.x; i < na * C; i += blockDim.x) { const int ai = i / C; cc[i] = cost_s[act_list[ai] * C + (i - ai * C)]; }
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned* cb32 = reinterpret_cast<unsigned*>(colbid);
            int* cr32 = reinterpret_cast<int*>(colbid) + C;
            if (C <= 32) auction_solve_lean32<1>(cc, R, C, act_list, na, row, col, price, owner, cb32, cr32);
            else auction_solve_lean32<2>(cc, R, C, act_list, na, row, col, price, owner, cb32, cr32);
        }
        __syncthreads();
    } else if (na <= 32) {
        if (threadIdx.x < 32)
            auction_solve_hybrid32(cost_s, R, C, act_list, na, row, col, price, owner,
                                   reinterpret_cast<unsigned*>(colbid), reinterpret_cast<int*>(colbid) + C);
        __syncthreads();
    } else {
        auction_solve_cta(cb, R, C, act, row, col, price, colbid, flags, threadIdx.x, blockDim.x);
    }
    for (int t = threadIdx.x; t < R; t += blockDim.x) row_out[(size_t)b * R + t] = row[t];
    for (int d = threadIdx.x; d < C; d += blockDim.x) col_out[(size_t)b * C + d] = col[d];
}

// =======================================================================================
// launchPoseNMS (nms.h:48-60; no definition upstream).  poses [n, nk*3], keep[i] in {0,1}.
// Pair rule: OKS of kernelComputeOKSMatrix (nms.cu:25-117): scale = larger keypoint-box
// area (conf > 0.2), 0 if scale < 32^2 or fewer than 3 valid keypoints on either side.
// One CTA: rank by (score desc, index asc), then greedy in rank order; each survivor
// strikes the lower ranks in parallel.
// =======================================================================================
__device__ float pose_oks_pair(const float* poses, const float* sig, int i, int j, int nk) {
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
    const float ai = (hxi - lxi) * (hyi - lyi), aj = (hxj - lxj) * (hyj - lyj);
    const float scale_sq = pb_max(ai, aj);
    if (scale_sq < 32.0f * 32.0f || vi < 3 || vj < 3) return 0.0f;
    float sum = 0.0f;
    int cnt = 0;
    for (int k = 0; k < nk; ++k) {
        if (a[k * 3 + 2] > 0.2f && b[k * 3 + 2] > 0.2f) {
            const float dx = a[k * 3] - b[k * 3], dy = a[k * 3 + 1] - b[k * 3 + 1];
            const float d2 = dx * dx + dy * dy;
            const float s = sig[k];
            sum += pb_expf(-d2 / (2.0f * scale_sq * 4.0f * s * s));
            ++cnt;
        }
    }
    return (cnt >= 3) ? (sum / (float)cnt) : 0.0f;
}

__global__ void __launch_bounds__(512)
pose_nms_kernel(const float* poses, const float* scores, const float* sigmas, int* keep, int n, int nk,
                float oks_thr, float score_thr) {
    extern __shared__ __align__(16) unsigned char sm[];
    int* order = reinterpret_cast<int*>(sm);            // [n] rank -> index (valid ranks first)
    unsigned* sup = reinterpret_cast<unsigned*>(order + n);
    __shared__ int s_m;
    const int tid = threadIdx.x, NT = blockDim.x;
    if (tid == 0) s_m = 0;
    for (int i = tid; i < (n + 31) / 32; i += NT) sup[i] = 0u;
    for (int i = tid; i < n; i += NT) keep[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
        const float si = scores[i];
        if (!(si >= score_thr)) continue;
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const float sj = scores[j];
            if (sj >= score_thr && (sj > si || (sj == si && j < i))) ++rank;
        }
        order[rank] = i;
        atomicAdd(&s_m, 1);
    }
    __syncthreads();
    const int m = s_m;
    for (int r = 0; r < m; ++r) {
        if ((sup[r >> 5] >> (r & 31)) & 1u) continue;   // uniform: read after barrier below
        const int i = order[r];
        if (tid == 0) keep[i] = 1;
        for (int q = r + 1 + tid; q < m; q += NT) {
            if ((sup[q >> 5] >> (q & 31)) & 1u) continue;
            const int j = order[q];
            const int lo = i < j ? i : j, hi = i < j ? j : i;
            if (pose_oks_pair(poses, sigmas, lo, hi, nk) > oks_thr) atomicOr(&sup[q >> 5], 1u << (q & 31));
        }
        __syncthreads();
    }
}

// =======================================================================================
// Host-legacy NMS rule set (NMSCuda::apply, nms.cu:142-306) on the device, one CTA per
// image.  dets = PoseDetection (56 floats: bbox[4], score, 17 x (x,y,conf)).
// =======================================================================================
__device__ __forceinline__ float legacy_iou(const float* a, const float* b) {     // nms.cu:166-181
    const float x1 = pb_max(a[0], b[0]), y1 = pb_max(a[1], b[1]);
    const float x2 = pb_min(a[2], b[2]), y2 = pb_min(a[3], b[3]);
    const float iw = pb_max(0.0f, x2 - x1), ih = pb_max(0.0f, y2 - y1);
    const float inter = iw * ih;
    const float a1 = (a[2] - a[0]) * (a[3] - a[1]);
    const float a2 = (b[2] - b[0]) * (b[3] - b[1]);
    const float uni = a1 + a2 - inter;
    return (uni > 0) ? (inter / uni) : 0.0f;
}

__device__ float legacy_oks(const float* p, const float* q) {                      // nms.cu:184-234
    const float* kp = p + 5;
    const float* kq = q + 5;
    float lx1 = 1e9f, ly1 = 1e9f, hx1 = -1e9f, hy1 = -1e9f;
    float lx2 = 1e9f, ly2 = 1e9f, hx2 = -1e9f, hy2 = -1e9f;
    int v1 = 0, v2 = 0;
    for (int k = 0; k < KP; ++k) {
        if (kp[k * 3 + 2] > 0.2f) {
            lx1 = pb_min(lx1, kp[k * 3]); ly1 = pb_min(ly1, kp[k * 3 + 1]);
            hx1 = pb_max(hx1, kp[k * 3]); hy1 = pb_max(hy1, kp[k * 3 + 1]); ++v1;
        }
        if (kq[k * 3 + 2] > 0.2f) {
            lx2 = pb_min(lx2, kq[k * 3]); ly2 = pb_min(ly2, kq[k * 3 + 1]);
            hx2 = pb_max(hx2, kq[k * 3]); hy2 = pb_max(hy2, kq[k * 3 + 1]); ++v2;
        }
    }
    if (v1 < 3 || v2 < 3) return 0.0f;
    const float a1 = (hx1 - lx1) * (hy1 - ly1), a2 = (hx2 - lx2) * (hy2 - ly2);
    float scale_sq = pb_max(a1, a2);
    if (scale_sq < 32.0f * 32.0f) scale_sq = 32.0f * 32.0f;
    float sum = 0.0f;
    int cnt = 0;
    for (int k = 0; k < KP; ++k) {
        if (kp[k * 3 + 2] > 0.2f && kq[k * 3 + 2] > 0.2f) {
            const float dx = kp[k * 3] - kq[k * 3], dy = kp[k * 3 + 1] - kq[k * 3 + 1];
            const float d2 = dx * dx + dy * dy;
            const float s = kSigmas[k];
            sum += pb_expf(-d2 / (2.0f * scale_sq * 4.0f * s * s));
            ++cnt;
        }
    }
    return (cnt >= 3) ? (sum / (float)cnt) : 0.0f;
}

__device__ bool legacy_suppresses(const float* a, const float* b) {                // nms.cu:259-301
    const float iou = legacy_iou(a, b);
    if (iou > 0.55f) return true;
    const float oks = legacy_oks(a, b);
    if (oks > 0.5f) return true;
    if (iou > 0.2f && oks > 0.4f) return true;
    const float cx1 = (a[0] + a[2]) / 2.0f, cy1 = (a[1] + a[3]) / 2.0f;
    const float cx2 = (b[0] + b[2]) / 2.0f, cy2 = (b[1] + b[3]) / 2.0f;
    const float w1 = a[2] - a[0], h1 = a[3] - a[1];
    float scale = pb_max(w1, h1);
    if (scale < 32.0f) scale = 32.0f;
    const float ddx = cx1 - cx2, ddy = cy1 - cy2;
    const float dist = sqrtf(ddx * ddx + ddy * ddy);
    const float nd = dist / scale;
    return nd < 0.3f && oks > 0.15f;
}

__global__ void __launch_bounds__(512)
nms_legacy_kernel(const float* dets, const int* offsets, int cap, float score_thr, int* keep_out, int* num_keep) {
    extern __shared__ __align__(16) unsigned char sm[];
    int* order = reinterpret_cast<int*>(sm);
    unsigned* sup = reinterpret_cast<unsigned*>(order + cap);
    __shared__ int s_m, s_nk;
    const int img = blockIdx.x, tid = threadIdx.x, NT = blockDim.x;
    const int off = offsets[img];
    int n = offsets[img + 1] - off;
    if (n > cap) n = cap;
    const float* base = dets + (size_t)off * 56;
    if (tid == 0) { s_m = 0; s_nk = 0; }
    for (int i = tid; i < (cap + 31) / 32; i += NT) sup[i] = 0u;
    __syncthreads();
    for (int i = tid; i < n; i += NT) {
        const float si = base[(size_t)i * 56 + 4];
        if (!(si >= score_thr)) continue;                          // nms.cu:154
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const float sj = base[(size_t)j * 56 + 4];
            if (sj >= score_thr && (sj > si || (sj == si && j < i))) ++rank;
        }
        order[rank] = i;
        atomicAdd(&s_m, 1);
    }
    __syncthreads();
    const int m = s_m;
    int* ko = keep_out + off;
    for (int r = 0; r < m; ++r) {
        if ((sup[r >> 5] >> (r & 31)) & 1u) continue;
        const int i = order[r];
        if (tid == 0) { ko[s_nk] = i; s_nk = s_nk + 1; }
        for (int q = r + 1 + tid; q < m; q += NT) {
            if ((sup[q >> 5] >> (q & 31)) & 1u) continue;
            if (legacy_suppresses(base + (size_t)i * 56, base + (size_t)order[q] * 56))
                atomicOr(&sup[q >> 5], 1u << (q & 31));
        }
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) num_keep[img] = s_nk;
}

}  // namespace pb

// =======================================================================================
// C ABI
// =======================================================================================
using namespace pb;

static inline unsigned nblk(int n, int t) { return (unsigned)((n + t - 1) / t); }
void pb_set_error(const char* fmt, ...);

extern "C" {

void launchPoseNMS(const float* poses, const float* scores, const float* sigmas, int* keep,
                   int num_detections, int num_keypoints, float oks_threshold,
                   float score_threshold, pb_stream_t stream) {
    if (num_detections <= 0) return;
    const size_t smem = (size_t)num_detections * 4 + (size_t)((num_detections + 31) / 32) * 4 + 16;
    ensure_dyn_smem((const void*)pose_nms_kernel, smem);
    pose_nms_kernel<<<1, 512, smem, (cudaStream_t)stream>>>(poses, scores, sigmas, keep, num_detections,
                                                           num_keypoints, oks_threshold, score_threshold);
    count_launch();
}

int pb_nms_legacy(const void* d_dets, const int* d_offsets, int num_images, int max_per_image,
                  float /*oks_threshold: ignored upstream, nms.cu:142*/, float score_threshold,
                  int* d_keep, int* d_num_keep, pb_stream_t stream) {
    if (num_images <= 0) return PB_OK;
    if (max_per_image <= 0 || max_per_image > 16384) { pb_set_error("pb_nms_legacy: max_per_image out of range"); return PB_ERR_INVALID; }
    const size_t smem = (size_t)max_per_image * 4 + (size_t)((max_per_image + 31) / 32) * 4 + 16;
    {
        cudaError_t e = ensure_dyn_smem((const void*)nms_legacy_kernel, smem);
        if (e != cudaSuccess) { pb_set_error("pb_nms_legacy: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    }
    nms_legacy_kernel<<<num_images, 512, smem, (cudaStream_t)stream>>>(
        static_cast<const float*>(d_dets), d_offsets, max_per_image, score_threshold, d_keep, d_num_keep);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { pb_set_error("pb_nms_legacy: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    return PB_OK;
}

int pb_auction_solve(const float* d_cost, int batch, int num_rows, int num_cols, int* d_row_assign,
                     int* d_col_assign, const int* d_row_active, pb_stream_t stream) {
    if (batch <= 0 || num_rows <= 0 || num_cols <= 0) return PB_OK;          // hungarian.cu:368
    size_t smem = (size_t)num_cols * 16 + (size_t)num_rows * 8 + 16;
    // small tables are staged in shared memory; problems with at most 32 active rows then take the
    // single-warp hybrid solve the tracker uses (auction.cuh), the others the CTA-wide one
    const size_t staged = smem + (size_t)num_cols * 4 + (size_t)num_rows * 4 + (size_t)num_rows * num_cols * 4 + (num_cols <= 64 ? (size_t)32 * num_cols * 4 : 0);
    const int stage_cost = (staged <= 96 * 1024 && num_cols <= 65535) ? 1 : 0;
    if (stage_cost) smem = staged;
    if (smem > 200 * 1024) { pb_set_error("pb_auction_solve: problem too large"); return PB_ERR_UNSUPPORTED; }
    {
        cudaError_t e = ensure_dyn_smem((const void*)auction_batch_kernel, smem);
        if (e != cudaSuccess) { pb_set_error("pb_auction_solve: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    }
    auction_batch_kernel<<<batch, 256, smem, (cudaStream_t)stream>>>(d_cost, num_rows, num_cols, d_row_assign,
                                                                     d_col_assign, d_row_active, stage_cost);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { pb_set_error("pb_auction_solve: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    return PB_OK;
}

#define PB_KF3_CHECK(name)                                                              \
    do { cudaError_t e_ = cudaGetLastError();                                           \
         if (e_ != cudaSuccess) { pb_set_error(name ": %s", cudaGetErrorString(e_)); return PB_ERR_CUDA; } \
         return PB_OK; } while (0)

int pb_kf3_initiate(float* d_means, float* d_diag, const float* d_dets, const int* d_slots, int n, pb_stream_t st) {
    if (n <= 0) return PB_OK;
    kf3_initiate_kernel<<<nblk(n * KP, 256), 256, 0, (cudaStream_t)st>>>(d_means, d_diag, d_dets, d_slots, n);
    count_launch();
    PB_KF3_CHECK("pb_kf3_initiate");
}
int pb_kf3_predict(float* d_means, float* d_diag, int n, float am, float jm, pb_stream_t st) {
    if (n <= 0) return PB_OK;
    kf3_predict_kernel<<<nblk(n * KP, 256), 256, 0, (cudaStream_t)st>>>(d_means, d_diag, n, am, jm);
    count_launch();
    PB_KF3_CHECK("pb_kf3_predict");
}
int pb_kf3_update(float* d_means, float* d_diag, const float* d_dets, const int* d_matches, int n, pb_stream_t st) {
    if (n <= 0) return PB_OK;
    kf3_update_kernel<<<nblk(n * KP, 256), 256, 0, (cudaStream_t)st>>>(d_means, d_diag, d_dets, d_matches, n);
    count_launch();
    PB_KF3_CHECK("pb_kf3_update");
}
int pb_kf3_extract(const float* d_means, float* d_out, const int* d_slots, int n, pb_stream_t st) {
    if (n <= 0) return PB_OK;
    kf3_extract_kernel<<<nblk(n * KP, 256), 256, 0, (cudaStream_t)st>>>(d_means, d_out, d_slots, n);
    count_launch();
    PB_KF3_CHECK("pb_kf3_extract");
}
int pb_kf3_materialize_cov(const float* d_diag, int track, float* d_cov, pb_stream_t st) {
    kf3_cov_kernel<<<nblk(136 * 136, 256), 256, 0, (cudaStream_t)st>>>(d_diag, track, d_cov);
    count_launch();
    PB_KF3_CHECK("pb_kf3_materialize_cov");
}

}  // extern "C"
