// next_rows.cu — the rows next to the hot path: OKSDistanceCUDA's device entry
// points and a deterministic GreedyMatcherCUDA, batched over independent problems.
//
//   pb_pose_distance   reference src/cuda/oks_distance.cu: kernelOKSDistance (:26-164, ungated OKS cost
//                      with the 0.05-confidence fallback), kernelExtractBboxes (:213-245, keypoint box of
//                      confidence > 0 with a 10 px margin) + kernelIoUDistance (:167-210), kernelCombineCosts
//                      (:248-261).  Upstream: up to five launches and two temporaries per call; here one
//                      launch, per-pose statistics computed once per CTA tile in shared memory.
//   pb_greedy_match    reference src/cuda/hungarian.cu: GreedyMatcherCUDA.  The device kernel upstream
//                      (kernelGreedyMatch :126-157) races on the columns, its outcome depends on thread
//                      timing; the class's own host path (:441-467) is deterministic — all cells below the
//                      threshold sorted by (cost, row, col), taken greedily — and is the rule used here.
//   pb_assign_legacy   reference src/cuda/hungarian.cu:235-339, LinearAssignmentCUDA::solve: greedy rows below
//                      100 cells (:198-233), else the auction over all rows with 3*rows iterations (:283) and
//                      the threshold filter of :328-336 — upstream one host round trip per iteration.
//   pb_letterbox_batch reference src/cuda/preprocess.cu:19-153: letterbox + BGR->RGB +
//                      /255 + HWC->CHW for a batch of frames of individual sizes, one launch; the letterbox
//                      parameters come out in the layout pb_set_output_transform takes.
#include "pb_common.cuh"
#include "auction.cuh"

namespace pb {

constexpr int PD_TILE = 16;

struct PoseStat { float area; float x1, y1, x2, y2; int valid01; };

// area of the keypoint box with confidence > 0.1 (oks_distance.cu:43-80), number of such keypoints,
// keypoint box with confidence > 0 and a 10 px margin (:213-245)
__device__ __forceinline__ PoseStat pose_stat(const float* p) {
    float lx = 1e9f, ly = 1e9f, hx = -1e9f, hy = -1e9f;
    float bx1 = 1e9f, by1 = 1e9f, bx2 = -1e9f, by2 = -1e9f;
    int n = 0;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        const float x = p[k * 3], y = p[k * 3 + 1], c = p[k * 3 + 2];
        if (c > 0.1f) { lx = pb_min(lx, x); ly = pb_min(ly, y); hx = pb_max(hx, x); hy = pb_max(hy, y); ++n; }
        if (c > 0.0f) { bx1 = pb_min(bx1, x); by1 = pb_min(by1, y); bx2 = pb_max(bx2, x); by2 = pb_max(by2, y); }
    }
    PoseStat s;
    s.area = (hx - lx) * (hy - ly);
    s.valid01 = n;
    s.x1 = bx1 - 10.0f; s.y1 = by1 - 10.0f; s.x2 = bx2 + 10.0f; s.y2 = by2 + 10.0f;
    return s;
}

__device__ __forceinline__ float oks_cost_ungated(const float* tp, const float* dp, float t_area, float d_area, int d_valid) {
    float scale_sq = (d_area + t_area) * 0.5f;                         // :78-86
    if (scale_sq < 1000.0f) scale_sq = 1000.0f;
    if (d_valid < 2) return 1.0f;                                      // :89-92
    float sum = 0.0f;
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        if (dp[k * 3 + 2] > 0.2f && tp[k * 3 + 2] > 0.2f) {            // :101-127
            const float dx = dp[k * 3] - tp[k * 3], dy = dp[k * 3 + 1] - tp[k * 3 + 1];
            const float d2 = dx * dx + dy * dy;
            const float sg = kSigmas[k] * 2.0f;
            sum += pb_expf(-d2 / (2.0f * scale_sq * (sg * sg)));
            ++cnt;
        }
    }
    float oks;
    if (cnt >= 3) {
        oks = sum / (float)cnt;
    } else {                                                           // fallback :134-160
        sum = 0.0f; cnt = 0;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            if (dp[k * 3 + 2] > 0.05f && tp[k * 3 + 2] > 0.05f) {
                const float dx = dp[k * 3] - tp[k * 3], dy = dp[k * 3 + 1] - tp[k * 3 + 1];
                const float d2 = dx * dx + dy * dy;
                const float sg = kSigmas[k] * 2.0f;
                sum += pb_expf(-d2 / (2.0f * scale_sq * (sg * sg)));
                ++cnt;
            }
        }
        oks = cnt > 0 ? (sum / (float)cnt) : 0.0f;
    }
    return 1.0f - oks;
}

__device__ __forceinline__ float iou_cost(const PoseStat& t, const PoseStat& d) {   // :183-209
    const float ix1 = pb_max(t.x1, d.x1), iy1 = pb_max(t.y1, d.y1);
    const float ix2 = pb_min(t.x2, d.x2), iy2 = pb_min(t.y2, d.y2);
    const float iw = pb_max(0.0f, ix2 - ix1), ih = pb_max(0.0f, iy2 - iy1);
    const float inter = iw * ih;
    const float ta = (t.x2 - t.x1) * (t.y2 - t.y1), da = (d.x2 - d.x1) * (d.y2 - d.y1);
    const float uni = ta + da - inter;
    const float iou = (uni > 0.0f) ? (inter / uni) : 0.0f;
    return 1.0f - iou;
}

// grid (ceil(nd/16), ceil(nt/16), batch), block 16x16: a CTA owns a 16x16 tile of one problem's matrix;
// the 16 track poses and 16 detection poses of the tile are staged in shared memory with their statistics.
__global__ void __launch_bounds__(PD_TILE * PD_TILE)
pose_distance_kernel(const float* __restrict__ tracks, const float* __restrict__ dets, int nt, int nd, int mode, float alpha,
                     float* __restrict__ out) {
    __shared__ float s_t[PD_TILE][POSE_F], s_d[PD_TILE][POSE_F];
    __shared__ PoseStat st_t[PD_TILE], st_d[PD_TILE];
    const int b = blockIdx.z;
    const int t0 = blockIdx.y * PD_TILE, d0 = blockIdx.x * PD_TILE;
    const int tid = threadIdx.y * PD_TILE + threadIdx.x;
    const float* tr = tracks + (size_t)b * nt * POSE_F;
    const float* de = dets + (size_t)b * nd * POSE_F;
    for (int i = tid; i < PD_TILE * POSE_F; i += PD_TILE * PD_TILE) {
        const int r = i / POSE_F, e = i - r * POSE_F;
        s_t[r][e] = (t0 + r < nt) ? tr[(size_t)(t0 + r) * POSE_F + e] : 0.0f;
        s_d[r][e] = (d0 + r < nd) ? de[(size_t)(d0 + r) * POSE_F + e] : 0.0f;
    }
    __syncthreads();
    if (tid < PD_TILE) st_t[tid] = pose_stat(s_t[tid]);
    else if (tid < 2 * PD_TILE) st_d[tid - PD_TILE] = pose_stat(s_d[tid - PD_TILE]);
    __syncthreads();
    const int t = t0 + threadIdx.y, d = d0 + threadIdx.x;
    if (t >= nt || d >= nd) return;
    const PoseStat& a = st_t[threadIdx.y];
    const PoseStat& c = st_d[threadIdx.x];
    float cost;
    if (mode == 0) cost = oks_cost_ungated(s_t[threadIdx.y], s_d[threadIdx.x], a.area, c.area, c.valid01);
    else if (mode == 1) cost = iou_cost(a, c);
    else {
        const float o = oks_cost_ungated(s_t[threadIdx.y], s_d[threadIdx.x], a.area, c.area, c.valid01);
        const float u = iou_cost(a, c);
        cost = alpha * o + (1.0f - alpha) * u;                         // :260
    }
    out[((size_t)b * nt + t) * nd + d] = cost;
}

// One CTA per problem.  Repeats: block-wide minimum of (cost, row, col) over the cells below the
// threshold whose row and column are still free; takes it.  Equals the sorted sweep of hungarian.cu:441-467.
__global__ void __launch_bounds__(256)
greedy_match_kernel(const float* __restrict__ cost, int R, int C, float threshold, int* __restrict__ row_matched) {
    extern __shared__ unsigned char sm[];
    unsigned* row_used = reinterpret_cast<unsigned*>(sm);              // [R] 0/1
    unsigned* col_used = row_used + R;                                 // [C]
    unsigned long long* best = reinterpret_cast<unsigned long long*>(col_used + C + ((R + C) & 1));   // [8] per-warp minima
    __shared__ unsigned long long s_min;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* cm = cost + (size_t)b * R * C;
    int* out = row_matched + (size_t)b * R;
    for (int r = tid; r < R; r += blockDim.x) { row_used[r] = 0u; out[r] = -1; }
    for (int c = tid; c < C; c += blockDim.x) col_used[c] = 0u;
    __syncthreads();
    const int steps = R < C ? R : C;
    for (int s = 0; s < steps; ++s) {
        // key = order-preserving cost bits << 32 | cell index (row-major: row then column ascending)
        unsigned long long mine = ~0ull;
        for (int i = tid; i < R * C; i += blockDim.x) {
            const int r = i / C, c = i - r * C;
            const float v = cm[i];
            if (v < threshold && !row_used[r] && !col_used[c]) {
                unsigned u = __float_as_uint(v);
                if (u == 0x80000000u) u = 0u;
                u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
                const unsigned long long key = ((unsigned long long)u << 32) | (unsigned)i;
                mine = key < mine ? key : mine;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, mine, off);
            mine = o < mine ? o : mine;
        }
        if (lane == 0) best[warp] = mine;
        __syncthreads();
        if (tid == 0) {
            unsigned long long m = ~0ull;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = best[w] < m ? best[w] : m;
            s_min = m;
            if (m != ~0ull) {
                const int i = (int)(m & 0xffffffffull);
                const int r = i / C, c = i - r * C;
                row_used[r] = 1u; col_used[c] = 1u; out[r] = c;
            }
        }
        __syncthreads();
        if (s_min == ~0ull) break;
    }
}

// ---------------------------------------------------------------------------------------
// LinearAssignmentCUDA::solve.  One CTA per problem.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
assign_legacy_kernel(const float* cost, int R, int C, float threshold, int* row_out, int* col_out, int* count_out) {
    extern __shared__ __align__(16) unsigned char sm[];
    unsigned long long* colbid = reinterpret_cast<unsigned long long*>(sm);
    float* price = reinterpret_cast<float*>(colbid + C);
    int* col = reinterpret_cast<int*>(price + C);
    int* row = col + C;
    int* flags = row + R;          // [4]
    const int b = blockIdx.x;
    const float* cb = cost + (size_t)b * R * C;
    if (R * C < 100) {                                                 // greedyAssign, :198-233 (sequential by definition)
        if (threadIdx.x == 0) {
            unsigned long long used_lo = 0ull, used_hi = 0ull;         // C < 100
            int n = 0;
            for (int c = 0; c < C; ++c) col[c] = -1;
            for (int r = 0; r < R; ++r) {
                float best = threshold;
                int bc = -1;
                for (int c = 0; c < C; ++c) {
                    const bool used = c < 64 ? ((used_lo >> c) & 1ull) : ((used_hi >> (c - 64)) & 1ull);
                    if (!used) { const float v = cb[r * C + c]; if (v < best) { best = v; bc = c; } }
                }
                row[r] = bc;
                if (bc >= 0) { col[bc] = r; if (bc < 64) used_lo |= 1ull << bc; else used_hi |= 1ull << (bc - 64); ++n; }
            }
            flags[2] = n;
        }
        __syncthreads();
    } else {
        auction_solve_cta(cb, R, C, nullptr, row, col, price, colbid, flags, threadIdx.x, blockDim.x, R * 3);
        if (threadIdx.x == 0) flags[2] = 0;
        __syncthreads();
        int n = 0;
        for (int r = threadIdx.x; r < R; r += blockDim.x) {            // :328-336
            const int c = row[r];
            if (c >= 0) {
                if (cb[(size_t)r * C + c] <= threshold) ++n;
                else { col[c] = -1; row[r] = -1; }
            }
        }
        if (n) atomicAdd(&flags[2], n);
        __syncthreads();
    }
    for (int r = threadIdx.x; r < R; r += blockDim.x) row_out[(size_t)b * R + r] = row[r];
    for (int c = threadIdx.x; c < C; c += blockDim.x) col_out[(size_t)b * C + c] = col[c];
    if (threadIdx.x == 0 && count_out) count_out[b] = flags[2];
}

// ---------------------------------------------------------------------------------------
// Batched letterbox pre-processing.  HBM-bound: 3 B read per source pixel touched, 12 B written per
// target pixel.  One thread produces four consecutive target pixels of one row (three 128-bit
// stores, one per colour plane); source taps go through the read-only path (neighbouring threads
// share their sectors).  Arithmetic as kernelPreprocess (:36-80) expression by expression.
// ---------------------------------------------------------------------------------------
struct LetterboxGeom { float scale; int new_w, new_h, pad_x, pad_y; };

__device__ __forceinline__ LetterboxGeom letterbox_geom(int w, int h, int tw, int th) {
    LetterboxGeom g;
    g.scale = pb_min(static_cast<float>(tw) / w, static_cast<float>(th) / h);              // :109-112
    g.new_w = static_cast<int>(w * g.scale);                                                // :114-115
    g.new_h = static_cast<int>(h * g.scale);
    g.pad_x = (tw - g.new_w) / 2;                                                           // :117-118
    g.pad_y = (th - g.new_h) / 2;
    return g;
}

__device__ __forceinline__ void letterbox_pixel(const unsigned char* __restrict__ img, int w, int h, const LetterboxGeom& g,
                                                int tx, int ty, float& r, float& gch, float& b) {
    if (tx < g.pad_x || tx >= g.pad_x + g.new_w || ty < g.pad_y || ty >= g.pad_y + g.new_h) {   // :39-47
        r = gch = b = 114.0f / 255.0f;
        return;
    }
    float sx = (tx - g.pad_x) / g.scale, sy = (ty - g.pad_y) / g.scale;                    // :50-51
    sx = fminf(fmaxf(sx, 0.0f), w - 1.001f);                                                // :54-55
    sy = fminf(fmaxf(sy, 0.0f), h - 1.001f);
    const int x0 = (int)sx, y0 = (int)sy;
    const int x1 = min(x0 + 1, w - 1), y1 = min(y0 + 1, h - 1);
    const float wx = sx - x0, wy = sy - y0;
    const unsigned char* p00 = img + ((size_t)y0 * w + x0) * 3;
    const unsigned char* p01 = img + ((size_t)y0 * w + x1) * 3;
    const unsigned char* p10 = img + ((size_t)y1 * w + x0) * 3;
    const unsigned char* p11 = img + ((size_t)y1 * w + x1) * 3;
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {                                                           // :66-80
        const float v00 = __ldg(p00 + c), v01 = __ldg(p01 + c), v10 = __ldg(p10 + c), v11 = __ldg(p11 + c);
        const float v = (1 - wx) * (1 - wy) * v00 + wx * (1 - wy) * v01 + (1 - wx) * wy * v10 + wx * wy * v11;
        o[c] = v / 255.0f;
    }
    r = o[2]; gch = o[1]; b = o[0];                                                         // BGR -> RGB
}

__global__ void __launch_bounds__(256)
letterbox_batch_kernel(const unsigned char* __restrict__ frames, size_t frame_stride, const int* __restrict__ wh,
                       int tw, int th, float* __restrict__ out, float* __restrict__ xform) {
    const int f = blockIdx.y;
    const int w = wh[2 * f], h = wh[2 * f + 1];
    const LetterboxGeom g = letterbox_geom(w, h, tw, th);
    if (blockIdx.x == 0 && threadIdx.x == 0 && xform) {                                     // :121-122
        xform[4 * f] = 1.0f / g.scale; xform[4 * f + 1] = 1.0f / g.scale;
        xform[4 * f + 2] = (float)g.pad_x; xform[4 * f + 3] = (float)g.pad_y;
    }
    const unsigned char* img = frames + (size_t)f * frame_stride;
    const size_t plane = (size_t)tw * th;
    float* o = out + (size_t)f * 3 * plane;
    const int quads_per_row = (tw + 3) >> 2;
    const int nquads = quads_per_row * th;
    const bool vec = (tw & 3) == 0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += gridDim.x * blockDim.x) {
        const int ty = q / quads_per_row, tx0 = (q - ty * quads_per_row) << 2;
        float r[4], gg[4], bb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            r[i] = gg[i] = bb[i] = 0.0f;
            if (tx0 + i < tw) letterbox_pixel(img, w, h, g, tx0 + i, ty, r[i], gg[i], bb[i]);
        }
        const size_t off = (size_t)ty * tw + tx0;
        if (vec) {
            __stcs(reinterpret_cast<float4*>(o + off), make_float4(r[0], r[1], r[2], r[3]));
            __stcs(reinterpret_cast<float4*>(o + plane + off), make_float4(gg[0], gg[1], gg[2], gg[3]));
            __stcs(reinterpret_cast<float4*>(o + 2 * plane + off), make_float4(bb[0], bb[1], bb[2], bb[3]));
        } else {
            for (int i = 0; i < 4 && tx0 + i < tw; ++i) { o[off + i] = r[i]; o[plane + off + i] = gg[i]; o[2 * plane + off + i] = bb[i]; }
        }
    }
}

void count_launch(int n);

}  // namespace pb

void pb_set_error(const char* fmt, ...);

extern "C" {

int pb_pose_distance(const float* d_tracks, const float* d_dets, int batch, int num_tracks, int num_dets, int mode, float alpha,
                     float* d_out_costs, pb_stream_t stream) {
    if (batch <= 0 || num_tracks <= 0 || num_dets <= 0) return PB_OK;            // oks_distance.cu:486
    if (!d_tracks || !d_dets || !d_out_costs || mode < 0 || mode > 2) { pb_set_error("pb_pose_distance: bad argument"); return PB_ERR_INVALID; }
    if (batch > 65535) { pb_set_error("pb_pose_distance: batch > 65535"); return PB_ERR_UNSUPPORTED; }
    dim3 grid((num_dets + pb::PD_TILE - 1) / pb::PD_TILE, (num_tracks + pb::PD_TILE - 1) / pb::PD_TILE, batch);
    pb::pose_distance_kernel<<<grid, dim3(pb::PD_TILE, pb::PD_TILE), 0, (cudaStream_t)stream>>>(d_tracks, d_dets, num_tracks, num_dets, mode,
                                                                                                 alpha, d_out_costs);
    pb::count_launch(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { pb_set_error("pb_pose_distance: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    return PB_OK;
}

int pb_greedy_match(const float* d_cost, int batch, int num_rows, int num_cols, float threshold, int* d_row_matched, pb_stream_t stream) {
    if (batch <= 0 || num_rows <= 0) return PB_OK;
    if (!d_cost || !d_row_matched || num_cols < 0) { pb_set_error("pb_greedy_match: bad argument"); return PB_ERR_INVALID; }
    if ((long long)num_rows * num_cols > 0x7fffffffLL) { pb_set_error("pb_greedy_match: matrix too large"); return PB_ERR_UNSUPPORTED; }
    const size_t smem = (size_t)(num_rows + num_cols + 2) * 4 + 8 * 8 + 16;
    if (smem > 200 * 1024) { pb_set_error("pb_greedy_match: rows + cols too large for shared memory"); return PB_ERR_UNSUPPORTED; }
    {
        cudaError_t e = pb::ensure_dyn_smem((const void*)pb::greedy_match_kernel, smem);
        if (e != cudaSuccess) { pb_set_error("pb_greedy_match: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    }
    pb::greedy_match_kernel<<<batch, 256, smem, (cudaStream_t)stream>>>(d_cost, num_rows, num_cols, threshold, d_row_matched);
    pb::count_launch(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { pb_set_error("pb_greedy_match: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    return PB_OK;
}

int pb_assign_legacy(const float* d_cost, int batch, int num_rows, int num_cols, float threshold, int* d_row_assign,
                     int* d_col_assign, int* d_count, pb_stream_t stream) {
    if (batch <= 0 || num_rows <= 0 || num_cols <= 0) return PB_OK;                  // hungarian.cu:243
    if (!d_cost || !d_row_assign || !d_col_assign) { pb_set_error("pb_assign_legacy: bad argument"); return PB_ERR_INVALID; }
    const size_t smem = (size_t)num_cols * 16 + (size_t)num_rows * 4 + 16;
    if (smem > 200 * 1024) { pb_set_error("pb_assign_legacy: problem too large"); return PB_ERR_UNSUPPORTED; }
    {
        cudaError_t e = pb::ensure_dyn_smem((const void*)pb::assign_legacy_kernel, smem);
        if (e != cudaSuccess) { pb_set_error("pb_assign_legacy: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    }
    pb::assign_legacy_kernel<<<batch, 256, smem, (cudaStream_t)stream>>>(d_cost, num_rows, num_cols, threshold, d_row_assign, d_col_assign, d_count);
    pb::count_launch(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { pb_set_error("pb_assign_legacy: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    return PB_OK;
}

int pb_letterbox_batch(const unsigned char* d_frames, size_t frame_stride_bytes, const int* d_sizes, int batch, int target_width,
                       int target_height, float* d_out, float* d_xform, pb_stream_t stream) {
    if (batch <= 0) return PB_OK;
    if (!d_frames || !d_sizes || !d_out || target_width <= 0 || target_height <= 0) { pb_set_error("pb_letterbox_batch: bad argument"); return PB_ERR_INVALID; }
    if (batch > 65535) { pb_set_error("pb_letterbox_batch: batch > 65535"); return PB_ERR_UNSUPPORTED; }
    const long long quads = (long long)((target_width + 3) / 4) * target_height;
    if (quads > 0x7fffffffLL) { pb_set_error("pb_letterbox_batch: target too large"); return PB_ERR_UNSUPPORTED; }
    int bx = (int)((quads + 255) / 256);
    // enough CTAs per frame to cover the 148 SMs several times over when the batch is small
    const int cap = batch >= 32 ? 64 : (2 * 148 * 8 + batch - 1) / batch;
    if (bx > cap) bx = cap;
    pb::letterbox_batch_kernel<<<dim3(bx, batch), 256, 0, (cudaStream_t)stream>>>(d_frames, frame_stride_bytes, d_sizes, target_width,
                                                                                 target_height, d_out, d_xform);
    pb::count_launch(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { pb_set_error("pb_letterbox_batch: %s", cudaGetErrorString(e)); return PB_ERR_CUDA; }
    return PB_OK;
}

}  // extern "C"
