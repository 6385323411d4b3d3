// tracker.cu — stand-alone tracker launch (pb_tracker_update, the serial and the three-kernel pipelined step): one CTA per
// stream running tracker_body (tracker_body.cuh); shared-memory plan; state reset.
#include <cstdlib>
#include "tracker_body.cuh"

namespace pb {

static void tk_offsets(int T, int Dm, int cost_s, int det_s, int pred_s, int term_floats, int cell_cap, SmemOffsets& o, int res_s = 0) {
    TkSmem t;
    tk_carve(nullptr, T, Dm, cost_s, det_s, pred_s, term_floats, cell_cap, &t, nullptr, res_s);
    o.off[36] = (unsigned)(uintptr_t)t.poses;
    o.off[37] = (unsigned)(uintptr_t)t.vel;
    o.off[38] = (unsigned)(uintptr_t)t.dirty;
    o.off[0] = (unsigned)(uintptr_t)t.active;
    o.off[1] = (unsigned)(uintptr_t)t.states;
    o.off[2] = (unsigned)(uintptr_t)t.hits;
    o.off[3] = (unsigned)(uintptr_t)t.ids;
    o.off[4] = (unsigned)(uintptr_t)t.ages;
    o.off[5] = (unsigned)(uintptr_t)t.row;
    o.off[6] = (unsigned)(uintptr_t)t.rowb;
    o.off[7] = (unsigned)(uintptr_t)t.act_list;
    o.off[8] = (unsigned)(uintptr_t)t.elig_list;
    o.off[9] = (unsigned)(uintptr_t)t.rowbc;
    o.off[10] = (unsigned)(uintptr_t)t.rowbid;
    o.off[11] = (unsigned)(uintptr_t)t.col;
    o.off[12] = (unsigned)(uintptr_t)t.colb;
    o.off[13] = (unsigned)(uintptr_t)t.slot_for_det;
    o.off[14] = (unsigned)(uintptr_t)t.out_list;
    o.off[15] = (unsigned)(uintptr_t)t.price;
    o.off[16] = (unsigned)(uintptr_t)t.dscore;
    o.off[17] = (unsigned)(uintptr_t)t.darea;
    o.off[18] = (unsigned)(uintptr_t)t.colbid;
    o.off[19] = (unsigned)(uintptr_t)t.tcent;
    o.off[20] = (unsigned)(uintptr_t)t.tarea;
    o.off[21] = (unsigned)(uintptr_t)t.tav;
    o.off[22] = (unsigned)(uintptr_t)t.dcent;
    o.off[23] = (unsigned)(uintptr_t)t.gate;
    o.off[24] = (unsigned)(uintptr_t)t.lgate;
    o.off[25] = (unsigned)(uintptr_t)t.colmask;
    o.off[26] = (unsigned)(uintptr_t)t.dup;
    o.off[27] = (unsigned)(uintptr_t)t.misc;
    o.off[28] = (unsigned)(uintptr_t)t.acc;
    o.off[29] = (unsigned)(uintptr_t)t.terms;
    o.off[30] = (unsigned)(uintptr_t)t.sig;
    o.off[31] = (unsigned)(uintptr_t)t.cell_list;
    o.off[32] = (unsigned)(uintptr_t)t.aowner;
    o.off[33] = (unsigned)(uintptr_t)t.cost;
    o.off[34] = (unsigned)(uintptr_t)t.det;
    o.off[35] = (unsigned)(uintptr_t)t.pred;
}

// compact: the plan of the fused per-stream kernel, which wants two CTAs per SM — a term buffer of 256 cells per
// round instead of 481+ and a cell list of 2048 instead of 4096 entries (more rounds for large frames, same results:
// cells are independent).  The term buffer also holds the compacted cost rows of the single-warp auction
// (at most 32 rows x 64 columns).
// resident: the layout of the resident tracker (pb_tracker_seq_kernel) — everything in shared memory, plus the stream's poses,
// velocities and dirty flags; plan.resident stays 0 when that does not fit.
TrackerPlan tracker_plan(int T, int Dm, bool compact, bool resident) {
    TrackerPlan p{};
    const size_t budget = 200 * 1024;
    p.cost_in_smem = p.det_in_smem = p.pred_in_smem = 0;
    // term buffer: at least one active row (Dm * 17 floats), 32 KB when it fits
    int term_floats = Dm * KP;
    const int floor_floats = compact ? 4352 : 8192;
    if (term_floats < floor_floats) term_floats = floor_floats;
    p.term_floats = term_floats;
    p.cell_cap = compact ? 2048 : 4096;
    if (p.cell_cap < Dm) p.cell_cap = Dm;
    size_t base = tk_carve(nullptr, T, Dm, 0, 0, 0, term_floats, p.cell_cap, nullptr);
    size_t cost_b = tk_align((size_t)T * Dm * 4), det_b = tk_align((size_t)Dm * POSE_F * 4), pred_b = tk_align((size_t)T * POSE_F * 4);
    size_t used = base;
    if (used + cost_b <= budget) { p.cost_in_smem = 1; used += cost_b; }
    if (used + det_b <= budget) { p.det_in_smem = 1; used += det_b; }
    if (used + pred_b <= budget) { p.pred_in_smem = 1; used += pred_b; }
    p.resident = 0;
    if (resident && p.cost_in_smem && p.det_in_smem && p.pred_in_smem &&
        tk_carve(nullptr, T, Dm, 1, 1, 1, term_floats, p.cell_cap, nullptr, nullptr, 1) <= 224 * 1024) p.resident = 1;
    p.smem_bytes = tk_carve(nullptr, T, Dm, p.cost_in_smem, p.det_in_smem, p.pred_in_smem, term_floats, p.cell_cap, nullptr, &p.prefix_bytes, p.resident);
    tk_offsets(T, Dm, p.cost_in_smem, p.det_in_smem, p.pred_in_smem, term_floats, p.cell_cap, p.so, p.resident);
    const long cells = (long)T * Dm;
    // small tables (the tracker's 128 x 64 case): 1024 threads — the auction runs in one warp whatever the
    // block size, every other stage (copies, gate, cost passes, outputs) is data-parallel and measured
    // 10 % shorter than with 256 threads; mid-size tables use the CTA-wide auction, which is fastest at 512
    p.threads = cells <= 16384 ? 1024 : (cells <= 65536 ? 512 : 1024);
    if (const char* e = getenv("PB_TRACKER_THREADS")) { const int t = atoi(e); if (t == 256 || t == 512 || t == 1024) p.threads = t; }
    return p;
}

// =======================================================================================
// Large tables (the 512 x 512 stress configuration): predict, keypoint-box centres, gates and the tier-1 OKS cost pass are
// row-parallel and dominate the frame (gate + cost pass: 60 % of a 512 x 512 frame on one SM).  This kernel runs them with
// plan.pre_slices CTAs per stream, each owning a slice of track slots, ahead of the one-CTA-per-stream kernel, which then
// takes the centres, areas, gate words and costs from global memory and goes straight to the auction.  Same expressions,
// same results.  Only for launches that are ordered on one CUDA stream (the state must be the previous frame's final one).
// =======================================================================================
constexpr int PRE_THREADS = 512;

__global__ void __launch_bounds__(PRE_THREADS)
pb_tracker_pre_kernel(TrackBuffers tb, TrackParams P, DetSource src, int rows_per_slice) {
    extern __shared__ __align__(16) unsigned char pre_smem[];
    const int b = blockIdx.y, T = P.T, Dm = P.Dm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NT = PRE_THREADS, nwarps = PRE_THREADS / 32;
    const int r0 = blockIdx.x * rows_per_slice;
    const int r1 = (r0 + rows_per_slice < T) ? r0 + rows_per_slice : T;
    const int nr = r1 > r0 ? r1 - r0 : 0;
    int n_in = src.num[b];
    const int D = n_in < Dm ? (n_in < 0 ? 0 : n_in) : Dm;
    const int Dw = (Dm + 31) / 32, words = (D + 31) / 32;
    // shared memory: det poses [Dm*51], dcent [Dm*4], darea [Dm], pred slice [rps*51], tcent [rps*4], tarea, tav, flags [rps], gate words [rps*Dw], cell list [rps*Dm]
    float* s_det = reinterpret_cast<float*>(pre_smem);
    float* s_dcent = s_det + (size_t)Dm * POSE_F;
    float* s_darea = s_dcent + (size_t)Dm * 4;
    float* s_pred = s_darea + Dm;
    float* s_tcent = s_pred + (size_t)rows_per_slice * POSE_F;
    float* s_tarea = s_tcent + (size_t)rows_per_slice * 4;
    float* s_tav = s_tarea + rows_per_slice;
    int* s_act = reinterpret_cast<int*>(s_tav + rows_per_slice);          // 1 active, 2 active and LOST
    unsigned* s_gate = reinterpret_cast<unsigned*>(s_act + rows_per_slice);   // [rps * Dw] tier-1 gate words of the slice
    __shared__ float s_sig[KP];
    float* g_poses = tb.poses + (size_t)b * T * POSE_F;
    float* g_vel = tb.vel + (size_t)b * T * 34;
    float* g_pred = tb.predicted + (size_t)b * T * POSE_F;
    float* g_tcent = tb.tcent + (size_t)b * T * 4;
    float* g_cost = tb.cost + (size_t)b * T * Dm;
    const int* g_states = tb.states + (size_t)b * T;
    const int* g_active = tb.active + (size_t)b * T;
    int* g_dirty = tb.pred_dirty + (size_t)b * T;
    unsigned* g_gate = tb.gate_g + (size_t)b * T * Dw;
    unsigned* g_lgate = tb.lgate_g + (size_t)b * T * Dw;
    float* g_tarea = tb.tarea_g + (size_t)b * T;
    const int na_total = tb.scalars[(size_t)b * 4 + 3];                   // active slots at frame start (= the count the previous frame left)
    if (tid < KP) s_sig[tid] = kSigmas[tid];
    for (int r = tid; r < nr; r += NT) {
        const int t = r0 + r;
        s_act[r] = (g_active[t] == 1) ? ((g_states[t] == ST_LOST) ? 2 : 1) : 0;
    }
    __syncthreads();
    // predict (:102-138): the slice's active rows
#pragma unroll 1
    for (int i = tid; i < nr * KP; i += NT) {
        const int r = i / KP, k = i - r * KP;
        if (s_act[r] == 0) continue;
        const int t = r0 + r;
        const int po = t * POSE_F + k * 3, vo = t * 34 + k * 2;
        const float x = g_poses[po], y = g_poses[po + 1], cf = g_poses[po + 2];
        const float vx = g_vel[vo], vy = g_vel[vo + 1];
        const float dt = 1.0f;
        const float px = x + vx * dt, py = y + vy * dt;
        g_pred[po] = px; g_pred[po + 1] = py; g_pred[po + 2] = cf;
        s_pred[r * POSE_F + k * 3] = px; s_pred[r * POSE_F + k * 3 + 1] = py; s_pred[r * POSE_F + k * 3 + 2] = cf;
        if (s_act[r] == 2) { g_vel[vo] = vx * 0.95f; g_vel[vo + 1] = vy * 0.95f; }
        if (k == 0) g_dirty[t] = 1;
    }
    const bool assoc12 = (na_total > 0) && (D > 0);
    if (!assoc12) return;                                                 // uniform over the grid's CTAs of this stream
    // this frame's detections (all of them, every slice needs them): poses, centres, areas
    const float* src_pose = src.poses + (size_t)b * src.stride * POSE_F;
#pragma unroll 1
    for (int i = tid; i < D * POSE_F; i += NT) s_det[i] = src_pose[i];
    __syncthreads();                                                      // predict's stores (global, velocity decay) and s_pred, s_det
#pragma unroll 1
    for (int d = tid; d < D; d += NT) { float area; pose_box(s_det + (size_t)d * POSE_F, &s_dcent[d * 4], &area); s_darea[d] = area; }
    // centres (:196-237): every slot of the slice whose predicted pose changed since its centre was derived
#pragma unroll 1
    for (int r = tid; r < nr; r += NT) {
        const int t = r0 + r;
        if (g_dirty[t] || s_act[r] != 0) {
            const float* pp = (s_act[r] != 0) ? (s_pred + (size_t)r * POSE_F) : (g_pred + (size_t)t * POSE_F);
            float area;
            pose_box(pp, &s_tcent[r * 4], &area);
            s_tarea[r] = area;
            g_tarea[t] = area;
            g_tcent[t * 4] = s_tcent[r * 4]; g_tcent[t * 4 + 1] = s_tcent[r * 4 + 1];
            g_tcent[t * 4 + 2] = s_tcent[r * 4 + 2]; g_tcent[t * 4 + 3] = s_tcent[r * 4 + 3];
            g_dirty[t] = 0;
        }
        if (s_act[r] != 0) {                                              // mean torso speed (:287-298), after the LOST decay
            const int torso[4] = {5, 6, 11, 12};
            float av = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float vx = g_vel[t * 34 + torso[i] * 2], vy = g_vel[t * 34 + torso[i] * 2 + 1];
                av += sqrtf(vx * vx + vy * vy);
            }
            s_tav[r] = av * 0.25f;
        }
    }
    __syncthreads();
    // gates (:241-317): one warp per (row of the slice, 32 detections); every word of the slice is written (0: inactive row)
#pragma unroll 1
    for (int i = warp; i < nr * Dw; i += nwarps) {
        const int r = i / Dw, w = i - r * Dw;
        const int t = r0 + r;
        const int d = w * 32 + lane;
        const bool act = s_act[r] != 0, lost = s_act[r] == 2;
        const float base = lost ? (3.0f * 1.3f) : 3.0f;
        bool g = false;
        if (act && w < words && d < D) g = gate_cell(&s_tcent[r * 4], &s_dcent[d * 4], s_tav[r], lost, base, P.gating_enabled != 0);
        const unsigned bm = __ballot_sync(0xffffffffu, g);
        if (lane == 0) {
            g_gate[t * Dw + w] = lost ? 0u : bm;
            g_lgate[t * Dw + w] = lost ? bm : 0u;
            s_gate[r * Dw + w] = lost ? 0u : bm;
        }
    }
    __syncthreads();
    // tier-1 cost pass (:333-425) on the gated cells of the slice's rows that are not LOST.  The gated cells (a sixth of the
    // table with the gate on) are compacted first — warp-aggregated appends, any order: cells are independent — so that the
    // 17-exponential loop runs on dense lanes.
    int* s_cells = reinterpret_cast<int*>(s_gate + (size_t)rows_per_slice * Dw);      // [rps * Dm]
    __shared__ int s_ncell;
    if (tid == 0) s_ncell = 0;
    __syncthreads();
#pragma unroll 1
    for (int i = warp; i < nr * words; i += nwarps) {
        const int r = i / words, w = i - r * words;
        const unsigned gw = s_gate[r * Dw + w];
        if (gw == 0u) continue;
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_ncell, __popc(gw));
        base = __shfl_sync(0xffffffffu, base, 0);
        if ((gw >> lane) & 1u) s_cells[base + __popc(gw & ((1u << lane) - 1u))] = (r << 16) | (w * 32 + lane);
    }
    __syncthreads();
    const int ncell = s_ncell;
#pragma unroll 1
    for (int i = tid; i < ncell; i += NT) {
        const int r = s_cells[i] >> 16, d = s_cells[i] & 0xffff;
        g_cost[(size_t)(r0 + r) * D + d] = oks_cell_cost(s_pred + (size_t)r * POSE_F, s_det + (size_t)d * POSE_F, s_tarea[r], s_darea[d], s_sig, 0.2f);
    }
}

static size_t pre_smem_bytes(int Dm, int rps) {
    const int Dw = (Dm + 31) / 32;
    return ((size_t)Dm * (POSE_F + 4 + 1) + (size_t)rps * (POSE_F + 4 + 1 + 1 + 1) + (size_t)rps * Dw + (size_t)rps * Dm) * 4 + 64;
}

// Decide whether (and how) the pre-kernel is used for a handle: tables of at least 64 K cells whose detections fit the
// pre-kernel's shared memory; slices so that the grid covers the device about twice.
void tracker_plan_pre(TrackerPlan& plan, int B, int T, int Dm, int sm_count, size_t smem_optin) {
    plan.pre_slices = 0; plan.pre_smem = 0;
    long min_cells = 65536;
    if (const char* e = getenv("PB_TRACKER_PRE")) { const long v = atol(e); if (v <= 0) return; min_cells = v; }
    if ((long)T * Dm < min_cells) return;
    int G = (2 * sm_count + B - 1) / B;
    if (G < 1) G = 1;
    if (G > (T + 7) / 8) G = (T + 7) / 8;
    const int rps = (T + G - 1) / G;
    G = (T + rps - 1) / rps;
    const size_t bytes = pre_smem_bytes(Dm, rps);
    if (bytes > smem_optin) return;
    plan.pre_slices = G; plan.pre_smem = bytes;
}

cudaError_t launch_tracker_pre(const TrackBuffers& tb, const TrackParams& p, const DetSource& src, const TrackerPlan& plan, cudaStream_t stream) {
    const cudaError_t e = ensure_dyn_smem((const void*)pb_tracker_pre_kernel, plan.pre_smem);
    if (e != cudaSuccess) return e;
    const int rps = (p.T + plan.pre_slices - 1) / plan.pre_slices;
    pb_tracker_pre_kernel<<<dim3(plan.pre_slices, p.B), PRE_THREADS, plan.pre_smem, stream>>>(tb, p, src, rps);
    count_launch();
    return cudaGetLastError();
}

template <int NTHREADS, bool ALLSMEM>
#ifdef PB_TRK_MAXNREG
__global__ void __maxnreg__(PB_TRK_MAXNREG)
#else
__global__ void __launch_bounds__(NTHREADS)
#endif
pb_tracker_kernel(TrackBuffers tb, TrackParams P, DetSource src) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    tracker_body<NTHREADS, ALLSMEM, false, false, ALLSMEM>(tb, P, src, blockIdx.x, smem_raw, 0, P.seq, P.frame_id, tb.outputs, tb.num_outputs);
}

// =======================================================================================
// Resident tracker: the tracker stages of Q.n consecutive frames of every video stream in ONE launch, one CTA per stream
// (the per-stream pipeline of the north star: a stream's CTA stays on its SM while the decode and NMS kernels of the following
// steps feed it).  Frame i is run as soon as the NMS kernel of its step has published the stream's kept detections
// (PostBuffers::ready, release / acquire per stream).  Streams never wait for each other; a frame whose auction runs to the
// iteration limit delays only the later frames of its own stream; frames 1.. of a launch need no hand-over between CTAs.
// The only device-side wait is for a kernel that was enqueued without any dependency on this one (pb_api.cu), so it cannot
// deadlock while the grid leaves SMs free for that kernel (the host layer checks); a 0.5 s time-out guards against misuse.
// =======================================================================================
template <int NTHREADS, bool ALLSMEM, bool RES>
#ifdef PB_TRK_MAXNREG
__global__ void __maxnreg__(PB_TRK_MAXNREG)
#else
__global__ void __launch_bounds__(NTHREADS, (ALLSMEM && NTHREADS <= 512) ? 1024 / NTHREADS : 1)   // at most 64 registers: room for other CTAs beside it
#endif
pb_tracker_seq_kernel(const __grid_constant__ TrackBuffers tb, const __grid_constant__ TrackParams P, const __grid_constant__ SeqTable Q) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_ok;
    const int b = blockIdx.x;
    unsigned long long waited = 0ull;                // telemetry: time spent waiting for the NMS kernels (thread 0)
#pragma unroll 1
    for (int i = 0; i < Q.n; ++i) {
        const int seq = P.seq + i;
        if (threadIdx.x == 0) {
            const int* flag = Q.ready[i] + b;
            const unsigned long long w0 = globaltimer_ns();
            unsigned long long w1 = w0;
            int ok = 1;
            for (;;) {
                int v;
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                if (v == seq) break;
                w1 = globaltimer_ns();
                if (w1 - w0 > 500000000ull) { ok = 0; break; }
                __nanosleep(64);
            }
            waited += w1 - w0;
            s_ok = ok;
        }
        __syncthreads();
        int r = -1;
        if (s_ok) {
            const DetSource src{Q.det_poses[i], Q.det_scores[i], Q.num_keep[i], Q.stride};
            r = tracker_body<NTHREADS, ALLSMEM, false, RES>(tb, P, src, b, smem_raw, 0, seq, P.frame_id + i, Q.outputs[i], Q.num_outputs[i], i == 0 ? 0 : 2,
                                                            true, i == 0, i == Q.n - 1);
        }
        if (r < 0) {
            // time-out (here or inside the frame): the state stays as it is, the sequence numbers of the launch are passed on
            // and the sticky error flag invalidates the results (every synchronising entry point reports it)
            if (threadIdx.x == 0) {
                atomicExch(tb.error_flag, 1);
                const int last = P.seq + Q.n - 1;
                __threadfence();
                asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(tb.seq_done + b), "r"(last) : "memory");
                asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(tb.out_done + b), "r"(last) : "memory");
                tb.chain[b] = chain_pack(0u, last + 1, 0);
            }
            return;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) tb.stage_ns[(size_t)b * 20 + 15] += waited;      // (slot 15: waits for a predecessor / for the detections)
}

cudaError_t launch_tracker_seq(const TrackBuffers& tb, TrackParams p, const SeqTable& q, const TrackerPlan& plan, cudaStream_t stream) {
    const bool allsmem = plan.cost_in_smem && plan.det_in_smem && plan.pred_in_smem;
    // variants: 1024 threads (one CTA per SM) or 512 threads (two per SM) with everything in shared memory, state resident
    // (plan.resident) or reloaded per frame; 256 / 512 / 1024 threads with parts of the working set in global memory
    int v;
    if (allsmem && plan.threads == 1024) v = plan.resident ? 5 : 3;
    else if (allsmem && plan.threads == 512) v = plan.resident ? 6 : 4;
    else if (allsmem && plan.threads == 256 && plan.resident) v = 7;
    else v = plan.threads == 256 ? 0 : (plan.threads == 512 ? 1 : 2);
    const void* fn = v == 0 ? (const void*)pb_tracker_seq_kernel<256, false, false> : v == 1 ? (const void*)pb_tracker_seq_kernel<512, false, false>
                   : v == 2 ? (const void*)pb_tracker_seq_kernel<1024, false, false> : v == 3 ? (const void*)pb_tracker_seq_kernel<1024, true, false>
                   : v == 4 ? (const void*)pb_tracker_seq_kernel<512, true, false> : v == 5 ? (const void*)pb_tracker_seq_kernel<1024, true, true>
                   : v == 6 ? (const void*)pb_tracker_seq_kernel<512, true, true> : (const void*)pb_tracker_seq_kernel<256, true, true>;
    {
        const cudaError_t e = ensure_dyn_smem(fn, plan.smem_bytes);
        if (e != cudaSuccess) return e;
    }
    p.cost_in_smem = plan.cost_in_smem; p.det_in_smem = plan.det_in_smem; p.pred_in_smem = plan.pred_in_smem;
    p.term_floats = plan.term_floats; p.cell_cap = plan.cell_cap;
    p.so = plan.so;
    p.precomputed = 0;
    switch (v) {
        case 0: pb_tracker_seq_kernel<256, false, false><<<p.B, 256, plan.smem_bytes, stream>>>(tb, p, q); break;
        case 1: pb_tracker_seq_kernel<512, false, false><<<p.B, 512, plan.smem_bytes, stream>>>(tb, p, q); break;
        case 2: pb_tracker_seq_kernel<1024, false, false><<<p.B, 1024, plan.smem_bytes, stream>>>(tb, p, q); break;
        case 3: pb_tracker_seq_kernel<1024, true, false><<<p.B, 1024, plan.smem_bytes, stream>>>(tb, p, q); break;
        case 4: pb_tracker_seq_kernel<512, true, false><<<p.B, 512, plan.smem_bytes, stream>>>(tb, p, q); break;
        case 5: pb_tracker_seq_kernel<1024, true, true><<<p.B, 1024, plan.smem_bytes, stream>>>(tb, p, q); break;
        case 6: pb_tracker_seq_kernel<512, true, true><<<p.B, 512, plan.smem_bytes, stream>>>(tb, p, q); break;
        default: pb_tracker_seq_kernel<256, true, true><<<p.B, 256, plan.smem_bytes, stream>>>(tb, p, q); break;
    }
    count_launch();
    return cudaGetLastError();
}

__global__ void pb_tracker_reset_kernel(TrackBuffers tb, int B, int T, int Dm, int seq) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nT = (size_t)B * T, nD = (size_t)B * Dm;
    if (i < nT * POSE_F) { tb.poses[i] = 0.f; tb.predicted[i] = 0.f; }
    if (i < nT * 34) tb.vel[i] = 0.f;
    if (i < nT * 4) tb.tcent[i] = 0.f;
    if (i < nT * Dm) tb.cost[i] = 0.f;
    if (i < nT) {
        tb.scores[i] = 0.f; tb.states[i] = 0; tb.ids[i] = 0; tb.hits[i] = 0; tb.ages[i] = 0;
        tb.last_frame[i] = 0; tb.active[i] = 0; tb.pred_dirty[i] = 0; tb.row_assign[i] = -1;
    }
    if (i < nD) { tb.det_scores[i] = 0.f; tb.col_assign[i] = -1; }
    if (i < nD * 4) tb.dcent[i] = 0.f;
    if (i < (size_t)B) {
        tb.scalars[i * 4 + 0] = 1; tb.scalars[i * 4 + 1] = 0; tb.scalars[i * 4 + 2] = 0; tb.scalars[i * 4 + 3] = 0;
        tb.num_outputs[i] = 0;
        tb.seq_done[i] = seq; tb.out_done[i] = seq;
        tb.chain[i] = chain_pack(0u, seq + 1, 0);
    }
    if (i == 0) *tb.error_flag = 0;
    if (i < (size_t)B * 20) tb.stage_ns[i] = 0ull;
}

cudaError_t launch_tracker_reset(const TrackBuffers& tb, int B, int T, int Dm, int seq, cudaStream_t stream) {
    size_t n = (size_t)B * T * (size_t)(Dm > POSE_F ? Dm : POSE_F);
    if (n < (size_t)B * Dm * 4) n = (size_t)B * Dm * 4;
    if (n < (size_t)B * 20) n = (size_t)B * 20;
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    pb_tracker_reset_kernel<<<blocks, threads, 0, stream>>>(tb, B, T, Dm, seq);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_tracker(const TrackBuffers& tb, TrackParams p, const DetSource& src,
                           const TrackerPlan& plan, cudaStream_t stream) {
    const bool allsmem = plan.cost_in_smem && plan.det_in_smem && plan.pred_in_smem;
    const int v = plan.threads == 256 ? 0 : (plan.threads == 512 ? 1 : (allsmem ? 3 : 2));
    const void* fn = v == 0 ? (const void*)pb_tracker_kernel<256, false> : v == 1 ? (const void*)pb_tracker_kernel<512, false>
                   : v == 2 ? (const void*)pb_tracker_kernel<1024, false> : (const void*)pb_tracker_kernel<1024, true>;
    {
        const cudaError_t e = ensure_dyn_smem(fn, plan.smem_bytes);
        if (e != cudaSuccess) return e;
    }
    p.cost_in_smem = plan.cost_in_smem; p.det_in_smem = plan.det_in_smem; p.pred_in_smem = plan.pred_in_smem;
    p.term_floats = plan.term_floats; p.cell_cap = plan.cell_cap;
    p.so = plan.so;
    if (v == 0) pb_tracker_kernel<256, false><<<p.B, 256, plan.smem_bytes, stream>>>(tb, p, src);
    else if (v == 1) pb_tracker_kernel<512, false><<<p.B, 512, plan.smem_bytes, stream>>>(tb, p, src);
    else if (v == 2) pb_tracker_kernel<1024, false><<<p.B, 1024, plan.smem_bytes, stream>>>(tb, p, src);
    else pb_tracker_kernel<1024, true><<<p.B, 1024, plan.smem_bytes, stream>>>(tb, p, src);
    count_launch();
    return cudaGetLastError();
}

}  // namespace pb
