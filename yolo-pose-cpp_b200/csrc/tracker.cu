// tracker.cu — stand-alone tracker launch (pb_tracker_update, the serial and the three-kernel pipelined step): one CTA per
// stream running tracker_body (tracker_body.cuh); shared-memory plan; state reset.
#include <cstdlib>
#include "tracker_body.cuh"

namespace pb {

static void tk_offsets(int T, int Dm, int cost_s, int det_s, int pred_s, int term_floats, int cell_cap, SmemOffsets& o) {
    TkSmem t;
    tk_carve(nullptr, T, Dm, cost_s, det_s, pred_s, term_floats, cell_cap, &t);
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
TrackerPlan tracker_plan(int T, int Dm, bool compact) {
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
    p.smem_bytes = tk_carve(nullptr, T, Dm, p.cost_in_smem, p.det_in_smem, p.pred_in_smem, term_floats, p.cell_cap, nullptr, &p.prefix_bytes);
    tk_offsets(T, Dm, p.cost_in_smem, p.det_in_smem, p.pred_in_smem, term_floats, p.cell_cap, p.so);
    const long cells = (long)T * Dm;
    // small tables (the tracker's 128 x 64 case): 1024 threads — the auction runs in one warp whatever the
    // block size, every other stage (copies, gate, cost passes, outputs) is data-parallel and measured
    // 10 % shorter than with 256 threads; mid-size tables use the CTA-wide auction, which is fastest at 512
    p.threads = cells <= 16384 ? 1024 : (cells <= 65536 ? 512 : 1024);
    if (const char* e = getenv("PB_TRACKER_THREADS")) { const int t = atoi(e); if (t == 256 || t == 512 || t == 1024) p.threads = t; }
    return p;
}

template <int NTHREADS, bool ALLSMEM>
__global__ void __launch_bounds__(NTHREADS)
pb_tracker_kernel(TrackBuffers tb, TrackParams P, DetSource src) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    tracker_body<NTHREADS, ALLSMEM, false>(tb, P, src, blockIdx.x, smem_raw, 0, P.seq, P.frame_id, tb.outputs, tb.num_outputs);
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
