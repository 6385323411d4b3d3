"""GPU tracker update (pb_tracker_update / pb_step through the C ABI) against the CPU checker:
track ids, assignments, lifecycle state bit-exact; poses, velocities, costs bit-exact too
(both sides evaluate the same IEEE operations, include/pb_math.h), which is stricter than the
1e-4 relative tolerance the contract states for floating point."""
import numpy as np
import pytest

from helpers import compare_state

pytestmark = pytest.mark.gpu


def run_sequence(pb, orc, torch, B, F, canvas=640, persons=20, period=64, occlusion=0, clumps=0, T=128, Dm=64,
                 max_age=10, min_hits=3, gating=1, stream0=0, check_state_every=1, fuse=2):
    cfg = pb.synth_config(canvas=canvas, persons=persons, period=period, occlusion=occlusion, clumps=clumps,
                          kp_drop_prob=0.15 if clumps else 0.05)
    heads = pb.synth_heads(cfg, stream0, B, 0, F, frame_major=True)
    pipe = pb.Pipeline(num_streams=B, num_anchors=cfg.num_anchors, max_tracks=T, max_detections=Dm, max_age=max_age,
                       min_hits=min_hits, gating_enabled=gating, fuse_stages=fuse)
    trk = [orc.Tracker(max_tracks=T, max_detections=Dm, max_age=max_age, min_hits=min_hits, gating_enabled=gating)
           for _ in range(B)]
    d = torch.from_numpy(heads).cuda()
    n_out = 0
    for f in range(F):
        pipe.step(d[f], f)
        torch.cuda.synchronize()
        na = pipe.get_num_active()
        for b in range(B):
            ref = orc.postprocess(heads[f, b])
            ra = trk[b].update(ref["poses"], ref["scores"], f)
            assert na[b] == ra, (f, b)
            rt, gt = trk[b].get_tracks(), pipe.get_tracks(b)
            assert gt.tobytes() == rt.tobytes(), (f, b, gt["track_id"], rt["track_id"])
            n_out += len(gt)
            if f % check_state_every == 0:
                rs = trk[b].get_state()
                bad = compare_state(pipe.get_state(b), rs, int(rs["scalars"][2]), T, f"f{f} b{b}")
                assert not bad, bad
    return n_out


@pytest.mark.parametrize("fuse", [0, 1])        # separate NMS / tracker kernels, fused per-stream kernel
def test_config1_like_sequence(pb, orc, cuda, fuse):
    assert run_sequence(pb, orc, cuda, B=3, F=48, fuse=fuse) > 1000


@pytest.mark.parametrize("fuse", [0, 1])
def test_occlusion_lost_and_recovery(pb, orc, cuda, fuse):
    assert run_sequence(pb, orc, cuda, B=4, F=90, period=90, occlusion=1, max_age=5, fuse=fuse) > 1000


def test_dense_crowd_1280(pb, orc, cuda):
    assert run_sequence(pb, orc, cuda, B=2, F=10, canvas=1280, persons=100, period=32, clumps=10, T=256, Dm=128) > 500


def test_small_tables_force_truncation_and_slot_reuse(pb, orc, cuda):
    assert run_sequence(pb, orc, cuda, B=2, F=40, persons=20, T=16, Dm=12, max_age=2, occlusion=1, period=40) > 50


@pytest.mark.parametrize("fuse", [0, 1])
def test_min_hits_one_and_gating_off(pb, orc, cuda, fuse):
    assert run_sequence(pb, orc, cuda, B=2, F=24, persons=10, min_hits=1, gating=0, fuse=fuse) > 100


def test_fused_kernel_spill_path(pb, orc, cuda, monkeypatch):
    """Fused per-stream kernel with a shared-memory tier of 64 candidates: streams with more candidates keep the NMS stage's
    per-candidate arrays in their global scratch (the spill path), the others in shared memory, in the same launch."""
    monkeypatch.setenv("PB_FUSED_TIER", "64")
    assert run_sequence(pb, orc, cuda, B=4, F=16, persons=14, fuse=1) > 300
    assert run_sequence(pb, orc, cuda, B=2, F=6, persons=4, fuse=1) > 10          # below the tier: the shared-memory path


def test_direct_detections_with_empty_and_varying_frames(pb, orc, cuda):
    """pb_tracker_update with caller-owned detection buffers (the GPUTracker::update signature),
    including frames with zero detections and a changing detection count (cost stride re-aliasing)."""
    torch = cuda
    B, T, Dm, F = 3, 64, 32, 40
    cfg = pb.synth_config(canvas=640, persons=24, period=40, occlusion=1)
    pipe = pb.Pipeline(num_streams=B, max_tracks=T, max_detections=Dm, max_age=3)
    trk = [orc.Tracker(max_tracks=T, max_detections=Dm, max_age=3) for _ in range(B)]
    stride = 40
    for f in range(F):
        poses = np.zeros((B, stride, 51), np.float32); scores = np.zeros((B, stride), np.float32)
        num = np.zeros(B, np.int32)
        for b in range(B):
            p, s = pb.synth_dets(cfg, b, f)
            if f % 7 == 3 and b == 1:
                p, s = p[:0], s[:0]                                   # empty frame
            if f % 5 == 0:
                p, s = p[: len(s) // 2], s[: len(s) // 2]
            poses[b, : len(s)] = p; scores[b, : len(s)] = s; num[b] = len(s)
        pipe.tracker_update(f, torch.from_numpy(poses).cuda(), torch.from_numpy(scores).cuda(),
                            torch.from_numpy(num).cuda(), stride)
        torch.cuda.synchronize()
        for b in range(B):
            trk[b].update(poses[b, : num[b]], scores[b, : num[b]], f)
            rs = trk[b].get_state()
            bad = compare_state(pipe.get_state(b), rs, int(rs["scalars"][2]), T, f"f{f} b{b}")
            assert not bad, bad
            assert pipe.get_tracks(b).tobytes() == trk[b].get_tracks().tobytes()


def test_tracker_stress_tables_in_global_memory(pb, orc, cuda):
    """Config 5 shape (512 x 512): the cost matrix does not fit in shared memory and is streamed
    from L2; three frames so that tracks exist and all three tiers run."""
    torch = cuda
    T = Dm = 512
    cfg = pb.synth_config(canvas=4096, persons=512, period=50)
    pipe = pb.Pipeline(num_streams=1, max_tracks=T, max_detections=Dm)
    trk = orc.Tracker(max_tracks=T, max_detections=Dm)
    for f in range(4):
        p, s = pb.synth_dets(cfg, 0, f)
        n = np.array([len(s)], np.int32)
        pipe.tracker_update(f, torch.from_numpy(p[None].copy()).cuda(), torch.from_numpy(s[None].copy()).cuda(),
                            torch.from_numpy(n).cuda(), len(s))
        torch.cuda.synchronize()
        trk.update(p, s, f)
        rs = trk.get_state()
        bad = compare_state(pipe.get_state(0), rs, int(rs["scalars"][2]), T, f"f{f}")
        assert not bad, bad
    assert len(pipe.get_tracks(0)) > 300


def test_reset_and_determinism(pb, orc, cuda):
    torch = cuda
    cfg = pb.synth_config(canvas=640, persons=12, period=16)
    heads = pb.synth_heads(cfg, 0, 2, 0, 12, frame_major=True)
    d = torch.from_numpy(heads).cuda()
    pipe = pb.Pipeline(num_streams=2)
    outs = []
    for rep in range(2):
        pipe.reset()
        acc = []
        for f in range(12):
            pipe.step(d[f], f)
            o, c = pipe.get_tracks_all()
            acc.append(c.tobytes() + b"".join(o[b, : c[b]].tobytes() for b in range(2)))
        outs.append(b"".join(acc))
    assert outs[0] == outs[1]


def feed_direct(pb, orc, torch, pipe, trks, frames_of_dets, check_every=1, T=None):
    """frames_of_dets[f][b] = (poses [n,51], scores [n]); pb_tracker_update against the checker, state dumps and records."""
    B = len(trks)
    stride = max(max(len(s) for _, s in fr) for fr in frames_of_dets) or 1
    lost_seen = recovered = 0
    prev_states = [None] * B
    for f, fr in enumerate(frames_of_dets):
        poses = np.zeros((B, stride, 51), np.float32); scores = np.zeros((B, stride), np.float32); num = np.zeros(B, np.int32)
        for b, (p, s) in enumerate(fr):
            poses[b, : len(s)] = p; scores[b, : len(s)] = s; num[b] = len(s)
        pipe.tracker_update(f, torch.from_numpy(poses).cuda(), torch.from_numpy(scores).cuda(), torch.from_numpy(num).cuda(), stride)
        torch.cuda.synchronize()
        na = pipe.get_num_active()
        for b, (p, s) in enumerate(fr):
            assert trks[b].update(p, s, f) == na[b], (f, b)
            rs = trks[b].get_state()
            st = np.where(rs["active"] == 1, rs["states"], -1)
            lost_seen += int((st == 2).sum())
            if prev_states[b] is not None:
                recovered += int(((prev_states[b] == 2) & (st == 1)).sum())
            prev_states[b] = st
            assert pipe.get_tracks(b).tobytes() == trks[b].get_tracks().tobytes(), (f, b)
            if f % check_every == 0 or f == len(frames_of_dets) - 1:
                bad = compare_state(pipe.get_state(b), rs, int(rs["scalars"][2]), T, f"f{f} b{b}")
                assert not bad, bad
    return lost_seen, recovered


@pytest.mark.parametrize("gating", [1, 0])
def test_config5_full_shape_lost_and_recovery(pb, orc, cuda, gating):
    """BASELINE config 5 at its full shape: 512 tracks x 512 detections per stream, 8 streams, spatial gating on and
    off, 36 frames with occlusion gaps and a short max-age so that CONFIRMED -> LOST, tier-3 recovery and removal
    all happen at this size (multi-chunk cost passes, CTA-wide auction, cost matrix spilled to global memory)."""
    T = Dm = 512
    B, F = 8, 36
    cfgs = [pb.synth_config(canvas=4096, persons=512, period=64, occlusion=(b & 1), max_speed=2.0, seed=0x5EEDB200 + b) for b in range(B)]
    pipe = pb.Pipeline(num_streams=B, num_anchors=64, max_candidates=64, max_keep=64, max_tracks=T, max_detections=Dm, max_age=3,
                       gating_enabled=gating)
    trks = [orc.Tracker(max_tracks=T, max_detections=Dm, max_age=3, gating_enabled=gating) for _ in range(B)]
    frames = []
    for f in range(F):
        fr = []
        for b in range(B):
            p, s = pb.synth_dets(cfgs[b], b, f)
            o = np.argsort(-s, kind="stable")
            fr.append((p[o], s[o]))
        frames.append(fr)
    lost, rec = feed_direct(pb, orc, cuda, pipe, trks, frames, check_every=6, T=T)
    assert lost > 20 and rec > 5, (lost, rec)
    assert max(len(pipe.get_tracks(b)) for b in range(B)) > 300


def test_dedup_overflow_takes_the_ordered_recompute(pb, orc, cuda):
    """More than 256 duplicate pairs in one frame (tracker.cu DUP_CAP): 40 near-identical detections start 40
    tracks (min_hits 1) whose predicted boxes all overlap in the next frame: 780 pairs, resolved with the
    sequential semantics of rule R5 by the overflow branch."""
    rng = np.random.default_rng(3)
    cfg = pb.synth_config(canvas=640, persons=1, period=8)
    base, _ = pb.synth_dets(cfg, 0, 0)
    n = 40
    frames = []
    for f in range(5):
        p = np.repeat(base[:1], n, 0).copy()
        p[:, 0::3] += rng.normal(0, 0.4, (n, 17)).astype(np.float32) + f
        p[:, 1::3] += rng.normal(0, 0.4, (n, 17)).astype(np.float32)
        s = np.sort(rng.uniform(0.5, 0.9, n).astype(np.float32))[::-1].copy()
        frames.append([(p, s), (p[:3], s[:3])])
    pipe = pb.Pipeline(num_streams=2, max_tracks=64, max_detections=48, min_hits=1)
    trks = [orc.Tracker(max_tracks=64, max_detections=48, min_hits=1) for _ in range(2)]
    feed_direct(pb, orc, cuda, pipe, trks, frames, T=64)
    st = trks[0].get_state()
    assert 1 <= int(st["active"].sum()) < n              # the duplicates were removed


def test_cell_list_clamp_is_unreachable(pb, cuda):
    """CELL_LIST_CAP (4096 gated cells per chunk of active rows) can only be exceeded by a single row with more than
    4096 detections; such a table does not fit the tracker's shared memory and pb_create refuses it."""
    with pytest.raises(pb.PbError) as e:
        pb.Pipeline(num_streams=1, max_tracks=4, max_detections=4097)
    assert e.value.status == pb.PB_ERR_UNSUPPORTED


@pytest.mark.parametrize("env", [{"PB_NO_BULK": "1"}, {"PB_NO_BULK": "2"}, {"PB_NO_BULK": "3"}, {"PB_NO_SUB_SOLVE": "1"}])
def test_ab_switches_keep_the_results(pb, orc, cuda, monkeypatch, env):
    """The stand-alone tracker kernel stages a stream's state slabs by bulk asynchronous copies (all of them, only the centres
    and the cost matrix, only the per-slot slabs, or element by element: PB_NO_BULK) and, on larger tables, solves tiers 2 and
    3 on the unmatched sub-problem (PB_NO_SUB_SOLVE turns that off): every variant against the checker, on the tracker's own
    table size with occlusion gaps (LOST tracks, tier-3 recoveries) and on a crowd (wide solve, sub-problem solve)."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    assert run_sequence(pb, orc, cuda, B=5, F=24, occlusion=1, max_age=4, fuse=0) > 0
    assert run_sequence(pb, orc, cuda, B=2, F=8, canvas=1280, persons=100, clumps=10, T=256, Dm=128, check_state_every=4, fuse=0) > 0
