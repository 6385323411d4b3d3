"""SURVEY.md §8f rows on the GPU: OKSDistanceCUDA entry points, deterministic greedy matcher,
un-letterboxing in the output stage, tracker state snapshot — CUDA against the checker (bit-exact)
and the checker against the reference's own code where it is deterministic."""
import numpy as np
import pytest

from test_oracle_next_rows import rand_poses

pytestmark = pytest.mark.gpu


def test_pose_distance_cuda_equals_checker_and_reference(pb, orc, cuda):
    torch = cuda
    rng = np.random.default_rng(8)
    for nt, nd in [(1, 1), (9, 13), (50, 100), (33, 17)]:
        batch = 3
        t = np.stack([rand_poses(rng, nt) for _ in range(batch)]); d = np.stack([rand_poses(rng, nd) for _ in range(batch)])
        d[:, : min(nt, nd), :, :2] = t[:, : min(nt, nd), :, :2] + rng.normal(0, 4, (batch, min(nt, nd), 17, 2)).astype(np.float32)
        dt, dd = torch.from_numpy(t.reshape(batch, nt, 51)).cuda(), torch.from_numpy(d.reshape(batch, nd, 51)).cuda()
        for mode, alpha in ((0, 0.0), (1, 0.0), (2, 0.7), (2, 0.25)):
            got = pb.pose_distance(dt, dd, mode, alpha).cpu().numpy()
            for b in range(batch):
                want = orc.pose_distance(t[b], d[b], mode, alpha)
                assert got[b].tobytes() == want.tobytes(), (nt, nd, mode, b)
    import ref_py
    if ref_py.available():                       # the reference's own kernels on the same poses
        t, d = rand_poses(rng, 20), rand_poses(rng, 30)
        d[:10, :, :2] = t[:10, :, :2] + rng.normal(0, 5, (10, 17, 2)).astype(np.float32)
        for mode, alpha in ((0, 0.0), (1, 0.0), (2, 0.7)):
            ref = ref_py.pose_distance(t, d, mode, alpha)
            want = orc.pose_distance(t, d, mode, alpha)
            assert np.abs(ref - want).max() < 1e-4, (mode, np.abs(ref - want).max())


def test_greedy_match_cuda_equals_checker_and_reference_host_rule(pb, orc, cuda):
    torch = cuda
    rng = np.random.default_rng(4)
    for R, C in [(1, 1), (5, 5), (12, 7), (7, 12), (128, 64), (200, 300)]:
        cost = rng.uniform(0, 1, (4, R, C)).astype(np.float32)
        cost[rng.uniform(0, 1, cost.shape) < 0.1] = 0.25
        got = pb.greedy_match(torch.from_numpy(cost).cuda(), 0.6).cpu().numpy()
        for b in range(4):
            assert np.array_equal(got[b], orc.greedy_match(cost[b], 0.6)), (R, C, b)
    import ref_py
    if ref_py.available():                       # GreedyMatcherCUDA::match takes its deterministic host path below 200 cells
        for R, C in [(5, 5), (12, 7), (7, 12), (13, 15)]:
            cost = rng.uniform(0, 1, (R, C)).astype(np.float32)
            assert np.array_equal(ref_py.greedy_match(cost, 0.6), orc.greedy_match(cost, 0.6)), (R, C)


def test_output_transform_is_scale_track_outputs(pb, orc, cuda):
    torch = cuda
    B, F = 2, 8
    scfg = pb.synth_config(canvas=640, persons=8, period=32)
    heads = pb.synth_heads(scfg, 3, B, 0, F, frame_major=True)
    d = torch.from_numpy(heads).cuda()
    plain, scaled = pb.Pipeline(num_streams=B), pb.Pipeline(num_streams=B)
    xf = np.array([[1.5, 1.5, 0, 80], [0.75, 2.0, 13, 7]], np.float32)       # scale_x, scale_y, pad_x, pad_y
    scaled.set_output_transform(xf)
    for f in range(F):
        plain.step(d[f], f); scaled.step(d[f], f)
    for b in range(B):
        a, s = plain.get_tracks(b), scaled.get_tracks(b)
        assert len(a) == len(s) > 0 and np.array_equal(a["track_id"], s["track_id"]) and a["score"].tobytes() == s["score"].tobytes()
        sx, sy, px, py = xf[b]
        want_kp = a["keypoints"].copy()
        want_kp[:, :, 0] = (a["keypoints"][:, :, 0] - px) * sx; want_kp[:, :, 1] = (a["keypoints"][:, :, 1] - py) * sy
        want_bb = a["bbox"].copy()
        want_bb[:, 0] = (a["bbox"][:, 0] - px) * sx; want_bb[:, 2] = (a["bbox"][:, 2] - px) * sx
        want_bb[:, 1] = (a["bbox"][:, 1] - py) * sy; want_bb[:, 3] = (a["bbox"][:, 3] - py) * sy
        assert s["keypoints"].tobytes() == want_kp.tobytes() and s["bbox"].tobytes() == want_bb.tobytes()
    scaled.set_output_transform(None)
    plain.step(d[0], F); scaled.step(d[0], F)
    assert plain.get_tracks(0).tobytes() == scaled.get_tracks(0).tobytes()


def test_state_snapshot_restores_exact_continuation(pb, cuda):
    torch = cuda
    B, F = 3, 30
    scfg = pb.synth_config(canvas=640, persons=10, period=64, occlusion=1)
    d = torch.from_numpy(pb.synth_heads(scfg, 20, B, 0, F, frame_major=True)).cuda()
    a = pb.Pipeline(num_streams=B, max_age=4)
    for f in range(15):
        a.step(d[f], f)
    blob = a.state_save()
    b = pb.Pipeline(num_streams=B, max_age=4)
    b.state_load(blob)
    for f in range(15, F):
        a.step(d[f], f); b.step(d[f], f)
        oa, ca = a.get_tracks_all(); ob, cb = b.get_tracks_all()
        assert np.array_equal(ca, cb)
        for s in range(B):
            assert oa[s, : ca[s]].tobytes() == ob[s, : cb[s]].tobytes()
    for s in range(B):
        sa, sb = a.get_state(s), b.get_state(s)
        for k in sa:
            assert sa[k].tobytes() == sb[k].tobytes(), k
    with pytest.raises(pb.PbError):
        pb.Pipeline(num_streams=B + 1).state_load(blob)
