"""Stage-level entry points on the GPU against the CPU checker: stand-alone auction, 3rd-order
Kalman filter, launchPoseNMS and the host-legacy NMS rule set."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIG = np.array([0.026, 0.025, 0.025, 0.035, 0.035, 0.079, 0.079, 0.072, 0.072, 0.062, 0.062, 0.107, 0.107,
                0.087, 0.087, 0.089, 0.089], np.float32)


def test_auction_batch(pb, orc, cuda):
    torch = cuda
    rng = np.random.default_rng(3)
    for R, C, batch in [(50, 50, 4), (128, 64, 3), (20, 37, 5), (512, 512, 2), (7, 1, 2)]:
        cost = rng.uniform(0, 1, (batch, R, C)).astype(np.float32)
        cost[rng.uniform(size=cost.shape) < 0.2] = 1e9
        act = (rng.uniform(size=(batch, R)) < 0.8).astype(np.int32)
        d_cost, d_act = torch.from_numpy(cost).cuda(), torch.from_numpy(act).cuda()
        d_row = torch.empty(batch, R, dtype=torch.int32, device="cuda"); d_col = torch.empty(batch, C, dtype=torch.int32, device="cuda")
        pb.check(pb.lib().pb_auction_solve(d_cost.data_ptr(), batch, R, C, d_row.data_ptr(), d_col.data_ptr(), d_act.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream))
        row, col = d_row.cpu().numpy(), d_col.cpu().numpy()
        for b in range(batch):
            r, c = orc.auction(cost[b], act[b])
            assert np.array_equal(row[b], r) and np.array_equal(col[b], c), (R, C, b)
        pb.check(pb.lib().pb_auction_solve(d_cost.data_ptr(), batch, R, C, d_row.data_ptr(), d_col.data_ptr(), None,
                                           torch.cuda.current_stream().cuda_stream))
        r, c = orc.auction(cost[0])
        assert np.array_equal(d_row[0].cpu().numpy(), r)


def _tracker_like_problem(rng, R, C, na, kind):
    """Cost tables shaped like the tracker's: few active rows, one cheap cell per row, everything else a tie
    at exactly 1.0 or locked at 1e9 (the states in which the auction runs to its iteration limit)."""
    cost = np.full((R, C), 1.0, np.float32)
    act = np.zeros(R, np.int32)
    rows = np.sort(rng.choice(R, na, replace=False))
    act[rows] = 1
    for i, r in enumerate(rows):
        if kind == "locked":
            cost[r] = 1e9
            for c in rng.choice(C, min(C, 3), replace=False):
                cost[r, c] = np.float32(rng.uniform(0.6, 1.0))
        elif kind == "ties":
            cost[r, rng.uniform(size=C) < 0.3] = np.float32(0.5)
        elif kind == "zero":
            cost[r, rng.uniform(size=C) < 0.2] = np.float32(0.0)
            cost[r, rng.uniform(size=C) < 0.1] = np.float32(-0.0)
        if i < C and kind != "ties":
            cost[r, i % C] = np.float32(rng.uniform(0.02, 0.4))
    if kind == "nan":
        cost[rows[0], :] = np.nan
        cost[rows[-1], C // 2] = np.nan
    return cost, act


def test_auction_few_active_rows(pb, orc, cuda):
    """At most 32 active rows: pb_auction_solve takes the single-warp hybrid solve (lane = column for up to four
    bidders, lane = row above); assignments must equal the checker's on over-subscribed, tied and locked tables."""
    torch = cuda
    rng = np.random.default_rng(17)
    st = torch.cuda.current_stream().cuda_stream
    shapes = [(128, 20, 21), (128, 20, 20), (128, 64, 32), (128, 33, 30), (64, 1, 5), (40, 100, 32), (32, 32, 32), (128, 19, 24),
              (16, 300, 16), (128, 5, 1)]
    for R, C, na in shapes:
        for kind in ("plain", "locked", "ties", "zero", "nan"):
            batch = 6
            costs, acts = zip(*[_tracker_like_problem(rng, R, C, na, kind) for _ in range(batch)])
            cost, act = np.stack(costs), np.stack(acts)
            d_cost, d_act = torch.from_numpy(cost).cuda(), torch.from_numpy(act).cuda()
            d_row = torch.empty(batch, R, dtype=torch.int32, device="cuda"); d_col = torch.empty(batch, C, dtype=torch.int32, device="cuda")
            pb.check(pb.lib().pb_auction_solve(d_cost.data_ptr(), batch, R, C, d_row.data_ptr(), d_col.data_ptr(), d_act.data_ptr(), st))
            row, col = d_row.cpu().numpy(), d_col.cpu().numpy()
            for b in range(batch):
                r, c = orc.auction(cost[b], act[b])
                assert np.array_equal(row[b], r) and np.array_equal(col[b], c), (R, C, na, kind, b)


def test_kf3(pb, orc, cuda):
    torch = cuda
    L = pb.lib()
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.default_rng(8)
    T = 50
    dets = rng.uniform(0, 1080, (T, 17, 3)).astype(np.float32); dets[..., 2] = rng.uniform(0, 1, (T, 17)); dets[3, 5, 2] = 0
    slots = rng.permutation(T).astype(np.int32)
    k = orc.KF3(T)
    means = torch.zeros(T, 136, device="cuda"); diag = torch.zeros(T, 136, device="cuda")
    d_dets = torch.from_numpy(dets.reshape(T, 51)).cuda(); d_slots = torch.from_numpy(slots).cuda()
    pb.check(L.pb_kf3_initiate(means.data_ptr(), diag.data_ptr(), d_dets.data_ptr(), d_slots.data_ptr(), T, st))
    k.initiate(dets.reshape(T, 51), slots)
    for it in range(4):
        pb.check(L.pb_kf3_predict(means.data_ptr(), diag.data_ptr(), T, 0.9, 0.9, st)); k.predict(T)
        z = (dets + rng.normal(0, 4, dets.shape)).astype(np.float32); z[..., 2] = rng.uniform(0, 1, (T, 17))
        m = np.stack([rng.permutation(T)[:30], rng.permutation(T)[:30]], 1).astype(np.int32)
        d_z, d_m = torch.from_numpy(z.reshape(T, 51)).cuda(), torch.from_numpy(m).cuda()
        pb.check(L.pb_kf3_update(means.data_ptr(), diag.data_ptr(), d_z.data_ptr(), d_m.data_ptr(), 30, st)); k.update(z.reshape(T, 51), m)
        rm, rd = k.state()
        assert means.cpu().numpy().tobytes() == rm.tobytes() and diag.cpu().numpy().tobytes() == rd.tobytes(), it
    out = torch.zeros(5, 51, device="cuda"); sl = torch.tensor([4, 0, 9, 49, 17], dtype=torch.int32, device="cuda")
    pb.check(L.pb_kf3_extract(means.data_ptr(), out.data_ptr(), sl.data_ptr(), 5, st))
    assert out.cpu().numpy().tobytes() == k.extract(sl.cpu().numpy()).tobytes()
    cov = torch.empty(136, 136, device="cuda")
    pb.check(L.pb_kf3_materialize_cov(diag.data_ptr(), 7, cov.data_ptr(), st))
    assert cov.cpu().numpy().tobytes() == k.full_state(7)[1].tobytes()


def _decoded(pb, orc, canvas, persons, clumps, stream):
    cfg = pb.synth_config(canvas=canvas, persons=persons, period=8, clumps=clumps)
    return orc.decode(pb.synth_heads(cfg, stream, 1, 2, 1)[0, 0], 0.30)


def test_launch_pose_nms(pb, orc, cuda):
    torch = cuda
    for canvas, persons, clumps in [(640, 20, 0), (1280, 80, 8)]:
        d = _decoded(pb, orc, canvas, persons, clumps, 1)
        n = d["num"]
        poses, scores = torch.from_numpy(d["poses"].copy()).cuda(), torch.from_numpy(d["scores"].copy()).cuda()
        sig = torch.from_numpy(SIG).cuda(); keep = torch.full((n,), -1, dtype=torch.int32, device="cuda")
        for oks_thr, sc_thr in [(0.65, 0.25), (0.3, 0.5)]:
            pb.lib().launchPoseNMS(poses.data_ptr(), scores.data_ptr(), sig.data_ptr(), keep.data_ptr(), n, 17, oks_thr, sc_thr,
                                   torch.cuda.current_stream().cuda_stream)
            ref = orc.pose_nms(d["poses"], d["scores"], SIG, oks_thr, sc_thr)
            assert np.array_equal(keep.cpu().numpy(), ref) and 0 < ref.sum() < n


def test_nms_legacy_batch(pb, orc, cuda):
    torch = cuda
    imgs = [_decoded(pb, orc, 640, 20, 0, s) for s in range(3)] + [_decoded(pb, orc, 1280, 80, 8, 9)]
    recs, offs = [], [0]
    for d in imgs:
        r = np.zeros(d["num"], pb.POSE_DETECTION)
        r["bbox"] = d["bboxes"]; r["score"] = d["scores"]; r["keypoints"] = d["poses"].reshape(-1, 17, 3)
        recs.append(r); offs.append(offs[-1] + d["num"])
    allr = np.concatenate(recs)
    d_dets = torch.from_numpy(allr.view(np.uint8).copy()).cuda()
    d_off = torch.tensor(offs, dtype=torch.int32, device="cuda")
    d_keep = torch.full((len(allr),), -1, dtype=torch.int32, device="cuda"); d_nk = torch.zeros(len(imgs), dtype=torch.int32, device="cuda")
    pb.check(pb.lib().pb_nms_legacy(d_dets.data_ptr(), d_off.data_ptr(), len(imgs), 1024, 0.65, 0.25, d_keep.data_ptr(), d_nk.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream))
    keep, nk = d_keep.cpu().numpy(), d_nk.cpu().numpy()
    for i, r in enumerate(recs):
        ref = orc.nms_legacy(r, 0.65, 0.25)
        assert nk[i] == len(ref) and np.array_equal(keep[offs[i]: offs[i] + nk[i]], ref), i
