"""The CPU checker's decode + native NMS (A1-A3) against an independent numpy restatement
and against its own literal (full-mask) formulation."""
import numpy as np
import pytest

SIG = np.array([0.026, 0.025, 0.025, 0.035, 0.035, 0.079, 0.079, 0.072, 0.072, 0.062, 0.062, 0.107, 0.107,
                0.087, 0.087, 0.089, 0.089], np.float32)


def np_decode(raw, thr, cap=1024):
    """gpu_postprocess.cu:30-81 in numpy; R1 ordering."""
    keep = np.nonzero(~(raw[4] < np.float32(thr)))[0][:cap]
    cx, cy, w, h = raw[0, keep], raw[1, keep], raw[2, keep], raw[3, keep]
    half = np.float32(0.5)
    bboxes = np.stack([cx - w * half, cy - h * half, cx + w * half, cy + h * half], 1)
    poses = raw[5:, keep].T.copy()
    return keep.astype(np.int32), poses, bboxes, raw[4, keep]


def literal_sweep(scores, mask, max_keep=256):
    """kernelSortByScore + kernelApplyNMSMask (gpu_postprocess.cu:178-242) on a full mask."""
    order = sorted(range(len(scores)), key=lambda i: (-scores[i], i))
    sup = np.zeros(len(scores), bool)
    keep = []
    for i in order:
        if len(keep) >= max_keep:
            break
        if sup[i]:
            continue
        keep.append(i)
        sup |= mask[i].astype(bool)
    return np.array(keep, np.int32)


@pytest.mark.parametrize("stream,frame", [(0, 0), (1, 7), (3, 19)])
def test_decode_matches_numpy(pb, orc, stream, frame):
    cfg = pb.synth_config(canvas=640, persons=20, period=64)
    raw = pb.synth_heads(cfg, stream, 1, frame, 1)[0, 0]
    a, poses, bboxes, scores = np_decode(raw, 0.30)
    d = orc.decode(raw, 0.30)
    assert d["num"] == len(a) and 60 <= d["num"] <= 1024
    assert np.array_equal(d["anchors"], a)
    assert d["poses"].tobytes() == poses.astype(np.float32).tobytes()
    assert d["bboxes"].tobytes() == bboxes.astype(np.float32).tobytes()
    assert d["scores"].tobytes() == scores.tobytes()


def test_decode_overflow_keeps_first_in_anchor_order(orc):
    rng = np.random.default_rng(0)
    raw = rng.uniform(0, 640, (56, 8400)).astype(np.float32)
    raw[4] = rng.uniform(0.25, 0.35, 8400)        # ~half above 0.30 -> far more than 1024
    d = orc.decode(raw, 0.30, max_cand=1024)
    expect = np.nonzero(~(raw[4] < np.float32(0.30)))[0][:1024]
    assert d["num"] == 1024 and np.array_equal(d["anchors"], expect)


def test_decode_empty_and_nan(orc):
    raw = np.zeros((56, 8400), np.float32)
    assert orc.decode(raw, 0.30)["num"] == 0
    raw[4, 17] = np.nan                            # NaN is not < thr: the reference keeps it
    assert list(orc.decode(raw, 0.30)["anchors"]) == [17]


@pytest.mark.parametrize("canvas,persons,clumps", [(640, 20, 0), (1280, 60, 6)])
def test_lazy_sweep_equals_literal_full_mask(pb, orc, canvas, persons, clumps):
    cfg = pb.synth_config(canvas=canvas, persons=persons, period=32, clumps=clumps)
    raw = pb.synth_heads(cfg, 5, 1, 3, 1)[0, 0]
    d = orc.decode(raw, 0.30)
    mask = orc.nms_mask(d["poses"], d["bboxes"], 0.65)
    assert np.array_equal(mask, mask.T), "overlap relation must be symmetric bit for bit"
    lit = literal_sweep(d["scores"], mask)
    r = orc.postprocess(raw, 0.30, 0.65)
    assert np.array_equal(r["keep_slots"], lit)
    assert np.array_equal(r["keep_anchors"], d["anchors"][lit])
    assert r["poses"].tobytes() == d["poses"][lit].tobytes()
    assert (np.diff(r["scores"]) <= 0).all()
    assert persons * 0.5 <= r["num_keep"] <= persons


def test_nms_pair_rule_against_numpy(pb, orc):
    """Overlap bits from an independent float32 numpy evaluation of gpu_postprocess.cu:113-168
    (np.exp instead of pb_expf): identical except for pairs within 1e-5 of a threshold."""
    cfg = pb.synth_config(canvas=640, persons=20, period=32)
    raw = pb.synth_heads(cfg, 2, 1, 9, 1)[0, 0]
    d = orc.decode(raw, 0.30)
    P, Bx = d["poses"].reshape(-1, 17, 3), d["bboxes"]
    n = len(P)
    mask = orc.nms_mask(d["poses"], d["bboxes"], 0.65)
    thr = np.float32(0.65)
    near = 0
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            ix1, iy1 = max(Bx[i, 0], Bx[j, 0]), max(Bx[i, 1], Bx[j, 1])
            ix2, iy2 = min(Bx[i, 2], Bx[j, 2]), min(Bx[i, 3], Bx[j, 3])
            inter = np.float32(max(np.float32(0), ix2 - ix1)) * np.float32(max(np.float32(0), iy2 - iy1))
            ai = (Bx[i, 2] - Bx[i, 0]) * (Bx[i, 3] - Bx[i, 1]); aj = (Bx[j, 2] - Bx[j, 0]) * (Bx[j, 3] - Bx[j, 1])
            uni = ai + aj - inter
            iou = inter / uni if uni > 0 else np.float32(0)
            bit = iou > thr
            margin = abs(float(iou) - 0.65)
            if not bit:
                vis = (P[i, :, 2] > 0.2) & (P[j, :, 2] > 0.2)
                if vis.sum() >= 3:
                    s2 = max(ai, aj, np.float32(1024.0))
                    d2 = ((P[i, vis, :2] - P[j, vis, :2]) ** 2).sum(1)
                    oks = float(np.exp(-d2.astype(np.float64) / (2.0 * float(s2) * 4.0 * SIG[vis].astype(np.float64) ** 2)).mean())
                    bit = oks > 0.65 or (oks > 0.4 and iou > 0.2)
                    margin = min(margin, abs(oks - 0.65), abs(oks - 0.4), abs(float(iou) - 0.2))
            if margin < 1e-5:
                near += 1
                continue
            assert bool(mask[i, j]) == bool(bit), (i, j)
    assert near < 5


def test_keep_cap_256(orc):
    """The sweep stops at 256 kept detections (gpu_postprocess.cu:224)."""
    rng = np.random.default_rng(3)
    n = 600
    raw = np.zeros((56, 8400), np.float32)
    idx = rng.choice(8400, n, replace=False)
    raw[4, idx] = rng.uniform(0.4, 0.9, n)
    raw[0, idx] = rng.uniform(0, 100000, n); raw[1, idx] = rng.uniform(0, 100000, n)   # far apart
    raw[2, idx] = 10; raw[3, idx] = 10
    r = orc.postprocess(raw, 0.30, 0.65)
    assert r["num_cand"] == n and r["num_keep"] == 256
    top = np.sort(raw[4, idx])[::-1][:256]
    assert np.array_equal(r["scores"], top)


def test_score_ties_are_stable(orc):
    raw = np.zeros((56, 8400), np.float32)
    for k, a in enumerate([100, 50, 7000, 300]):
        raw[4, a] = 0.5
        raw[0, a] = 1000.0 * (k + 1); raw[1, a] = 50; raw[2, a] = 20; raw[3, a] = 20
    r = orc.postprocess(raw, 0.30, 0.65)
    assert list(r["keep_anchors"]) == [50, 100, 300, 7000]     # equal scores: lower anchor first (R2)
