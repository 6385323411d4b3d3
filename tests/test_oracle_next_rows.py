"""SURVEY.md §8f row f1 on the CPU: the checker's restatement of OKSDistanceCUDA and of
GreedyMatcherCUDA's host rule against independent numpy / Python formulations."""
import numpy as np

SIG = np.array([0.026, 0.025, 0.025, 0.035, 0.035, 0.079, 0.079, 0.072, 0.072, 0.062, 0.062, 0.107, 0.107, 0.087, 0.087, 0.089, 0.089])


def rand_poses(rng, n, spread=400.0, drop=0.2):
    p = np.zeros((n, 17, 3), np.float32)
    c = rng.uniform(100, 100 + spread, (n, 1, 2))
    p[:, :, :2] = c + rng.normal(0, 25, (n, 17, 2))
    p[:, :, 2] = rng.uniform(0.0, 1.0, (n, 17))
    p[rng.uniform(0, 1, (n, 17)) < drop, 2] = 0.0
    return p


def np_oks_cost(t, d):
    def area(p, thr):
        m = p[:, 2] > thr
        if not m.any():
            return (-1e9 - 1e9) * (-1e9 - 1e9), 0
        return (p[m, 0].max() - p[m, 0].min()) * (p[m, 1].max() - p[m, 1].min()), int(m.sum())
    da, dv = area(d.astype(np.float64), 0.1); ta, _ = area(t.astype(np.float64), 0.1)
    scale = max((da + ta) * 0.5, 1000.0)
    if dv < 2:
        return 1.0
    def oks(thr):
        m = (d[:, 2] > thr) & (t[:, 2] > thr)
        if not m.any():
            return 0.0, 0
        d2 = ((d[m, :2].astype(np.float64) - t[m, :2]) ** 2).sum(1)
        return float(np.exp(-d2 / (2 * scale * (2 * SIG[m]) ** 2)).sum()), int(m.sum())
    s, n = oks(0.2)
    if n >= 3:
        return 1.0 - s / n
    s, n = oks(0.05)
    return 1.0 - (s / n if n > 0 else 0.0)


def np_iou_cost(t, d):
    def box(p):
        m = p[:, 2] > 0
        return np.array([p[m, 0].min() - 10, p[m, 1].min() - 10, p[m, 0].max() + 10, p[m, 1].max() + 10], np.float64)
    a, b = box(t), box(d)
    iw = max(0.0, min(a[2], b[2]) - max(a[0], b[0])); ih = max(0.0, min(a[3], b[3]) - max(a[1], b[1]))
    inter = iw * ih
    uni = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    return 1.0 - (inter / uni if uni > 0 else 0.0)


def test_pose_distance_against_numpy(orc):
    rng = np.random.default_rng(5)
    t, d = rand_poses(rng, 9), rand_poses(rng, 13)
    d[:4, :, :2] = t[:4, :, :2] + rng.normal(0, 3, (4, 17, 2)).astype(np.float32)      # near matches
    d[5, :, 2] = 0.0; d[5, 0, 2] = 0.5                                                   # one usable keypoint: degenerate, cost 1
    d[6, :, 2] = np.where(np.arange(17) < 2, 0.9, 0.15)                                  # fallback threshold path
    o, u, c = orc.pose_distance(t, d, 0), orc.pose_distance(t, d, 1), orc.pose_distance(t, d, 2, 0.6)
    for i in range(9):
        for j in range(13):
            assert abs(o[i, j] - np_oks_cost(t[i], d[j])) < 2e-5, (i, j)
            assert abs(u[i, j] - np_iou_cost(t[i], d[j])) < 2e-5, (i, j)
    assert (o[:, 5] == 1.0).all()
    assert np.allclose(c, np.float32(0.6) * o + (np.float32(1) - np.float32(0.6)) * u, rtol=0, atol=1e-6)
    assert o[:4, :4].diagonal().max() < 0.5 < o[0, 8]


def test_greedy_match_rule(orc):
    rng = np.random.default_rng(2)
    for R, C in [(5, 5), (12, 7), (7, 12), (40, 33)]:
        cost = rng.uniform(0, 1, (R, C)).astype(np.float32)
        cost[rng.uniform(0, 1, (R, C)) < 0.2] = 0.25            # exact ties: (cost, row, col) order decides
        got = orc.greedy_match(cost, 0.6)
        cells = sorted((float(cost[r, c]), r, c) for r in range(R) for c in range(C) if cost[r, c] < 0.6)
        want = np.full(R, -1, np.int32); cu = set()
        for v, r, c in cells:
            if want[r] < 0 and c not in cu:
                want[r] = c; cu.add(c)
        assert np.array_equal(got, want)
        assert len(set(got[got >= 0].tolist())) == (got >= 0).sum()
    assert (orc.greedy_match(np.ones((3, 4), np.float32), 0.5) == -1).all()
