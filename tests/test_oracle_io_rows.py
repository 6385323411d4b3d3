"""CPU checker for the rows either side of the path: letterbox pre-processing
(preprocess.cu:19-153) and LinearAssignmentCUDA::solve (hungarian.cu:235-339).  Properties here; the
comparison with the reference's own code is the golden fixture tests/golden/ref_io_rows.npz (recorded
from the reference on a B200 by tools/make_golden_io.py) plus the live cross-check in
tests/test_gpu_io_rows.py."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "ref_io_rows.npz")


def make_frame(rng, w, h):
    """Smooth gradients + noise: interpolation errors would show, values cover 0..255."""
    yy, xx = np.mgrid[0:h, 0:w]
    img = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) % 256)], -1).astype(np.int32)
    img = (img + rng.integers(-20, 21, img.shape)).clip(0, 255)
    return img.astype(np.uint8)


def test_letterbox_geometry_and_padding(orc):
    rng = np.random.default_rng(1)
    for (w, h, tw, th) in [(1920, 1080, 640, 640), (1080, 1920, 640, 640), (640, 480, 640, 640), (333, 517, 320, 256), (64, 64, 640, 640)]:
        img = make_frame(rng, w, h)
        out, xf = orc.letterbox(img, tw, th)
        scale = min(np.float32(tw) / np.float32(w), np.float32(th) / np.float32(h))
        nw, nh = int(np.float32(w) * scale), int(np.float32(h) * scale)
        px, py = (tw - nw) // 2, (th - nh) // 2
        assert xf[2] == px and xf[3] == py and xf[0] == xf[1] == np.float32(1.0) / scale
        gray = np.float32(114.0) / np.float32(255.0)
        mask = np.ones((th, tw), bool); mask[py:py + nh, px:px + nw] = False
        assert (out[:, mask] == gray).all()
        inner = out[:, ~mask]
        assert inner.min() >= 0.0 and inner.max() <= 1.0 + 1e-6       # bilinear weights sum to 1 only up to rounding


def test_letterbox_same_size_is_channel_swap_over_255(orc):
    """scale == 1: every tap but the clamped last row/column lands on a pixel centre."""
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (96, 128, 3), dtype=np.uint8)
    out, xf = orc.letterbox(img, 128, 96)
    assert list(xf) == [1.0, 1.0, 0.0, 0.0]
    want = (img[:-1, :-1, ::-1].astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1)
    assert np.array_equal(out[:, :-1, :-1], want)


def test_letterbox_constant_image_stays_constant(orc):
    img = np.full((200, 300, 3), 77, np.uint8)
    out, xf = orc.letterbox(img, 640, 640)
    py = int(xf[3])
    inner = out[:, py:640 - py - 1, :]
    assert np.abs(inner - np.float32(77) / np.float32(255)).max() < 1e-6


def test_assign_legacy_greedy_below_100_cells(orc):
    cost = np.array([[0.4, 0.1, 0.9], [0.2, 0.1, 0.3], [0.6, 0.7, 0.8]], np.float32)
    row, col, n = orc.assign_legacy(cost, 0.5)
    assert list(row) == [1, 0, -1] and list(col) == [1, 0, -1] and n == 2       # row 2 has nothing below 0.5
    cost = np.full((3, 4), 0.25, np.float32)                                    # ties: lowest free column
    row, col, n = orc.assign_legacy(cost, 0.5)
    assert list(row) == [0, 1, 2] and n == 3


def test_assign_legacy_auction_and_threshold_filter(orc):
    rng = np.random.default_rng(5)
    for R, C in [(10, 10), (50, 50), (30, 12), (12, 30), (128, 64)]:
        cost = rng.uniform(0, 1, (R, C)).astype(np.float32)
        row, col, n = orc.assign_legacy(cost, 0.5)
        assert n == (row >= 0).sum() == (col >= 0).sum()
        for r, c in enumerate(row):
            if c >= 0:
                assert col[c] == r and cost[r, c] <= 0.5
        row1, col1, n1 = orc.assign_legacy(cost, 1.0)                           # nothing filtered
        assert n1 >= n and all(row[r] == row1[r] for r in range(R) if row[r] >= 0)
    # at most 50 iterations decide the tracker's solve; the legacy one runs to 3*rows
    cost = rng.uniform(0, 1, (40, 40)).astype(np.float32)
    r50, _ = orc.auction(cost)
    r120, _, _ = orc.assign_legacy(cost, 1e9)
    assert (r120 >= 0).sum() >= (r50 >= 0).sum()


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden fixture not recorded yet")
def test_checker_equals_reference_goldens(orc):
    g = np.load(GOLD)
    for i in range(int(g["n_frames"])):
        img = g[f"img{i}"]
        tw, th = (int(v) for v in g[f"target{i}"])
        out, xf = orc.letterbox(img, tw, th)
        ref = g[f"out{i}"]
        assert np.array_equal(xf, g[f"xf{i}"]), i
        assert np.abs(out - ref).max() <= 1e-4 * max(1.0, float(np.abs(ref).max())), (i, np.abs(out - ref).max())
    for i in range(int(g["n_costs"])):
        cost, thr = g[f"cost{i}"], float(g[f"thr{i}"])
        row, col, n = orc.assign_legacy(cost, thr)
        assert np.array_equal(row, g[f"row{i}"]) and np.array_equal(col, g[f"col{i}"]) and n == int(g[f"cnt{i}"]), i
