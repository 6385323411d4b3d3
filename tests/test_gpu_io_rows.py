"""GPU: batched letterbox pre-processing (pb_letterbox_batch) and the legacy assignment solve
(pb_assign_legacy) against the CPU checker bit for bit, against the golden fixture recorded from the
reference, and — where oracle/_ref is present — against the reference running live on this box."""
import os

import numpy as np
import pytest

from test_oracle_io_rows import GOLD, make_frame

pytestmark = pytest.mark.gpu


def _pack(torch, frames):
    stride = max(f.size for f in frames)
    stride = (stride + 255) // 256 * 256
    buf = np.zeros((len(frames), stride), np.uint8)
    for i, f in enumerate(frames):
        buf[i, : f.size] = f.reshape(-1)
    sizes = np.array([[f.shape[1], f.shape[0]] for f in frames], np.int32)
    return torch.from_numpy(buf).cuda(), torch.from_numpy(sizes).cuda()


def test_letterbox_batch_equals_checker(pb, orc, cuda):
    torch = cuda
    rng = np.random.default_rng(3)
    shapes = [(1920, 1080), (1080, 1920), (640, 480), (333, 517), (64, 64), (641, 359), (2, 2), (1, 7)]
    frames = [make_frame(rng, w, h) for (w, h) in shapes]
    for tw, th in [(640, 640), (1280, 1280), (322, 250), (33, 17)]:          # the last two: widths not divisible by 4
        d_frames, d_sizes = _pack(torch, frames)
        out, xf = pb.letterbox_batch(d_frames, d_sizes, tw, th)
        out, xf = out.cpu().numpy(), xf.cpu().numpy()
        for i, f in enumerate(frames):
            want, wxf = orc.letterbox(f, tw, th)
            assert np.array_equal(xf[i], wxf), (tw, th, i, xf[i], wxf)
            assert out[i].tobytes() == want.tobytes(), (tw, th, shapes[i], np.abs(out[i] - want).max())


def test_letterbox_feeds_output_transform(pb, orc, cuda):
    """The xform rows are what pb_set_output_transform takes: a point of the source frame mapped into
    the letterboxed frame comes back under (v - pad) * scale (main.cpp:48-68)."""
    torch = cuda
    rng = np.random.default_rng(5)
    frames = [make_frame(rng, 1920, 1080), make_frame(rng, 720, 1280)]
    d_frames, d_sizes = _pack(torch, frames)
    _, xf = pb.letterbox_batch(d_frames, d_sizes, 640, 640)
    xf = xf.cpu().numpy()
    for (w, h), t in zip([(1920, 1080), (720, 1280)], xf):
        sx, sy, px, py = t
        x_src, y_src = 0.37 * w, 0.61 * h
        x_lb, y_lb = x_src / sx + px, y_src / sy + py
        assert abs((x_lb - px) * sx - x_src) < 1e-2 and abs((y_lb - py) * sy - y_src) < 1e-2
        assert 0 <= x_lb < 640 and 0 <= y_lb < 640


def test_assign_legacy_equals_checker(pb, orc, cuda):
    torch = cuda
    rng = np.random.default_rng(6)
    for R, C in [(3, 3), (9, 11), (1, 99), (10, 10), (50, 50), (30, 12), (12, 30), (128, 64), (200, 150)]:
        batch = 4
        cost = rng.uniform(0, 1, (batch, R, C)).astype(np.float32)
        cost[rng.uniform(size=cost.shape) < 0.05] = 0.25                        # ties
        for thr in (0.5, 1.0, 0.05):
            row, col, cnt = pb.assign_legacy(torch.from_numpy(cost).cuda(), thr)
            row, col, cnt = row.cpu().numpy(), col.cpu().numpy(), cnt.cpu().numpy()
            for b in range(batch):
                wr, wc, wn = orc.assign_legacy(cost[b], thr)
                assert np.array_equal(row[b], wr) and np.array_equal(col[b], wc) and cnt[b] == wn, (R, C, thr, b)


@pytest.mark.skipif(not os.path.exists(GOLD), reason="golden fixture not recorded yet")
def test_cuda_path_equals_reference_goldens(pb, cuda):
    torch = cuda
    g = np.load(GOLD)
    for i in range(int(g["n_frames"])):
        img = g[f"img{i}"]
        tw, th = (int(v) for v in g[f"target{i}"])
        d_frames, d_sizes = _pack(torch, [img])
        out, xf = pb.letterbox_batch(d_frames, d_sizes, tw, th)
        ref = g[f"out{i}"]
        assert np.array_equal(xf[0].cpu().numpy(), g[f"xf{i}"])
        err = np.abs(out[0].cpu().numpy() - ref).max()
        assert err <= 1e-4 * max(1.0, float(np.abs(ref).max())), (i, err)       # fp32, reference contracts FMAs
    for i in range(int(g["n_costs"])):
        cost, thr = g[f"cost{i}"], float(g[f"thr{i}"])
        row, col, cnt = pb.assign_legacy(torch.from_numpy(cost[None]).cuda(), thr)
        assert np.array_equal(row[0].cpu().numpy(), g[f"row{i}"]) and np.array_equal(col[0].cpu().numpy(), g[f"col{i}"])
        assert int(cnt[0]) == int(g[f"cnt{i}"])


def test_checker_equals_reference_live(orc, cuda):
    import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(12)
    for (w, h, tw, th) in [(1280, 720, 640, 640), (500, 900, 640, 640), (640, 640, 1280, 1280)]:
        img = make_frame(rng, w, h)
        with ref_py.quiet():
            ref, rxf = ref_py.preprocess(img, tw, th)
        want, wxf = orc.letterbox(img, tw, th)
        assert np.array_equal(rxf, wxf)
        assert np.abs(ref - want).max() <= 1e-4
    for R, C in [(4, 5), (10, 10), (40, 40), (64, 20), (20, 64)]:
        cost = rng.uniform(0, 1, (R, C)).astype(np.float32)
        for thr in (0.5, 1.0):
            with ref_py.quiet():
                rr, rc, rn = ref_py.assign_solve(cost, thr)
            wr, wc, wn = orc.assign_legacy(cost, thr)
            assert np.array_equal(rr, wr) and np.array_equal(rc, wc) and rn == wn, (R, C, thr)
