"""GPU decode + NMS (pb_postprocess, through the C ABI) against the CPU checker: bit-exact
candidate counts, kept anchors / slots and kept poses, boxes, scores."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run_post(pb, torch, heads, conf=0.30, nms=0.65, **cfg):
    B, _, N = heads.shape
    pipe = pb.Pipeline(num_streams=B, num_anchors=N, **cfg)
    d = torch.from_numpy(np.ascontiguousarray(heads)).cuda()
    pipe.postprocess(d, conf, nms)
    torch.cuda.synchronize()
    return pipe, [pipe.get_kept(b) for b in range(B)]


def assert_same(got, ref, tag=""):
    assert got["num_cand"] == ref["num_cand"], tag
    assert np.array_equal(got["keep_anchors"], ref["keep_anchors"]), tag
    assert np.array_equal(got["keep_slots"], ref["keep_slots"]), tag
    for k in ("poses", "bboxes", "scores"):
        assert got[k].tobytes() == ref[k].tobytes(), (tag, k)


@pytest.mark.parametrize("canvas,persons,clumps,B,frames", [(640, 20, 0, 8, 3), (1280, 100, 10, 4, 2), (640, 3, 0, 3, 1)])
def test_postprocess_matches_checker(pb, orc, cuda, canvas, persons, clumps, B, frames):
    cfg = pb.synth_config(canvas=canvas, persons=persons, period=16, clumps=clumps, kp_drop_prob=0.15 if clumps else 0.05)
    heads = pb.synth_heads(cfg, 0, B, 0, frames, frame_major=True)
    for f in range(frames):
        for mode in (0, 1, 3):               # complete sweep, lazy keypoints, deferred OKS tests on complete records
            _, got = run_post(pb, cuda, heads[f], keypoint_fetch=mode)
            for b in range(B):
                assert_same(got[b], orc.postprocess(heads[f, b]), f"f{f} b{b} mode{mode}")


def test_nms_kernel_in_the_trackers_shared_memory_configuration(pb, orc, cuda, monkeypatch):
    """By default the NMS kernel is sized to the shared-memory configuration of the tracker kernel (163 KB beside a 137 KB
    tracker CTA instead of 224 KB for the full cap of 1024 candidates: SMs then keep their L1 / shared-memory split between the
    two).  Streams with more candidates than that working set holds take the spill path inside the same launch: a crowd of 130
    persons (about 900 candidates) beside one of 60 — same results as the checker, and as the plain kernel."""
    cfg_big = pb.synth_config(canvas=1280, persons=130, period=16, clumps=10, kp_drop_prob=0.15)
    cfg_small = pb.synth_config(canvas=1280, persons=60, period=16, clumps=6, kp_drop_prob=0.15)
    heads = np.concatenate([pb.synth_heads(cfg_big, 1, 1, 0, 1, frame_major=True)[0], pb.synth_heads(cfg_small, 2, 1, 0, 1, frame_major=True)[0]])
    pipe, got = run_post(pb, cuda, heads)
    plan = pipe.nms_plan()
    assert plan["threads"] == 1024 and plan["tier_candidates"] < 1024 and plan["smem_bytes"] <= 163 * 1024, plan
    assert got[0]["num_cand"] > plan["tier_candidates"] > got[1]["num_cand"], (got[0]["num_cand"], got[1]["num_cand"], plan)
    for b in range(2):
        assert_same(got[b], orc.postprocess(heads[b]), f"b{b}")
    monkeypatch.setenv("PB_PIPE_NMS_TIER", "0")
    pipe0, got0 = run_post(pb, cuda, heads)
    assert pipe0.nms_plan()["tier_candidates"] == 1024
    for b in range(2):
        assert_same(got0[b], got[b], f"plain kernel b{b}")


def test_empty_and_single(pb, orc, cuda):
    heads = np.zeros((3, 56, 8400), np.float32)
    heads[1, 4, 4242] = 0.9; heads[1, :4, 4242] = [100, 100, 40, 80]
    heads[2, 4, 8399] = 0.31; heads[2, 4, 0] = 0.31                      # first and last anchor, equal score
    _, got = run_post(pb, cuda, heads)
    assert got[0]["num_keep"] == 0 and got[0]["num_cand"] == 0
    for b in range(3):
        assert_same(got[b], orc.postprocess(heads[b]), f"b{b}")
    assert list(got[2]["keep_anchors"]) == [0, 8399]


def test_candidate_overflow_and_keep_cap(pb, orc, cuda):
    rng = np.random.default_rng(0)
    heads = rng.uniform(0, 640, (2, 56, 8400)).astype(np.float32)
    heads[0, 4] = rng.uniform(0.25, 0.35, 8400)                          # ~4000 above threshold -> first 1024
    n = 700
    heads[1, 4] = 0.0
    idx = rng.choice(8400, n, replace=False)
    heads[1, 4, idx] = rng.uniform(0.4, 0.9, n)
    heads[1, 0, idx] = rng.uniform(0, 1e5, n); heads[1, 1, idx] = rng.uniform(0, 1e5, n)
    heads[1, 2, idx] = 10; heads[1, 3, idx] = 10                         # disjoint boxes: nothing suppressed
    _, got = run_post(pb, cuda, heads)
    ref0, ref1 = orc.postprocess(heads[0]), orc.postprocess(heads[1])
    assert got[0]["num_cand"] == 1024 and got[1]["num_keep"] == 256
    assert_same(got[0], ref0, "overflow"); assert_same(got[1], ref1, "cap")


def test_threshold_sweep_and_small_capacities(pb, orc, cuda):
    cfg = pb.synth_config(canvas=640, persons=15, period=16)
    heads = pb.synth_heads(cfg, 7, 2, 4, 1)[0]
    for conf, nms in [(0.05, 0.65), (0.5, 0.3), (0.30, 0.95), (0.30, 0.05)]:
        cap = 1024 if conf > 0.1 else 1024
        _, got = run_post(pb, cuda, heads, conf, nms)
        for b in range(2):
            assert_same(got[b], orc.postprocess(heads[b], conf, nms), f"conf{conf} nms{nms} b{b}")
    _, got = run_post(pb, cuda, heads, max_candidates=64, max_keep=8)
    for b in range(2):
        assert_same(got[b], orc.postprocess(heads[b], max_cand=64, max_keep=8), f"small b{b}")


def test_unaligned_anchor_count(pb, orc, cuda):
    """N not a multiple of 4: the scalar scan path."""
    rng = np.random.default_rng(5)
    N = 8403
    heads = rng.uniform(0, 640, (2, 56, N)).astype(np.float32)
    heads[:, 4] = rng.uniform(0, 0.2, (2, N))
    hit = rng.choice(N, 90, replace=False)
    heads[:, 4, hit] = rng.uniform(0.35, 0.9, (2, 90))
    heads[0, 4, N - 1] = 0.8
    _, got = run_post(pb, cuda, heads)
    for b in range(2):
        assert_same(got[b], orc.postprocess(heads[b]), f"b{b}")


def test_lazy_sweep_rounds_match_checker(pb, orc, cuda):
    """Lazy sweep: per tile the NMS kernel first sweeps by IoU alone and fetches keypoints only for the
    speculative survivors; if a rank shadowed only by a survivor that fell to an OKS rule turns up, the
    tile is redone with the keypoints of all its live ranks.  Stream 0: ordinary duplicates
    (IoU decides: fast path).  Stream 1: two anchors with the same skeleton but boxes that overlap
    too little for the IoU rule — only the OKS rule removes the weaker one (complete path).
    Stream 2: same skeletons, IoU in (0.2, thr] and OKS in (0.4, thr]: the combined rule fires."""
    cfg = pb.synth_config(canvas=640, persons=6, period=16)
    base = pb.synth_heads(cfg, 9, 1, 0, 1, frame_major=True)[0, 0]
    heads = np.stack([base, base.copy(), base.copy()])
    tmpl = np.array([[0.0, -1.5], [-0.1, -1.6], [0.1, -1.6], [-0.2, -1.5], [0.2, -1.5], [-0.5, -1.0], [0.5, -1.0], [-0.8, -0.3],
                     [0.8, -0.3], [-1.0, 0.3], [1.0, 0.3], [-0.3, 0.0], [0.3, 0.0], [-0.3, 0.8], [0.3, 0.8], [-0.3, 1.5], [0.3, 1.5]], np.float32)

    def put(h, anchor, score, cx, cy, w, hh, kps, kshift=0.0):
        h[:4, anchor] = [cx, cy, w, hh]; h[4, anchor] = score
        for k in range(17):
            h[5 + 3 * k, anchor] = kps[k, 0] + kshift; h[6 + 3 * k, anchor] = kps[k, 1]; h[7 + 3 * k, anchor] = 0.9
    kps = np.array([320.0, 520.0], np.float32) + tmpl * 40.0
    # stream 1: identical skeletons, boxes 120x200 shifted by 70 px -> IoU = 50*200/(2*24000-10000) = 0.26 < 0.65, OKS = 1
    put(heads[1], 7000, 0.93, 320, 520, 120, 200, kps)
    put(heads[1], 7003, 0.91, 390, 520, 120, 200, kps)
    # stream 2: skeleton shifted by 14 px (OKS between 0.4 and 0.65 at this scale), boxes shifted 40 px (IoU = 0.5)
    put(heads[2], 7000, 0.93, 320, 520, 120, 200, kps)
    put(heads[2], 7003, 0.91, 360, 520, 120, 200, kps, kshift=14.0)
    # stream 0: A kept; B (same skeleton, IoU(A,B) = 0.26) falls to the OKS rule; C's box overlaps B's (IoU 0.85) but
    # not A's (0.2) and its skeleton is far from A's: C is shadowed only by B, so it must be KEPT — second round
    put(heads[0], 7000, 0.93, 320, 520, 120, 200, kps)
    put(heads[0], 7003, 0.91, 390, 520, 120, 200, kps)
    put(heads[0], 7010, 0.89, 400, 520, 120, 200, kps, kshift=200.0)
    refs = [orc.postprocess(heads[b]) for b in range(3)]
    for mode in (2, 3, 1):                   # complete sweep / deferred OKS tests / lazy keypoints: same results
        pipe, got = run_post(pb, cuda, heads, keypoint_fetch=mode)
        for b in range(3):
            assert_same(got[b], refs[b], f"mode {mode} stream {b}")
    assert 7000 in got[1]["keep_anchors"] and 7003 not in got[1]["keep_anchors"]      # removed by the OKS rule alone
    assert {7000, 7010} <= set(got[0]["keep_anchors"].tolist()) and 7003 not in got[0]["keep_anchors"]
    counts = pipe.nms_path_counts()          # the lazy run: 3 stream-frames
    assert counts["stream_frames"] == 3, counts
