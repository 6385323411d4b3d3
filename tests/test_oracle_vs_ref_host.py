"""Host-legacy NMS: the checker's restatement against NMSCuda::apply COMPILED FROM THE REFERENCE
(reference src/cuda/nms.cu:142-306 is host code; oracle/_ref/libposebyte_ref.so is built from the
reference's own sources by `make -C oracle ref` and runs without a GPU).  Skipped only where the
reference tree was never available to build it."""
import numpy as np
import pytest

import ref_py

pytestmark = pytest.mark.skipif(not ref_py.available(), reason="oracle/_ref not built (needs /root/reference at build time)")


def make_dets(rng, persons, dup, canvas=640.0, jitter=3.0, drop=0.1):
    tmpl = np.array([[0.0, -1.5], [-0.1, -1.6], [0.1, -1.6], [-0.2, -1.5], [0.2, -1.5], [-0.5, -1.0], [0.5, -1.0],
                     [-0.8, -0.3], [0.8, -0.3], [-1.0, 0.3], [1.0, 0.3], [-0.3, 0.0], [0.3, 0.0], [-0.3, 0.8],
                     [0.3, 0.8], [-0.3, 1.5], [0.3, 1.5]], np.float32)
    out = []
    for _ in range(persons):
        c = rng.uniform(0.1 * canvas, 0.9 * canvas, 2); s = rng.uniform(20, 60)
        for _ in range(rng.integers(1, dup + 1)):
            d = np.zeros((), ref_py.POSE_DETECTION)
            kp = np.zeros((17, 3), np.float32)
            kp[:, :2] = c + tmpl * s + rng.normal(0, jitter, (17, 2))
            kp[:, 2] = rng.uniform(0.25, 1.0, 17)
            kp[rng.uniform(0, 1, 17) < drop, 2] = rng.uniform(0, 0.15)
            lo, hi = kp[:, :2].min(0), kp[:, :2].max(0)
            d["bbox"] = [lo[0] - 3, lo[1] - 3, hi[0] + 3, hi[1] + 3]
            d["keypoints"] = kp
            d["score"] = rng.uniform(0.05, 0.99)
            out.append(d)
    return np.array(out, ref_py.POSE_DETECTION)


@pytest.mark.parametrize("seed,persons,dup,jitter", [(0, 10, 4, 3.0), (1, 30, 6, 1.0), (2, 60, 3, 8.0), (3, 5, 12, 0.5),
                                                     (4, 100, 5, 4.0), (5, 1, 1, 0.0)])
def test_nms_legacy_equals_reference_host_code(orc, seed, persons, dup, jitter):
    rng = np.random.default_rng(seed)
    dets = make_dets(rng, persons, dup, jitter=jitter)
    for score_thr in (0.25, 0.5):
        ref = ref_py.nms_apply(dets, 0.65, score_thr)
        got = orc.nms_legacy(dets, 0.65, score_thr)
        assert np.array_equal(ref, got), (seed, score_thr, ref[:10], got[:10])
    assert len(ref) < len(dets)          # something was suppressed or filtered


def test_nms_legacy_crowd_with_overlap(orc):
    rng = np.random.default_rng(11)
    dets = make_dets(rng, 40, 6, canvas=200.0, jitter=5.0)       # heavy overlap: all four rules fire
    ref = ref_py.nms_apply(dets, 0.65, 0.25)
    got = orc.nms_legacy(dets, 0.65, 0.25)
    assert np.array_equal(ref, got)


def test_nms_legacy_empty_and_all_below_threshold(orc):
    dets = make_dets(np.random.default_rng(3), 4, 2)
    dets["score"] = 0.1
    assert len(ref_py.nms_apply(dets, 0.65, 0.25)) == 0 == len(orc.nms_legacy(dets, 0.65, 0.25))
    assert len(orc.nms_legacy(dets[:0], 0.65, 0.25)) == 0
