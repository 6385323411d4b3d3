"""GPU cross-checks against the reference (needs a B200).

(1) The product's CUDA path against the committed golden vectors (outputs of the reference's own
    kernels, tests/golden/): kept detections bit-exact, TrackOutput records and final state equal
    up to a one-to-one renaming of ids, floats within 1e-4 relative.
(2) The reference run LIVE on this box (oracle/_ref/libposebyte_ref.so = its unmodified sources
    compiled for sm_100a; a built artefact, /root/reference is not read) on inputs that are NOT
    in the fixtures, against the CPU checker in replay mode: every discrete output equal frame
    by frame.  This is the same procedure that produced the fixtures (tools/make_golden.py)."""
import json
import os
import sys

import numpy as np
import pytest

import golden_util as gu

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tools"))
META = json.load(open(os.path.join(HERE, "golden", "ref_b200.json")))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("key", META["selfcheck_at_generation"]["rules_R3_R4_equal_up_to_id_renaming"])
def test_cuda_path_equals_reference_goldens(pb, cuda, key):
    torch = cuda
    name, s = key.split("/s"); s = int(s)
    sc = META["scenarios"][name]
    gold = gu.golden_stream(np.load(os.path.join(HERE, "golden", "ref_b200.npz")), name, s)
    scfg = pb.synth_config(**sc["synth"])
    heads = torch.from_numpy(pb.synth_heads(scfg, s, 1, 0, sc["frames"], frame_major=True)).cuda()   # [F,1,56,N]
    pipe = pb.Pipeline(num_streams=1, num_anchors=scfg.num_anchors, new_track_thresh=sc["conf"], high_thresh=sc["conf"],
                       low_thresh=sc["conf"] / 2, **sc["trk"])
    bad, ids = [], gu.IdMap()
    for f in range(sc["frames"]):
        pipe.step(heads[f], f, sc["conf"], sc["nms"])
        w = f"{name} stream {s} frame {f}"
        k = pipe.get_kept(0)
        n = int(gold["num_keep"][f])
        bad += gu.compare_kept(dict(num_keep=n, scores=gold["kept_scores"][f], poses=gold["kept_poses"][f],
                                    bboxes=gold["kept_bboxes"][f]), k, w)
        if int(gold["num_active"][f]) != int(pipe.get_num_active()[0]):
            bad.append(f"{w}: num_active {pipe.get_num_active()[0]} vs reference {gold['num_active'][f]}")
        bad += gu.compare_tracks(gold["tracks"][f][: int(gold["num_tracks"][f])], pipe.get_tracks(0), ids, w)
        assert len(bad) < 10, "\n".join(bad)
    ref_state = {k2[len("state_"):]: v for k2, v in gold.items() if k2.startswith("state_")}
    bad += gu.compare_final_state(ref_state, pipe.get_state(0), ids, f"{name} stream {s} final")
    assert not bad, "\n".join(bad[:10])


def _need_ref():
    import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref/libposebyte_ref.so was not built (needs /root/reference at build time)")
    return ref_py


@pytest.mark.parametrize("name,synth,streams,frames,trk", [
    ("live640", dict(canvas=640, persons=16, period=40), [11, 12], 30, dict(max_tracks=128, max_detections=64, max_age=10, min_hits=3)),
    ("liveoccl", dict(canvas=640, persons=8, period=60, occlusion=1), [21], 60, dict(max_tracks=64, max_detections=32, max_age=3, min_hits=2)),
])
def test_live_reference_equals_checker_in_replay_mode(pb, orc, cuda, name, synth, streams, frames, trk):
    ref = _need_ref()
    import make_golden as mg
    sc = dict(synth=synth, streams=streams, frames=frames, trk=trk, conf=0.30, nms=0.65)
    rec = mg.run_scenario(name, sc, pb, ref, cuda)

    class G:                       # the in-memory recording, shaped like an .npz
        files = list(rec.keys())
        def __getitem__(self, k): return rec[k]
    for s in streams:
        bad, ident = gu.check_stream_against(pb, orc, sc, name, s, gu.golden_stream(G(), name, s), replay=True)
        assert not bad, "\n".join(bad[:10])
        assert ident


def test_live_reference_auction_and_kf3_equal_checker(pb, orc, cuda):
    ref = _need_ref()
    rng = np.random.default_rng(99)
    for R, Cc in [(7, 9), (40, 40), (128, 64), (64, 128)]:
        cost = rng.uniform(0, 1, (R, Cc)).astype(np.float32)
        cost[rng.uniform(0, 1, (R, Cc)) < 0.5] = 1.0
        act = (rng.uniform(0, 1, R) < 0.7).astype(np.int32)
        r_row, r_col = ref.auction(cost, act)
        o_row, o_col = orc.auction(cost, act)
        assert np.array_equal(r_row, o_row) and np.array_equal(r_col, o_col), (R, Cc)
    T = 8
    rk, ok = ref.KF3(T), orc.KF3(T)
    dets = np.zeros((T, 17, 3), np.float32)
    dets[:, :, :2] = rng.uniform(10, 1200, (T, 17, 2)); dets[:, :, 2] = rng.uniform(0, 1, (T, 17))
    slots = np.arange(T, dtype=np.int32)
    rk.initiate(dets.reshape(T, 51), slots); ok.initiate(dets.reshape(T, 51), slots)
    for it in range(4):
        rk.predict(T, 0.9, 0.9); ok.predict(T, 0.9, 0.9)
        d2 = dets.copy(); d2[:, :, :2] += rng.normal(0, 2 + it, (T, 17, 2)).astype(np.float32)
        m = np.stack([np.arange(T), rng.permutation(T)], 1).astype(np.int32)[: 4 + it]
        rk.update(d2.reshape(T, 51), m); ok.update(d2.reshape(T, 51), m)
        rm, rd, off = rk.state(); om, od = ok.state()
        assert off == 0.0
        assert gu.close(rm, om, atol=gu.VEL_ATOL).all(), gu.max_rel(rm, om)
        assert gu.close(rd, od).all(), gu.max_rel(rd, od)
