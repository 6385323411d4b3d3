"""Tracker restatement (gpu_tracker.cu): lifecycle rules, id issue order, persistence quirks."""
import numpy as np
import pytest

TEMPLATE = np.array([[0, -1.5], [-0.1, -1.6], [0.1, -1.6], [-0.2, -1.5], [0.2, -1.5], [-0.5, -1.0], [0.5, -1.0],
                     [-0.8, -0.3], [0.8, -0.3], [-1.0, 0.3], [1.0, 0.3], [-0.3, 0.0], [0.3, 0.0], [-0.3, 0.8],
                     [0.3, 0.8], [-0.3, 1.5], [0.3, 1.5]], np.float32)


def person(cx, cy, s=40.0, conf=0.9):
    p = np.zeros((17, 3), np.float32)
    p[:, 0] = cx + TEMPLATE[:, 0] * s
    p[:, 1] = cy + TEMPLATE[:, 1] * s
    p[:, 2] = conf
    return p.reshape(51)


def dets(centres, scores=None):
    poses = np.stack([person(x, y) for x, y in centres]) if centres else np.zeros((0, 51), np.float32)
    sc = np.array(scores if scores is not None else [0.9 - 0.01 * i for i in range(len(centres))], np.float32)
    return poses, sc


def test_ids_issued_in_detection_order_and_confirmation(orc):
    t = orc.Tracker()
    cs = [(100, 100), (300, 120), (500, 300)]
    for f in range(5):
        p, s = dets([(x + 2 * f, y) for x, y in cs])
        na = t.update(p, s, f)
        out = t.get_tracks()
        assert na == 3
        if f < 2:
            assert len(out) == 0                      # tentative, hits < 3 (gpu_tracker.cu:1600)
        else:
            assert list(out["track_id"]) == [1, 2, 3]  # R4: ids follow detection order
    st = t.get_state()
    assert list(st["states"][:3]) == [1, 1, 1] and list(st["hits"][:3]) == [5, 5, 5]
    assert st["scalars"][0] == 4 and st["scalars"][1] == 3          # next id, slot hint


def test_tentative_dies_after_three_misses_confirmed_goes_lost_then_removed(orc):
    t = orc.Tracker(max_age=4)
    p, s = dets([(100, 100)])
    t.update(p, s, 0)
    e = dets([])
    for f in range(1, 3):
        assert t.update(*e, f) == 1                    # age 1, 2
    assert t.update(*e, 3) == 0                        # tentative, age 3 > 2 -> removed
    # confirmed track
    t = orc.Tracker(max_age=4)
    for f in range(3):
        t.update(p, s, f)
    assert t.get_state()["states"][0] == 1
    states = []
    for f in range(3, 3 + 16):
        t.update(*e, f)
        st = t.get_state()
        states.append((int(st["active"][0]), int(st["states"][0]), int(st["ages"][0])))
    assert states[3] == (1, 1, 4) and states[4] == (1, 2, 5)        # age > max_age -> LOST
    assert states[13] == (1, 2, 14) and states[14][0] == 0          # age > max_age + 10 -> removed


def test_lost_track_is_recovered_in_tier3_with_same_id(orc):
    t = orc.Tracker(max_age=2)
    p, s = dets([(200, 200), (1500, 200)])     # far apart: outside each other's spatial gate
    for f in range(4):
        t.update(p, s, f)
    only_second = dets([(1500, 200)])
    for f in range(4, 9):
        t.update(*only_second, f)
    st = t.get_state()
    assert st["states"][0] == 2 and st["active"][0] == 1            # first track is LOST
    t.update(p, s, 9)                                                # it reappears nearby
    st = t.get_state()
    assert st["states"][0] == 1 and st["ages"][0] == 0              # LOST -> CONFIRMED (gpu_tracker.cu:644)
    out = t.get_tracks()
    assert sorted(out["track_id"]) == [1, 2]


def test_truncation_to_max_detections_and_slot_exhaustion(orc):
    t = orc.Tracker(max_tracks=4, max_detections=3)
    p, s = dets([(100 + 150 * i, 100) for i in range(6)])
    assert t.update(p, s, 0) == 3                      # only the first 3 detections are seen (:1066)
    t = orc.Tracker(max_tracks=2, max_detections=8)
    p, s = dets([(100 + 400 * i, 100) for i in range(5)])
    assert t.update(p, s, 0) == 2                      # two slots: detections 0 and 1 get them
    st = t.get_state()
    assert list(st["ids"]) == [1, 2] and st["scalars"][0] == 3
    assert st["scalars"][1] == 5                        # the hint advances for every qualifying detection
    assert list(st["col_assign"][:5]) == [0, 1, -1, -1, -1]


def test_stale_cells_let_unmatched_tracks_take_far_detections(orc):
    """Reference behaviour worth pinning: the match threshold is never applied
    (hungarian.cu:358-405) and gated-out cells keep old values (Q1), so an unmatched track whose
    row still holds the 1.0 written while the slot was inactive bids on any free detection."""
    t = orc.Tracker(max_tracks=4, max_detections=3)
    p, s = dets([(100 + 150 * i, 100) for i in range(3)])
    t.update(p, s, 0)
    far, sf = dets([(100 + 150 * i, 3000) for i in range(3)])
    assert t.update(far, sf, 1) == 3                   # no new tracks: the three old ones took the far detections
    st = t.get_state()
    assert sorted(st["row_assign"][:3]) == [0, 1, 2] and st["scalars"][0] == 4


def test_low_score_detections_do_not_start_tracks(orc):
    t = orc.Tracker(new_track_thresh=0.5)
    p, s = dets([(100, 100), (300, 100)], scores=[0.6, 0.4])
    assert t.update(p, s, 0) == 1


def test_duplicate_tracks_are_removed_keep_more_hits(orc):
    """Two confirmed tracks whose centre boxes overlap > 0.7 IoU: the one with fewer hits goes."""
    t = orc.Tracker()
    a, sa = dets([(200, 200)])
    for f in range(3):
        t.update(a, sa, f)
    both, sb = dets([(200, 200), (420, 200)])
    for f in range(3, 7):
        t.update(both, sb, f)
    # move the second person onto the first: both detections now at the same place
    near, sn = dets([(200, 200), (203, 200)])
    for f in range(7, 12):
        t.update(near, sn, f)
    st = t.get_state()
    assert st["active"][0] == 1 and st["active"][1] == 0            # track 2 (fewer hits) deactivated


def test_stale_cost_cells_persist_across_frames(orc):
    """Q1: a gated-out cell keeps what the last writer left (here the 1e9 of last frame's lock)."""
    t = orc.Tracker()
    p, s = dets([(100, 100), (2000, 2000)])
    for f in range(3):
        t.update(p, s, f)
    st = t.get_state()
    cost = st["cost"][: 128 * 2].reshape(128, 2)
    assert cost[0, 1] == np.float32(1e9) and cost[1, 0] == np.float32(1e9)   # far apart: never recomputed
    assert (cost[2:] == 1.0).all()                                            # inactive rows


def test_determinism(orc, pb):
    cfg = pb.synth_config(canvas=640, persons=10, period=40, occlusion=1)
    heads = pb.synth_heads(cfg, 0, 2, 0, 40, frame_major=False)
    a = orc.run_streams(heads, False, threads=1, max_age=5)
    b = orc.run_streams(heads, False, threads=2, max_age=5)
    assert np.array_equal(a["hashes"], b["hashes"]) and a["tracks_total"] == b["tracks_total"] > 0
    assert a["hashes"][0] != a["hashes"][1]


def test_gating_off_extension(orc):
    """gating_enabled=0 (config 5 'gating off'): far detections are still scored."""
    on, off = orc.Tracker(), orc.Tracker(gating_enabled=0)
    p, s = dets([(100, 100)])
    for f in range(3):
        on.update(p, s, f); off.update(p, s, f)
    far, sf = dets([(3000, 3000)])
    n_on, n_off = on.update(far, sf, 3), off.update(far, sf, 3)
    # gate on: the cell keeps last frame's lock value (1e9), no match, a second track starts;
    # gate off: the far pair is scored (cost ~1) and, with no threshold applied, matched.
    assert n_on == 2 and on.get_state()["row_assign"][0] == -1
    assert n_off == 1 and off.get_state()["row_assign"][0] == 0
