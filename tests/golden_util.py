"""Replay of the golden scenarios (tools/make_golden.py) through the CPU checker and comparison
rules shared by tests/test_golden.py (committed fixtures produced by the reference itself on a
B200) and tests/test_gpu_ref_crosscheck.py (the reference run live beside the checker).

Comparison rules (BASELINE.json north_star): kept detections, assignments and track life-cycle
counters must be EQUAL; track IDs equal (the reference issues them in atomics order, rule R4 of
DESIGN.md fixes ascending detection order — on the recorded runs both coincide, and the test
says so if they ever do not, by falling back to a consistent one-to-one renaming); keypoints,
boxes and filter states within 1e-4 relative."""
import numpy as np

RTOL = 1e-4
# Velocities are differences of pixel coordinates (|x| ~ 1e2..1e3 px, ulp ~ 6e-5 px): their
# absolute error is set by the coordinates' last bit, not by their own (small) magnitude.
VEL_ATOL = 5e-4

TRACK_OUTPUT = np.dtype([("track_id", "<i4"), ("score", "<f4"), ("bbox", "<f4", (4,)),
                         ("keypoints", "<f4", (17, 3))])


def close(a, b, rtol=RTOL, atol=0.0):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b)) + atol


def max_rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-30)
    return float(d.max()) if d.size else 0.0


class IdMap:
    """One-to-one renaming reference id <-> checker id, grown frame by frame."""

    def __init__(self):
        self.fwd, self.bwd, self.identity = {}, {}, True

    def bind(self, ref_id: int, got_id: int) -> bool:
        if ref_id != got_id:
            self.identity = False
        if self.fwd.setdefault(ref_id, got_id) != got_id:
            return False
        return self.bwd.setdefault(got_id, ref_id) == ref_id


def compare_kept(ref: dict, got: dict, where: str) -> list:
    """ref/got: dict(num_keep, scores[n], poses[n,51], bboxes[n,4]) in score order."""
    bad = []
    if ref["num_keep"] != got["num_keep"]:
        return [f"{where}: num_keep ref {ref['num_keep']} got {got['num_keep']}"]
    n = ref["num_keep"]
    # scores and keypoints are verbatim copies of head values on both sides: bit-exact
    if ref["scores"][:n].tobytes() != got["scores"][:n].tobytes():
        bad.append(f"{where}: kept scores differ")
    if ref["poses"][:n].tobytes() != got["poses"][:n].tobytes():
        bad.append(f"{where}: kept poses differ")
    if not close(ref["bboxes"][:n], got["bboxes"][:n]).all():
        bad.append(f"{where}: kept bboxes differ by {max_rel(ref['bboxes'][:n], got['bboxes'][:n])}")
    return bad


def compare_tracks(ref: np.ndarray, got: np.ndarray, ids: IdMap, where: str) -> list:
    """TrackOutput arrays of one frame (detection order on both sides)."""
    if len(ref) != len(got):
        return [f"{where}: {len(ref)} track outputs in the reference, {len(got)} here"]
    bad = []
    for i in range(len(ref)):
        if not ids.bind(int(ref["track_id"][i]), int(got["track_id"][i])):
            bad.append(f"{where}: output {i} id {got['track_id'][i]} is not a consistent renaming of {ref['track_id'][i]}")
        if ref["score"][i].tobytes() != got["score"][i].tobytes():
            bad.append(f"{where}: output {i} score differs")
        if not close(ref["keypoints"][i], got["keypoints"][i]).all():
            bad.append(f"{where}: output {i} keypoints differ by {max_rel(ref['keypoints'][i], got['keypoints'][i])}")
        if not close(ref["bbox"][i], got["bbox"][i]).all():
            bad.append(f"{where}: output {i} bbox differs by {max_rel(ref['bbox'][i], got['bbox'][i])}")
    return bad


def compare_final_state(ref: dict, got: dict, ids: IdMap, where: str) -> list:
    bad = []
    act = ref["active"] == 1
    if not np.array_equal(ref["active"], got["active"]):
        return [f"{where}: active slots differ"]
    for k in ("states", "hits", "ages", "last_frame"):
        if not np.array_equal(ref[k][act], got[k][act]):
            bad.append(f"{where}: {k} differs on active slots")
    for t in np.nonzero(act)[0]:
        if not ids.bind(int(ref["ids"][t]), int(got["ids"][t])):
            bad.append(f"{where}: slot {t} id {got['ids'][t]} vs reference {ref['ids'][t]} breaks the renaming")
    if not close(ref["poses"][act], got["poses"][act]).all():
        bad.append(f"{where}: track poses differ by {max_rel(ref['poses'][act], got['poses'][act])}")
    if not close(ref["vel"][act], got["vel"][act], atol=VEL_ATOL).all():
        bad.append(f"{where}: velocities differ by abs {np.abs(ref['vel'][act] - got['vel'][act]).max()}")
    if ref["scores"][act].tobytes() != got["scores"][act].tobytes():
        bad.append(f"{where}: track scores differ")
    if not np.array_equal(ref["scalars"][:2], got["scalars"][:2]):
        bad.append(f"{where}: next_id / slot hint ref {ref['scalars'][:2]} got {got['scalars'][:2]}")
    D = int(ref["scalars"][2])
    if not np.array_equal(ref["col_assign"][:D], got["col_assign"][:D]):
        bad.append(f"{where}: final col_assign differs")
    if not np.array_equal(ref["row_assign"][act], got["row_assign"][act]):
        bad.append(f"{where}: final row_assign differs on active rows")
    return bad


def golden_stream(G, name: str, s: int) -> dict:
    pre = f"{name}/s{s}/"
    d = {k[len(pre):]: G[k] for k in G.files if k.startswith(pre)}
    F, Dm = d["tracks"].shape[:2]
    d["tracks"] = np.ascontiguousarray(d["tracks"]).view(TRACK_OUTPUT).reshape(F, Dm)
    return d


FRAME_KEYS = ["states", "ids", "hits", "ages", "active", "row_assign", "col_assign"]


def recorded_new_tracks(gold: dict, f: int):
    """(slots[d], ids[d]) of the tracks the reference created in frame f, by detection index:
    the outcome of its slot/id atomics race (gpu_tracker.cu:715-722, :757), read off the
    recorded per-frame state.  A slot is new in frame f iff it is active with hits == 1 and
    age == 0 and it was inactive or carried another id in frame f-1."""
    D = int(gold["num_keep"][f])
    Dm = gold["frame_col_assign"].shape[1]
    D = min(D, Dm)
    slots = np.full(Dm, -1, np.int32); ids = np.zeros(Dm, np.int32)
    act, hits, ages, tid = gold["frame_active"][f], gold["frame_hits"][f], gold["frame_ages"][f], gold["frame_ids"][f]
    for d in range(D):
        t = int(gold["frame_col_assign"][f][d])
        if t < 0 or act[t] != 1 or hits[t] != 1 or ages[t] != 0:
            continue
        if f > 0 and gold["frame_active"][f - 1][t] == 1 and gold["frame_ids"][f - 1][t] == tid[t]:
            continue
        slots[d] = t; ids[d] = tid[t]
    return slots, ids


def replay_checker(pb, orc, sc: dict, s: int, gold: dict = None):
    """Yield per frame (kept dict, n_active, tracks, state) from the CPU checker.  With `gold`
    the checker runs in replay mode: new tracks take the slots/ids the reference gave them."""
    scfg = pb.synth_config(**sc["synth"])
    heads = pb.synth_heads(scfg, s, 1, 0, sc["frames"], frame_major=False)[0]
    trk = orc.Tracker(new_track_thresh=sc["conf"], high_thresh=sc["conf"], low_thresh=sc["conf"] / 2, **sc["trk"])
    for f in range(sc["frames"]):
        k = orc.postprocess(heads[f], sc["conf"], sc["nms"])
        if gold is not None:
            trk.force_new(*recorded_new_tracks(gold, f))
        na = trk.update(k["poses"], k["scores"], f)
        yield f, k, na, trk.get_tracks(), trk.get_state()
    if gold is not None:
        assert trk.forced_errors() == 0, f"replay: {trk.forced_errors()} recorded slots were unusable"


def compare_frame_state(gold: dict, f: int, st: dict, where: str) -> list:
    """Replay mode: the discrete tracker state after frame f must equal the recorded one."""
    bad = []
    act = gold["frame_active"][f] == 1
    if not np.array_equal(gold["frame_active"][f], st["active"]):
        return [f"{where}: active slots ref {np.nonzero(act)[0]} got {np.nonzero(st['active'] == 1)[0]}"]
    for k in ("states", "ids", "hits", "ages", "row_assign"):
        if not np.array_equal(gold["frame_" + k][f][act], st[k][act]):
            bad.append(f"{where}: {k} ref {gold['frame_' + k][f][act]} got {st[k][act]}")
    D = min(int(gold["num_keep"][f]), len(st["col_assign"]))
    if not np.array_equal(gold["frame_col_assign"][f][:D], st["col_assign"][:D]):
        bad.append(f"{where}: col_assign ref {gold['frame_col_assign'][f][:D]} got {st['col_assign'][:D]}")
    return bad


def check_stream_against(pb, orc, sc: dict, name: str, s: int, gold: dict, replay: bool = False) -> tuple:
    """Run one golden stream through the checker; returns (mismatch list, ids_identical).
    replay=False: rules R3/R4 decide slots and ids (ids compared up to a consistent renaming).
    replay=True: the recorded race outcome is supplied; every discrete output must be equal."""
    bad, ids = [], IdMap()
    last = sc["frames"] - 1
    for f, k, na, tracks, st in replay_checker(pb, orc, sc, s, gold if replay else None):
        w = f"{name} stream {s} frame {f}"
        n = int(gold["num_keep"][f])
        ref_k = dict(num_keep=n, scores=gold["kept_scores"][f], poses=gold["kept_poses"][f], bboxes=gold["kept_bboxes"][f])
        bad += compare_kept(ref_k, k, w)
        if int(gold["num_active"][f]) != na:
            bad.append(f"{w}: update() returned {na}, reference {gold['num_active'][f]}")
        bad += compare_tracks(gold["tracks"][f][: int(gold["num_tracks"][f])], tracks, ids, w)
        if replay:
            bad += compare_frame_state(gold, f, st, w)
        if f == last:
            ref_state = {k2[len("state_"):]: v for k2, v in gold.items() if k2.startswith("state_")}
            bad += compare_final_state(ref_state, st, ids, w)
        if len(bad) > 10:
            break
    if replay and not ids.identity:
        bad.append(f"{name} stream {s}: replay mode but ids are not identical")
    return bad, ids.identity
