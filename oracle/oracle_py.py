"""ctypes binding of the CPU checker (oracle/_build/libposebyte_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never by the product package."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PB_ORACLE_LIB") or os.path.join(_HERE, "_build", "libposebyte_oracle.so")
BUILD_FLAGS = "portable build (-O2 -march=x86-64-v3)"

TRACK_OUTPUT = np.dtype([("track_id", "<i4"), ("score", "<f4"), ("bbox", "<f4", (4,)),
                         ("keypoints", "<f4", (17, 3))])
POSE_DETECTION = np.dtype([("bbox", "<f4", (4,)), ("score", "<f4"), ("keypoints", "<f4", (17, 3))])


class TrackerConfig(C.Structure):
    _fields_ = [("max_tracks", C.c_int), ("max_detections", C.c_int), ("match_threshold", C.c_float),
                ("high_thresh", C.c_float), ("low_thresh", C.c_float), ("new_track_thresh", C.c_float),
                ("max_age", C.c_int), ("min_hits", C.c_int), ("gating_enabled", C.c_int)]


def tracker_config(max_tracks=128, max_detections=64, match_threshold=0.5, high_thresh=0.30, low_thresh=0.15,
                   new_track_thresh=0.30, max_age=10, min_hits=3, gating_enabled=1) -> TrackerConfig:
    return TrackerConfig(max_tracks, max_detections, match_threshold, high_thresh, low_thresh,
                         new_track_thresh, max_age, min_hits, gating_enabled)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} missing: run `make -C oracle`")
        L = C.CDLL(LIB_PATH)
        vp, ip, fp = C.c_void_p, C.c_int, C.c_float
        L.orc_decode.argtypes = [vp, ip, fp, ip, vp, vp, vp, vp]
        L.orc_nms_native.argtypes = [vp, vp, vp, ip, fp, ip, vp]
        L.orc_postprocess.argtypes = [vp, ip, fp, fp, ip, ip, vp, vp, vp, vp, vp, vp]
        L.orc_nms_legacy.argtypes = [vp, ip, fp, fp, vp]
        L.orc_pose_nms.argtypes = [vp, vp, vp, vp, ip, ip, fp, fp]
        L.orc_pose_nms.restype = None
        L.orc_auction.argtypes = [vp, ip, ip, vp, vp, vp]
        L.orc_auction.restype = None
        L.orc_pose_distance.argtypes = [vp, vp, ip, ip, ip, fp, vp]
        L.orc_pose_distance.restype = None
        L.orc_greedy_match.argtypes = [vp, ip, ip, fp, vp]
        L.orc_greedy_match.restype = None
        L.orc_assign_legacy.argtypes = [vp, ip, ip, fp, vp, vp]
        L.orc_letterbox.argtypes = [vp, ip, ip, ip, ip, vp, vp]
        L.orc_letterbox.restype = None
        L.orc_tracker_create.argtypes = [C.POINTER(TrackerConfig)]
        L.orc_tracker_create.restype = vp
        L.orc_tracker_destroy.argtypes = [vp]
        L.orc_tracker_destroy.restype = None
        L.orc_tracker_update.argtypes = [vp, vp, vp, ip, ip]
        L.orc_tracker_get_tracks.argtypes = [vp, vp, ip]
        L.orc_tracker_force_new.argtypes = [vp, vp, vp, ip]
        L.orc_tracker_force_new.restype = None
        L.orc_tracker_forced_errors.argtypes = [vp]
        L.orc_tracker_get_state.argtypes = [vp] * 16
        L.orc_tracker_get_state.restype = None
        L.orc_kf3_create.argtypes = [ip]
        L.orc_kf3_create.restype = vp
        L.orc_kf3_destroy.argtypes = [vp]
        L.orc_kf3_destroy.restype = None
        L.orc_kf3_initiate.argtypes = [vp, vp, vp, ip]
        L.orc_kf3_predict.argtypes = [vp, ip, fp, fp]
        L.orc_kf3_update.argtypes = [vp, vp, vp, ip]
        L.orc_kf3_extract.argtypes = [vp, vp, vp, ip]
        L.orc_kf3_get_state.argtypes = [vp, ip, vp, vp]
        L.orc_kf3_get_diag.argtypes = [vp, vp, vp]
        for f in (L.orc_kf3_initiate, L.orc_kf3_predict, L.orc_kf3_update, L.orc_kf3_extract,
                  L.orc_kf3_get_state, L.orc_kf3_get_diag):
            f.restype = None
        L.orc_expf_array.argtypes = [vp, vp, ip]
        L.orc_expf_array.restype = None
        L.orc_nms_mask.argtypes = [vp, vp, ip, fp, vp]
        L.orc_nms_mask.restype = None
        L.orc_run_streams.argtypes = [vp, ip, ip, ip, ip, fp, fp, ip, ip, C.POINTER(TrackerConfig), ip, vp, vp, vp]
        L.orc_run_streams.restype = C.c_double
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def expf(x: np.ndarray) -> np.ndarray:
    x = _f32(x).ravel()
    y = np.empty_like(x)
    lib().orc_expf_array(x.ctypes.data, y.ctypes.data, x.size)
    return y


def nms_mask(poses, bboxes, thr) -> np.ndarray:
    poses = _f32(poses).reshape(-1, 51); bboxes = _f32(bboxes).reshape(-1, 4)
    n = len(poses)
    m = np.zeros((n, n), np.uint8)
    lib().orc_nms_mask(poses.ctypes.data, bboxes.ctypes.data, n, thr, m.ctypes.data)
    return m


def decode(raw: np.ndarray, conf_thr: float, max_cand: int = 1024):
    raw = _f32(raw)
    N = raw.shape[-1]
    poses = np.zeros((max_cand, 51), np.float32); bboxes = np.zeros((max_cand, 4), np.float32)
    scores = np.zeros(max_cand, np.float32); anchors = np.zeros(max_cand, np.int32)
    c = lib().orc_decode(raw.ctypes.data, N, conf_thr, max_cand, poses.ctypes.data, bboxes.ctypes.data,
                         scores.ctypes.data, anchors.ctypes.data)
    return dict(poses=poses[:c], bboxes=bboxes[:c], scores=scores[:c], anchors=anchors[:c], num=c)


def postprocess(raw: np.ndarray, conf_thr=0.30, nms_thr=0.65, max_cand=1024, max_keep=256):
    raw = _f32(raw)
    N = raw.shape[-1]
    poses = np.zeros((max_cand, 51), np.float32); bboxes = np.zeros((max_cand, 4), np.float32)
    scores = np.zeros(max_cand, np.float32)
    slots = np.zeros(max_cand, np.int32); anchors = np.zeros(max_cand, np.int32)
    nc = C.c_int(0)
    k = lib().orc_postprocess(raw.ctypes.data, N, conf_thr, nms_thr, max_cand, max_keep, poses.ctypes.data,
                              bboxes.ctypes.data, scores.ctypes.data, slots.ctypes.data, anchors.ctypes.data,
                              C.addressof(nc))
    return dict(poses=poses[:k], bboxes=bboxes[:k], scores=scores[:k], keep_slots=slots[:k],
                keep_anchors=anchors[:k], num_keep=k, num_cand=nc.value)


def nms_legacy(dets: np.ndarray, oks_thr=0.65, score_thr=0.25) -> np.ndarray:
    dets = np.ascontiguousarray(dets, dtype=POSE_DETECTION)
    keep = np.zeros(max(len(dets), 1), np.int32)
    n = lib().orc_nms_legacy(dets.ctypes.data, len(dets), oks_thr, score_thr, keep.ctypes.data)
    return keep[:n].copy()


def pose_nms(poses, scores, sigmas, oks_thr, score_thr) -> np.ndarray:
    poses = _f32(poses); scores = _f32(scores); sigmas = _f32(sigmas)
    n = len(scores)
    keep = np.zeros(max(n, 1), np.int32)
    lib().orc_pose_nms(poses.ctypes.data, scores.ctypes.data, sigmas.ctypes.data, keep.ctypes.data, n,
                       len(sigmas), oks_thr, score_thr)
    return keep[:n]


def auction(cost: np.ndarray, row_active=None):
    cost = _f32(cost)
    R, Cc = cost.shape
    row = np.full(R, -1, np.int32); col = np.full(Cc, -1, np.int32)
    ra = None if row_active is None else np.ascontiguousarray(row_active, dtype=np.int32)
    lib().orc_auction(cost.ctypes.data, R, Cc, row.ctypes.data, col.ctypes.data, None if ra is None else ra.ctypes.data)
    return row, col


def pose_distance(tracks, dets, mode=0, alpha=0.7) -> np.ndarray:
    t = _f32(tracks).reshape(-1, 51); d = _f32(dets).reshape(-1, 51)
    out = np.zeros((len(t), len(d)), np.float32)
    lib().orc_pose_distance(t.ctypes.data, d.ctypes.data, len(t), len(d), mode, alpha, out.ctypes.data)
    return out


def greedy_match(cost, threshold) -> np.ndarray:
    c = _f32(cost)
    R, Cc = c.shape
    out = np.full(R, -1, np.int32)
    lib().orc_greedy_match(c.ctypes.data, R, Cc, threshold, out.ctypes.data)
    return out


def assign_legacy(cost, threshold):
    """LinearAssignmentCUDA::solve (hungarian.cu:235-339) -> (row, col, count)."""
    c = _f32(cost)
    R, Cc = c.shape
    row = np.full(R, -1, np.int32); col = np.full(Cc, -1, np.int32)
    n = lib().orc_assign_legacy(c.ctypes.data, R, Cc, threshold, row.ctypes.data, col.ctypes.data)
    return row, col, int(n)


def letterbox(bgr: np.ndarray, tw=640, th=640):
    """PreprocessorCUDA::preprocess (preprocess.cu:19-153): bgr [h,w,3] u8 -> ([3,th,tw] fp32, xform[4])."""
    img = np.ascontiguousarray(bgr, dtype=np.uint8)
    h, w, _ = img.shape
    out = np.empty((3, th, tw), np.float32); xf = np.empty(4, np.float32)
    lib().orc_letterbox(img.ctypes.data, w, h, tw, th, out.ctypes.data, xf.ctypes.data)
    return out, xf


class Tracker:
    def __init__(self, **kw):
        self.cfg = tracker_config(**kw)
        self._t = lib().orc_tracker_create(C.byref(self.cfg))
        self.T, self.Dm = self.cfg.max_tracks, self.cfg.max_detections

    def __del__(self):
        if getattr(self, "_t", None):
            try:
                lib().orc_tracker_destroy(self._t)
            except TypeError:           # interpreter shutdown: the module's globals are gone already
                pass
            self._t = None

    def update(self, det_poses, det_scores, frame_id: int) -> int:
        p = _f32(det_poses).reshape(-1, 51); s = _f32(det_scores)
        return lib().orc_tracker_update(self._t, p.ctypes.data, s.ctypes.data, len(s), frame_id)

    def force_new(self, slots, ids):
        """Replay mode: slot / id per detection index for the tracks the next update creates."""
        s = np.ascontiguousarray(slots, np.int32); i = np.ascontiguousarray(ids, np.int32)
        lib().orc_tracker_force_new(self._t, s.ctypes.data, i.ctypes.data, len(s))

    def forced_errors(self) -> int:
        return lib().orc_tracker_forced_errors(self._t)

    def get_tracks(self) -> np.ndarray:
        out = np.zeros(self.Dm, dtype=TRACK_OUTPUT)
        n = lib().orc_tracker_get_tracks(self._t, out.ctypes.data, self.Dm)
        return out[:n]

    def get_state(self) -> dict:
        T, Dm = self.T, self.Dm
        st = dict(poses=np.zeros((T, 51), np.float32), vel=np.zeros((T, 34), np.float32),
                  scores=np.zeros(T, np.float32), states=np.zeros(T, np.int32), ids=np.zeros(T, np.int32),
                  hits=np.zeros(T, np.int32), ages=np.zeros(T, np.int32), last_frame=np.zeros(T, np.int32),
                  active=np.zeros(T, np.int32), row_assign=np.zeros(T, np.int32),
                  col_assign=np.zeros(Dm, np.int32), cost=np.zeros(T * Dm, np.float32),
                  predicted=np.zeros((T, 51), np.float32), centers=np.zeros((T, 4), np.float32),
                  scalars=np.zeros(4, np.int32))
        order = ["poses", "vel", "scores", "states", "ids", "hits", "ages", "last_frame", "active",
                 "row_assign", "col_assign", "cost", "predicted", "centers", "scalars"]
        lib().orc_tracker_get_state(self._t, *[st[k].ctypes.data for k in order])
        return st


class KF3:
    def __init__(self, max_tracks: int):
        self.T = max_tracks
        self._k = lib().orc_kf3_create(max_tracks)

    def __del__(self):
        if getattr(self, "_k", None):
            lib().orc_kf3_destroy(self._k)
            self._k = None

    def initiate(self, dets, slots):
        d = _f32(dets).reshape(-1, 51); s = np.ascontiguousarray(slots, np.int32)
        lib().orc_kf3_initiate(self._k, d.ctypes.data, s.ctypes.data, len(s))

    def predict(self, n, am=0.9, jm=0.9):
        lib().orc_kf3_predict(self._k, n, am, jm)

    def update(self, dets, matches):
        d = _f32(dets).reshape(-1, 51); m = np.ascontiguousarray(matches, np.int32).reshape(-1, 2)
        lib().orc_kf3_update(self._k, d.ctypes.data, m.ctypes.data, len(m))

    def extract(self, slots):
        s = np.ascontiguousarray(slots, np.int32)
        out = np.zeros((len(s), 51), np.float32)
        lib().orc_kf3_extract(self._k, out.ctypes.data, s.ctypes.data, len(s))
        return out

    def state(self):
        m = np.zeros((self.T, 136), np.float32); d = np.zeros((self.T, 136), np.float32)
        lib().orc_kf3_get_diag(self._k, m.ctypes.data, d.ctypes.data)
        return m, d

    def full_state(self, track):
        m = np.zeros(136, np.float32); c = np.zeros((136, 136), np.float32)
        lib().orc_kf3_get_state(self._k, track, m.ctypes.data, c.ctypes.data)
        return m, c


def run_streams(heads: np.ndarray, frame_major: bool, conf_thr=0.30, nms_thr=0.65, max_cand=1024, max_keep=256,
                threads=1, **trk_kw):
    """heads [F,B,56,N] (frame_major) or [B,F,56,N].  Returns dict(wall_s, hashes[B], tracks_total, stage_s[3])."""
    heads = _f32(heads)
    if frame_major:
        F, B, _, N = heads.shape
    else:
        B, F, _, N = heads.shape
    cfg = tracker_config(**trk_kw)
    hashes = np.zeros(B, np.uint64)
    total = C.c_longlong(0)
    stage = np.zeros(3, np.float64)
    wall = lib().orc_run_streams(heads.ctypes.data, B, F, N, int(frame_major), conf_thr, nms_thr, max_cand,
                                 max_keep, C.byref(cfg), threads, hashes.ctypes.data, C.addressof(total),
                                 stage.ctypes.data)
    return dict(wall_s=wall, hashes=hashes, tracks_total=total.value, stage_s=stage)
