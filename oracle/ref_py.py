"""ctypes binding of oracle/_ref/libposebyte_ref.so: the reference's OWN sources
(/root/reference/src/cuda/*.cu, unmodified) compiled for sm_100a behind oracle/ref_harness.cu.

TEST INFRASTRUCTURE ONLY.  Used to pin the CPU restatement (posebyte_oracle.cpp) against the real
reference: NMSCuda::apply runs on the host (no GPU needed); GPUPostprocess / GPUTracker /
LinearAssignmentCUDA / KalmanFilterCUDA need a B200 (tests marked gpu, tools/make_golden.py).
The library is built by `make -C oracle ref` where /root/reference exists and travels to the GPU
box as a built artefact; nothing here reads /root/reference at run time."""
from __future__ import annotations

import contextlib
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libposebyte_ref.so")

TRACK_OUTPUT = np.dtype([("track_id", "<i4"), ("score", "<f4"), ("bbox", "<f4", (4,)),
                         ("keypoints", "<f4", (17, 3))])
POSE_DETECTION = np.dtype([("bbox", "<f4", (4,)), ("score", "<f4"), ("keypoints", "<f4", (17, 3))])

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


@contextlib.contextmanager
def quiet():
    """The reference prints banners / CUDA_CHECK messages to stdout and stderr; silence them."""
    import sys
    sys.stdout.flush(); sys.stderr.flush()
    dn = os.open(os.devnull, os.O_WRONLY)
    s1, s2 = os.dup(1), os.dup(2)
    os.dup2(dn, 1); os.dup2(dn, 2)
    try:
        yield
    finally:
        os.dup2(s1, 1); os.dup2(s2, 2)
        os.close(dn); os.close(s1); os.close(s2)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        vp, ip, fp = C.c_void_p, C.c_int, C.c_float
        L.ref_post_create.argtypes = [ip, ip]; L.ref_post_create.restype = vp
        L.ref_post_destroy.argtypes = [vp]; L.ref_post_destroy.restype = None
        L.ref_post_process.argtypes = [vp, vp, fp, fp]
        L.ref_post_get.argtypes = [vp, ip, vp, vp, vp]; L.ref_post_get.restype = None
        L.ref_post_poses_dev.argtypes = [vp]; L.ref_post_poses_dev.restype = vp
        L.ref_post_scores_dev.argtypes = [vp]; L.ref_post_scores_dev.restype = vp
        L.ref_tracker_create.argtypes = [ip, ip, fp, fp, fp, fp, ip, ip]; L.ref_tracker_create.restype = vp
        L.ref_tracker_destroy.argtypes = [vp]; L.ref_tracker_destroy.restype = None
        L.ref_tracker_update.argtypes = [vp, vp, vp, ip, ip]
        L.ref_tracker_get_tracks.argtypes = [vp, vp, ip]
        L.ref_tracker_get_state.argtypes = [vp] * 16; L.ref_tracker_get_state.restype = None
        L.ref_nms_apply.argtypes = [vp, ip, fp, fp, vp]
        L.ref_auction.argtypes = [vp, ip, ip, vp, vp, vp]; L.ref_auction.restype = None
        L.ref_pose_distance.argtypes = [vp, vp, ip, ip, ip, fp, vp]; L.ref_pose_distance.restype = None
        L.ref_greedy_match.argtypes = [vp, ip, ip, fp, vp]
        L.ref_assign_solve.argtypes = [vp, ip, ip, fp, vp, vp]
        L.ref_preprocess.argtypes = [vp, ip, ip, ip, ip, vp, vp]; L.ref_preprocess.restype = None
        L.ref_kf3_create.argtypes = [ip]; L.ref_kf3_create.restype = vp
        L.ref_kf3_destroy.argtypes = [vp]; L.ref_kf3_destroy.restype = None
        L.ref_kf3_initiate.argtypes = [vp, vp, vp, ip]; L.ref_kf3_initiate.restype = None
        L.ref_kf3_predict.argtypes = [vp, ip, fp, fp]; L.ref_kf3_predict.restype = None
        L.ref_kf3_update.argtypes = [vp, vp, ip, vp, ip]; L.ref_kf3_update.restype = None
        L.ref_kf3_get.argtypes = [vp, ip, vp, vp]; L.ref_kf3_get.restype = C.c_float
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def nms_apply(dets: np.ndarray, oks_thr=0.65, score_thr=0.25) -> np.ndarray:
    """NMSCuda::apply (reference nms.cu:142-306), host code."""
    dets = np.ascontiguousarray(dets, dtype=POSE_DETECTION)
    keep = np.zeros(max(len(dets), 1), np.int32)
    with quiet():
        n = lib().ref_nms_apply(dets.ctypes.data, len(dets), oks_thr, score_thr, keep.ctypes.data)
    return keep[:n].copy()


def auction(cost: np.ndarray, row_active=None):
    """LinearAssignmentCUDA::solveDeviceAsyncWithActive on the GPU (hungarian.cu:358-405)."""
    cost = _f32(cost)
    R, Cc = cost.shape
    row = np.full(R, -1, np.int32); col = np.full(Cc, -1, np.int32)
    ra = None if row_active is None else np.ascontiguousarray(row_active, dtype=np.int32)
    with quiet():
        lib().ref_auction(cost.ctypes.data, R, Cc, row.ctypes.data, col.ctypes.data, None if ra is None else ra.ctypes.data)
    return row, col


def _as_dets(poses) -> np.ndarray:
    p = _f32(poses).reshape(-1, 17, 3)
    d = np.zeros(len(p), POSE_DETECTION)
    d["keypoints"] = p
    return d


def pose_distance(tracks, dets, mode=0, alpha=0.7) -> np.ndarray:
    """OKSDistanceCUDA::computeOKSDistance / computeIoUDistance / computeCombinedDistance (GPU)."""
    t, d = _as_dets(tracks), _as_dets(dets)
    out = np.zeros((len(t), len(d)), np.float32)
    with quiet():
        lib().ref_pose_distance(t.ctypes.data, d.ctypes.data, len(t), len(d), mode, alpha, out.ctypes.data)
    return out


def assign_solve(cost, threshold):
    """LinearAssignmentCUDA::solve (host entry point) -> (row, col, count)."""
    c = _f32(cost)
    R, Cc = c.shape
    row = np.full(R, -1, np.int32); col = np.full(Cc, -1, np.int32)
    n = lib().ref_assign_solve(c.ctypes.data, R, Cc, threshold, row.ctypes.data, col.ctypes.data)
    return row, col, int(n)


def preprocess(bgr, tw=640, th=640):
    """PreprocessorCUDA::preprocess -> ([3,th,tw] fp32, xform[4])."""
    img = np.ascontiguousarray(bgr, dtype=np.uint8)
    h, w, _ = img.shape
    out = np.empty((3, th, tw), np.float32); xf = np.empty(4, np.float32)
    lib().ref_preprocess(img.ctypes.data, w, h, tw, th, out.ctypes.data, xf.ctypes.data)
    return out, xf


def greedy_match(cost, threshold) -> np.ndarray:
    """GreedyMatcherCUDA::match (deterministic host path below 200 cells)."""
    c = _f32(cost)
    R, Cc = c.shape
    out = np.full(R, -1, np.int32)
    with quiet():
        lib().ref_greedy_match(c.ctypes.data, R, Cc, threshold, out.ctypes.data)
    return out


class Postprocess:
    """GPUPostprocess (gpu_postprocess.cu:319-476) on the current CUDA device."""

    def __init__(self, max_detections=1024, num_anchors=8400):
        with quiet():
            self._p = lib().ref_post_create(max_detections, num_anchors)

    def __del__(self):
        if getattr(self, "_p", None):
            lib().ref_post_destroy(self._p)
            self._p = None

    def process(self, d_raw_ptr: int, conf=0.30, nms=0.65) -> dict:
        with quiet():
            n = lib().ref_post_process(self._p, d_raw_ptr, conf, nms)
        poses = np.zeros((max(n, 1), 51), np.float32); bboxes = np.zeros((max(n, 1), 4), np.float32)
        scores = np.zeros(max(n, 1), np.float32)
        lib().ref_post_get(self._p, n, poses.ctypes.data, bboxes.ctypes.data, scores.ctypes.data)
        return dict(num_keep=n, poses=poses[:n], bboxes=bboxes[:n], scores=scores[:n])

    def poses_dev(self) -> int:
        return lib().ref_post_poses_dev(self._p)

    def scores_dev(self) -> int:
        return lib().ref_post_scores_dev(self._p)


class Tracker:
    """GPUTracker (gpu_tracker.cu) on the current CUDA device."""

    def __init__(self, max_tracks=128, max_detections=64, match_threshold=0.5, high_thresh=0.30, low_thresh=0.15,
                 new_track_thresh=0.30, max_age=10, min_hits=3):
        self.T, self.Dm = max_tracks, max_detections
        with quiet():
            self._t = lib().ref_tracker_create(max_tracks, max_detections, match_threshold, high_thresh, low_thresh,
                                               new_track_thresh, max_age, min_hits)

    def __del__(self):
        if getattr(self, "_t", None):
            with quiet():
                lib().ref_tracker_destroy(self._t)
            self._t = None

    def update(self, d_poses_ptr: int, d_scores_ptr: int, n: int, frame_id: int) -> int:
        with quiet():
            return lib().ref_tracker_update(self._t, d_poses_ptr, d_scores_ptr, n, frame_id)

    def get_tracks(self) -> np.ndarray:
        out = np.zeros(self.Dm, dtype=TRACK_OUTPUT)
        with quiet():
            n = lib().ref_tracker_get_tracks(self._t, out.ctypes.data, self.Dm)
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
        lib().ref_tracker_get_state(self._t, *[st[k].ctypes.data for k in order])
        return st


class KF3:
    """KalmanFilterCUDA (kalman_filter.cu) on the current CUDA device."""

    def __init__(self, max_tracks: int):
        self.T = max_tracks
        with quiet():
            self._k = lib().ref_kf3_create(max_tracks)

    def __del__(self):
        if getattr(self, "_k", None):
            with quiet():
                lib().ref_kf3_destroy(self._k)
            self._k = None

    def initiate(self, dets, slots):
        d = _f32(dets).reshape(-1, 51); s = np.ascontiguousarray(slots, np.int32)
        lib().ref_kf3_initiate(self._k, d.ctypes.data, s.ctypes.data, len(s))

    def predict(self, n, am=0.9, jm=0.9):
        lib().ref_kf3_predict(self._k, n, am, jm)

    def update(self, dets, matches):
        d = _f32(dets).reshape(-1, 51); m = np.ascontiguousarray(matches, np.int32).reshape(-1, 2)
        lib().ref_kf3_update(self._k, d.ctypes.data, len(d), m.ctypes.data, len(m))

    def state(self):
        """means [T,136], covariance diagonal [T,136], largest |off-diagonal| element."""
        m = np.zeros((self.T, 136), np.float32); d = np.zeros((self.T, 136), np.float32)
        off = lib().ref_kf3_get(self._k, self.T, m.ctypes.data, d.ctypes.data)
        return m, d, float(off)
