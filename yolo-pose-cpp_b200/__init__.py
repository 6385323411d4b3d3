"""posebyte-b200: B200-native PoseBYTE post-inference path (decode -> pose-NMS -> tracker).

Python is the harness only: this module is a ctypes binding of the C ABI declared in
``include/posebyte_b200.h`` (the product is ``lib/libposebyte_b200.so``, hand-written
CUDA for sm_100a behind a C++17 host layer).  It mirrors the reference's call sequence
(``GPUPostprocess::process`` -> ``GPUTracker::update`` -> ``getActiveTracks``,
reference src/main.cpp:207-224) for B independent streams per call.

There is no CPU fallback: creating a :class:`Pipeline` without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.environ.get("PB_LIB_PATH") or os.path.join(LIB_DIR, "libposebyte_b200.so")   # PB_LIB_PATH: development builds (telemetry)
SYNTH_PATH = os.path.join(LIB_DIR, "libpb_synth.so")

PB_OK, PB_ERR_INVALID, PB_ERR_CUDA, PB_ERR_UNSUPPORTED, PB_ERR_NO_DEVICE = 0, -1, -2, -3, -4

# TrackOutput (include/types.h): 228-byte record.
TRACK_OUTPUT = np.dtype([("track_id", "<i4"), ("score", "<f4"), ("bbox", "<f4", (4,)),
                         ("keypoints", "<f4", (17, 3))])
assert TRACK_OUTPUT.itemsize == 228
# PoseDetection (include/types.h): 224-byte record.
POSE_DETECTION = np.dtype([("bbox", "<f4", (4,)), ("score", "<f4"), ("keypoints", "<f4", (17, 3))])
assert POSE_DETECTION.itemsize == 224


class PbError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"posebyte_b200 status {status}: {msg}")
        self.status = status


class PbConfig(C.Structure):
    _fields_ = [("num_streams", C.c_int), ("num_anchors", C.c_int), ("max_candidates", C.c_int),
                ("max_keep", C.c_int), ("max_tracks", C.c_int), ("max_detections", C.c_int),
                ("match_threshold", C.c_float), ("high_thresh", C.c_float), ("low_thresh", C.c_float),
                ("new_track_thresh", C.c_float), ("max_age", C.c_int), ("min_hits", C.c_int),
                ("use_cuda_graph", C.c_int), ("gating_enabled", C.c_int), ("device", C.c_int),
                ("pipeline_depth", C.c_int), ("keypoint_fetch", C.c_int), ("fuse_stages", C.c_int)]


class PbTiming(C.Structure):
    _fields_ = [(n, C.c_longlong) for n in ("predict_us", "gate_us", "high_assoc_us", "low_assoc_us",
                                            "lost_assoc_us", "update_us", "age_us", "new_track_us",
                                            "dedup_us", "total_us")] + [("frame_count", C.c_int)]


class PbDeviceViews(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("det_poses", "det_bboxes", "det_scores", "num_keep", "num_cand",
                                          "keep_slots", "keep_anchors", "track_poses", "track_scores",
                                          "track_states", "track_ids", "track_outputs", "num_outputs",
                                          "num_active")]


class SynthConfig(C.Structure):
    _fields_ = [("num_anchors", C.c_int), ("canvas", C.c_int), ("persons", C.c_int), ("period", C.c_int),
                ("clumps", C.c_int), ("occlusion", C.c_int), ("kp_drop_prob", C.c_float),
                ("max_speed", C.c_float), ("seed", C.c_ulonglong)]


# Every symbol include/posebyte_b200.h declares (tests check the library exports them all).
ABI_SYMBOLS = [
    "pb_last_error", "pb_version", "pb_default_config", "pb_create", "pb_destroy", "pb_reset",
    "pb_postprocess", "pb_tracker_update", "pb_step", "pb_step_seq", "pb_step_path", "pb_nms_plan", "pb_join", "pb_step_host", "pb_submit_host", "pb_wait", "pb_get_tracks",
    "pb_get_tracks_all", "pb_get_num_active", "pb_get_kept", "pb_get_state", "pb_get_device_views",
    "pb_get_timing", "pb_get_stream_stage_ns", "pb_debug_timeline", "pb_launch_count", "pb_set_profiling", "pb_get_nms_path_counts", "pb_get_kernel_ms", "pb_get_kernel_us", "pb_get_post_stage_us", "launchPoseNMS", "pb_nms_legacy", "pb_auction_solve",
    "pb_set_output_transform", "pb_state_size", "pb_state_save", "pb_state_load", "pb_pose_distance", "pb_greedy_match",
    "pb_assign_legacy", "pb_letterbox_batch",
    "pb_kf3_initiate", "pb_kf3_predict", "pb_kf3_update", "pb_kf3_extract", "pb_kf3_materialize_cov",
]

_lib = None
_synth = None


def lib() -> C.CDLL:
    """The product library.  Raises if it has not been built (``__graft_entry__.build()``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        L.pb_last_error.restype = C.c_char_p
        L.pb_version.restype = C.c_char_p
        L.pb_launch_count.restype = C.c_longlong
        L.launchPoseNMS.restype = None
        vp, ip, fp = C.c_void_p, C.c_int, C.c_float
        L.pb_default_config.argtypes = [C.POINTER(PbConfig)]
        L.pb_create.argtypes = [C.POINTER(PbConfig), C.POINTER(vp)]
        L.pb_destroy.argtypes = [vp]
        L.pb_reset.argtypes = [vp, vp]
        L.pb_postprocess.argtypes = [vp, vp, fp, fp, vp]
        L.pb_tracker_update.argtypes = [vp, vp, vp, vp, ip, ip, vp]
        L.pb_step.argtypes = [vp, vp, fp, fp, ip, vp]
        L.pb_step_seq.argtypes = [vp, vp, C.c_size_t, ip, ip, ip, fp, fp, ip, vp]
        L.pb_step_path.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.pb_nms_plan.argtypes = [vp] + [C.POINTER(C.c_int)] * 4
        L.pb_step_host.argtypes = [vp, vp, fp, fp, ip, vp, vp]
        L.pb_join.argtypes = [vp, vp]
        L.pb_submit_host.argtypes = [vp, vp, fp, fp, ip, vp, vp]
        L.pb_wait.argtypes = [vp]
        L.pb_get_tracks.argtypes = [vp, ip, vp, ip, C.POINTER(ip)]
        L.pb_get_tracks_all.argtypes = [vp, vp, vp]
        L.pb_get_num_active.argtypes = [vp, vp]
        L.pb_get_kept.argtypes = [vp, ip, vp, vp, vp, vp, vp, ip, C.POINTER(ip), C.POINTER(ip)]
        L.pb_get_state.argtypes = [vp, ip] + [vp] * 15
        L.pb_get_device_views.argtypes = [vp, C.POINTER(PbDeviceViews)]
        L.pb_get_timing.argtypes = [vp, C.POINTER(PbTiming)]
        L.pb_get_stream_stage_ns.argtypes = [vp, vp]
        L.pb_debug_timeline.argtypes = [vp, vp]
        L.pb_get_post_stage_us.argtypes = [vp, C.POINTER(C.c_double * 5)]
        L.pb_set_profiling.argtypes = [vp, ip]
        L.pb_get_kernel_ms.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(ip), C.POINTER(C.c_double), C.POINTER(ip)]
        L.pb_get_nms_path_counts.argtypes = [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        L.pb_get_kernel_us.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(ip)]
        L.launchPoseNMS.argtypes = [vp, vp, vp, vp, ip, ip, fp, fp, vp]
        L.pb_nms_legacy.argtypes = [vp, vp, ip, ip, fp, fp, vp, vp, vp]
        L.pb_auction_solve.argtypes = [vp, ip, ip, ip, vp, vp, vp, vp]
        L.pb_set_output_transform.argtypes = [vp, vp]
        L.pb_state_size.argtypes = [vp, C.POINTER(C.c_size_t)]
        L.pb_state_save.argtypes = [vp, vp, C.c_size_t]
        L.pb_state_load.argtypes = [vp, vp, C.c_size_t]
        L.pb_pose_distance.argtypes = [vp, vp, ip, ip, ip, ip, fp, vp, vp]
        L.pb_greedy_match.argtypes = [vp, ip, ip, ip, fp, vp, vp]
        L.pb_assign_legacy.argtypes = [vp, ip, ip, ip, fp, vp, vp, vp, vp]
        L.pb_letterbox_batch.argtypes = [vp, C.c_size_t, vp, ip, ip, ip, vp, vp, vp]
        L.pb_kf3_initiate.argtypes = [vp, vp, vp, vp, ip, vp]
        L.pb_kf3_predict.argtypes = [vp, vp, ip, fp, fp, vp]
        L.pb_kf3_update.argtypes = [vp, vp, vp, vp, ip, vp]
        L.pb_kf3_extract.argtypes = [vp, vp, vp, ip, vp]
        L.pb_kf3_materialize_cov.argtypes = [vp, ip, vp, vp]
        _lib = L
    return _lib


def check(status: int) -> None:
    if status != PB_OK:
        raise PbError(status, lib().pb_last_error().decode())


def default_config(**kw) -> PbConfig:
    cfg = PbConfig()
    lib().pb_default_config(C.byref(cfg))
    for k, v in kw.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg


def _ptr(x):
    """Device/host address of a torch tensor, numpy array, int or None."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    return x.data_ptr()          # torch.Tensor


def _dev_f32(x, numel: int, device: int, what: str):
    """Address of a device buffer of at least `numel` contiguous float32 elements on `device`.  Torch tensors are
    checked (dtype, device, contiguity, size); a raw int address is taken on trust, as in C."""
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray) or not hasattr(x, "data_ptr"):
        raise ValueError(f"{what}: expected a CUDA tensor (or a raw device address), got {type(x).__name__}")
    import torch
    if x.dtype != torch.float32:
        raise ValueError(f"{what}: dtype {x.dtype}, expected float32")
    if not x.is_cuda or x.device.index != device:
        raise ValueError(f"{what}: tensor on {x.device}, the handle lives on cuda:{device}")
    if not x.is_contiguous():
        raise ValueError(f"{what}: tensor is not contiguous")
    if x.numel() < numel:
        raise ValueError(f"{what}: {x.numel()} elements, at least {numel} needed")
    return x.data_ptr()


def _host_array(x, dtype, nbytes: int, what: str, writable=False):
    if not isinstance(x, np.ndarray):
        raise ValueError(f"{what}: expected a numpy array")
    dtypes = dtype if isinstance(dtype, tuple) else (dtype,)
    if not any(x.dtype == np.dtype(d) for d in dtypes):
        raise ValueError(f"{what}: dtype {x.dtype}, expected {' or '.join(str(np.dtype(d)) for d in dtypes)}")
    if not x.flags.c_contiguous:
        raise ValueError(f"{what}: array is not C-contiguous")
    if x.nbytes < nbytes:
        raise ValueError(f"{what}: {x.nbytes} bytes, at least {nbytes} needed")
    if writable and not x.flags.writeable:
        raise ValueError(f"{what}: array is read-only")
    return x.ctypes.data


def _stream_ptr(stream):
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, int):
        return stream
    return stream.cuda_stream


class Pipeline:
    """B streams of decode + NMS + PoseBYTE tracking on one GPU (one handle)."""

    def __init__(self, **cfg_kw):
        self.cfg = default_config(**cfg_kw)
        self._h = C.c_void_p()
        check(lib().pb_create(C.byref(self.cfg), C.byref(self._h)))
        self.B = self.cfg.num_streams
        self.T = self.cfg.max_tracks
        self.Dm = self.cfg.max_detections
        self._head_numel = self.B * 56 * self.cfg.num_anchors

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().pb_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self, stream=None):
        check(lib().pb_reset(self._h, _stream_ptr(stream)))

    # -- the path ---------------------------------------------------------------------
    def postprocess(self, heads, conf=0.30, nms=0.65, stream=None):
        check(lib().pb_postprocess(self._h, _dev_f32(heads, self._head_numel, self.cfg.device, "heads"), conf, nms, _stream_ptr(stream)))

    def tracker_update(self, frame_id, det_poses=None, det_scores=None, num_dets=None, det_stride=0, stream=None):
        if det_poses is not None and not isinstance(det_poses, int):
            dev = self.cfg.device
            _dev_f32(det_poses, self.B * det_stride * 51, dev, "det_poses")
            _dev_f32(det_scores, self.B * det_stride, dev, "det_scores")
            import torch
            if not (hasattr(num_dets, "dtype") and num_dets.dtype == torch.int32 and num_dets.is_cuda and num_dets.device.index == dev
                    and num_dets.is_contiguous() and num_dets.numel() >= self.B):
                raise ValueError("num_dets: expected a contiguous int32 CUDA tensor with one count per stream")
        check(lib().pb_tracker_update(self._h, _ptr(det_poses), _ptr(det_scores), _ptr(num_dets), det_stride,
                                      frame_id, _stream_ptr(stream)))

    def step(self, heads, frame_id, conf=0.30, nms=0.65, stream=None):
        check(lib().pb_step(self._h, _dev_f32(heads, self._head_numel, self.cfg.device, "heads"), conf, nms, frame_id, _stream_ptr(stream)))

    def step_seq(self, heads, first, n_steps, frame0, conf=0.30, nms=0.65, stream=None):
        """n_steps steps on the batches heads[(first + i) % len(heads)] ([F,B,56,N] CUDA tensor), frame ids frame0 + i."""
        F = heads.shape[0]
        ptr = _dev_f32(heads, F * self._head_numel, self.cfg.device, "heads")
        check(lib().pb_step_seq(self._h, ptr, self._head_numel, F, first % F, n_steps, conf, nms, frame0, _stream_ptr(stream)))

    def step_path(self) -> dict:
        """How this handle runs its steps: per-step path and the chunk of the resident-tracker path of step_seq (0: none)."""
        a, b = C.c_int(0), C.c_int(0)
        check(lib().pb_step_path(self._h, C.byref(a), C.byref(b)))
        return {"per_step": ("serial", "pipelined three-kernel step", "fused per-stream kernel")[a.value], "seq_chunk": b.value}

    def nms_plan(self) -> dict:
        """Launch plan of the NMS kernel of the per-step paths (pb_nms_plan)."""
        v = [C.c_int(0) for _ in range(4)]
        check(lib().pb_nms_plan(self._h, *[C.byref(x) for x in v]))
        return dict(threads=v[0].value, ctas_per_sm=v[1].value, smem_bytes=v[2].value, tier_candidates=v[3].value)

    def join(self, stream=None):
        """Make `stream` wait for work a pipelined step left on the internal streams."""
        check(lib().pb_join(self._h, _stream_ptr(stream)))

    def step_host(self, heads_np: np.ndarray, frame_id, conf=0.30, nms=0.65, out=None, counts=None):
        """Host buffers in, TrackOutput records out.  `heads_np` (and optionally `out` [B,Dm]
        TRACK_OUTPUT / `counts` [B] int32) may be views of page-locked memory: the input is then
        read in place by the GPU and the results are copied straight into `out` / `counts`."""
        if out is None:
            out = np.zeros((self.B, self.Dm), dtype=TRACK_OUTPUT)
        if counts is None:
            counts = np.zeros(self.B, dtype=np.int32)
        hp = _host_array(heads_np, np.float32, self._head_numel * 4, "heads")
        op = _host_array(out, (TRACK_OUTPUT, np.uint8), self.B * self.Dm * 228, "out", writable=True)
        cp = _host_array(counts, np.int32, self.B * 4, "counts", writable=True)
        check(lib().pb_step_host(self._h, hp, conf, nms, frame_id, op, cp))
        return out, counts

    def submit_host(self, heads_np: np.ndarray, frame_id, out: np.ndarray, counts: np.ndarray, conf=0.30, nms=0.65):
        """Asynchronous step_host: all three arrays must be views of page-locked memory and stay
        untouched until wait() returns."""
        hp = _host_array(heads_np, np.float32, self._head_numel * 4, "heads")
        op = _host_array(out, (TRACK_OUTPUT, np.uint8), self.B * self.Dm * 228, "out", writable=True)
        cp = _host_array(counts, np.int32, self.B * 4, "counts", writable=True)
        check(lib().pb_submit_host(self._h, hp, conf, nms, frame_id, op, cp))

    def wait(self):
        check(lib().pb_wait(self._h))

    # -- rows next to the path ---------------------------------------------------------
    def set_output_transform(self, xform):
        """xform [B,4] = scale_x, scale_y, pad_x, pad_y per stream (scaleTrackOutputs), or None."""
        if xform is None:
            check(lib().pb_set_output_transform(self._h, None))
        else:
            x = np.ascontiguousarray(xform, np.float32).reshape(self.B, 4)
            check(lib().pb_set_output_transform(self._h, x.ctypes.data))

    def state_save(self) -> bytes:
        n = C.c_size_t(0)
        check(lib().pb_state_size(self._h, C.byref(n)))
        buf = np.zeros(n.value, np.uint8)
        check(lib().pb_state_save(self._h, buf.ctypes.data, n.value))
        return buf.tobytes()

    def state_load(self, blob: bytes):
        n = C.c_size_t(0)
        check(lib().pb_state_size(self._h, C.byref(n)))
        if len(blob) < n.value:
            raise PbError(PB_ERR_INVALID, f"state_load: blob of {len(blob)} bytes, the handle's snapshots have {n.value}")
        buf = np.frombuffer(blob, np.uint8)
        check(lib().pb_state_load(self._h, buf.ctypes.data, buf.size))

    # -- results ----------------------------------------------------------------------
    def get_tracks(self, b: int) -> np.ndarray:
        out = np.zeros(self.Dm, dtype=TRACK_OUTPUT)
        n = C.c_int(0)
        check(lib().pb_get_tracks(self._h, b, out.ctypes.data, self.Dm, C.byref(n)))
        return out[: n.value]

    def get_tracks_all(self):
        out = np.zeros((self.B, self.Dm), dtype=TRACK_OUTPUT)
        counts = np.zeros(self.B, dtype=np.int32)
        check(lib().pb_get_tracks_all(self._h, out.ctypes.data, counts.ctypes.data))
        return out, counts

    def get_num_active(self) -> np.ndarray:
        out = np.zeros(self.B, dtype=np.int32)
        check(lib().pb_get_num_active(self._h, out.ctypes.data))
        return out

    def get_kept(self, b: int) -> dict:
        K = self.cfg.max_keep
        poses = np.zeros((K, 51), np.float32); bboxes = np.zeros((K, 4), np.float32)
        scores = np.zeros(K, np.float32); slots = np.zeros(K, np.int32); anchors = np.zeros(K, np.int32)
        nk, nc = C.c_int(0), C.c_int(0)
        check(lib().pb_get_kept(self._h, b, poses.ctypes.data, bboxes.ctypes.data, scores.ctypes.data,
                                slots.ctypes.data, anchors.ctypes.data, K, C.byref(nk), C.byref(nc)))
        n = nk.value
        return dict(poses=poses[:n], bboxes=bboxes[:n], scores=scores[:n], keep_slots=slots[:n],
                    keep_anchors=anchors[:n], num_keep=n, num_cand=nc.value)

    def get_state(self, b: int) -> dict:
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
        check(lib().pb_get_state(self._h, b, *[st[k].ctypes.data for k in order]))
        return st

    def device_views(self) -> PbDeviceViews:
        v = PbDeviceViews()
        check(lib().pb_get_device_views(self._h, C.byref(v)))
        return v

    def post_stage_us(self) -> dict:
        a = (C.c_double * 5)()
        check(lib().pb_get_post_stage_us(self._h, C.byref(a)))
        return dict(zip(("list", "rank", "load", "nms", "output"), [round(x, 2) for x in a]))

    def tracker_stage_us(self) -> dict:
        t = self.timing()
        n = max(t.frame_count, 1)
        return {k: round(getattr(t, k) / n, 2) for k, _ in t._fields_ if k != "frame_count"}

    def set_profiling(self, on: bool):
        check(lib().pb_set_profiling(self._h, int(on)))

    def kernel_ms(self) -> dict:
        """Summed per-kernel device milliseconds since the last call (needs set_profiling(True))."""
        pm, tm, pn, tn = C.c_double(0), C.c_double(0), C.c_int(0), C.c_int(0)
        check(lib().pb_get_kernel_ms(self._h, C.byref(pm), C.byref(pn), C.byref(tm), C.byref(tn)))
        return dict(post_ms=pm.value, post_launches=pn.value, track_ms=tm.value, track_launches=tn.value)

    def debug_timeline(self) -> np.ndarray:
        out = np.zeros((64, self.cfg.num_streams, 6), np.uint64)
        check(lib().pb_debug_timeline(self._h, out.ctypes.data))
        return out

    def stream_stage_ns(self) -> np.ndarray:
        """[B,20] per-stream nanosecond accumulators of the tracker stages (see the header)."""
        out = np.zeros((self.B, 20), np.uint64)
        check(lib().pb_get_stream_stage_ns(self._h, out.ctypes.data))
        return out

    def nms_path_counts(self) -> dict:
        """Stream-frames the NMS kernel processed, and how many 64-rank tiles of the lazy sweep
        needed their second round (keypoints for every live rank)."""
        f, c, k = C.c_longlong(0), C.c_longlong(0), C.c_longlong(0)
        check(lib().pb_get_nms_path_counts(self._h, C.byref(f), C.byref(c), C.byref(k)))
        return dict(stream_frames=f.value, second_rounds=c.value, keypoint_fetches=k.value)

    def kernel_us(self) -> dict:
        """Mean device microseconds per launch of the three kernels since the last call (needs set_profiling(True))."""
        g, n, t, k = C.c_double(0), C.c_double(0), C.c_double(0), C.c_int(0)
        check(lib().pb_get_kernel_us(self._h, C.byref(g), C.byref(n), C.byref(t), C.byref(k)))
        return dict(gather_us=g.value, nms_us=n.value, track_us=t.value, launches=k.value)

    def timing(self) -> PbTiming:
        t = PbTiming()
        check(lib().pb_get_timing(self._h, C.byref(t)))
        return t


def pose_distance(tracks, dets, mode=0, alpha=0.7, stream=None):
    """tracks [batch, nt, 51], dets [batch, nd, 51] CUDA tensors -> costs [batch, nt, nd]."""
    import torch
    batch, nt, _ = tracks.shape
    nd = dets.shape[1]
    out = torch.empty(batch, nt, nd, device=tracks.device, dtype=torch.float32)
    check(lib().pb_pose_distance(tracks.data_ptr(), dets.data_ptr(), batch, nt, nd, mode, alpha, out.data_ptr(), _stream_ptr(stream)))
    return out


def greedy_match(cost, threshold, stream=None):
    """cost [batch, R, C] CUDA tensor -> row_matched [batch, R] int32."""
    import torch
    batch, R, Cc = cost.shape
    out = torch.empty(batch, R, device=cost.device, dtype=torch.int32)
    check(lib().pb_greedy_match(cost.data_ptr(), batch, R, Cc, threshold, out.data_ptr(), _stream_ptr(stream)))
    return out


def assign_legacy(cost, threshold, stream=None):
    """cost [batch, R, C] CUDA tensor -> (row [batch, R], col [batch, C], count [batch]) int32 (LinearAssignmentCUDA::solve)."""
    import torch
    batch, R, Cc = cost.shape
    row = torch.empty(batch, R, device=cost.device, dtype=torch.int32)
    col = torch.empty(batch, Cc, device=cost.device, dtype=torch.int32)
    cnt = torch.zeros(batch, device=cost.device, dtype=torch.int32)
    check(lib().pb_assign_legacy(cost.data_ptr(), batch, R, Cc, threshold, row.data_ptr(), col.data_ptr(), cnt.data_ptr(), _stream_ptr(stream)))
    return row, col, cnt


def letterbox_batch(frames, sizes, tw=640, th=640, stream=None):
    """frames [batch, stride] uint8 CUDA tensor (BGR HWC images at the start of each row), sizes [batch, 2] int32
    (width, height) -> (out [batch, 3, th, tw] fp32, xform [batch, 4])."""
    import torch
    batch = frames.shape[0]
    out = torch.empty(batch, 3, th, tw, device=frames.device, dtype=torch.float32)
    xf = torch.empty(batch, 4, device=frames.device, dtype=torch.float32)
    check(lib().pb_letterbox_batch(frames.data_ptr(), frames.stride(0), sizes.data_ptr(), batch, tw, th, out.data_ptr(), xf.data_ptr(),
                                   _stream_ptr(stream)))
    return out, xf


def launch_count() -> int:
    return int(lib().pb_launch_count())


# ---- synthetic input (host generator, lib/libpb_synth.so) -------------------------------
def synth_lib() -> C.CDLL:
    global _synth
    if _synth is None:
        if not os.path.exists(SYNTH_PATH):
            raise FileNotFoundError(f"{SYNTH_PATH} missing: run __graft_entry__.build()")
        S = C.CDLL(SYNTH_PATH)
        S.pb_synth_head.argtypes = [C.POINTER(SynthConfig), C.c_int, C.c_int, C.c_void_p]
        S.pb_synth_heads.argtypes = [C.POINTER(SynthConfig), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_int]
        S.pb_synth_dets.argtypes = [C.POINTER(SynthConfig), C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        S.pb_synth_dets.restype = C.c_int
        _synth = S
    return _synth


def num_anchors_for(canvas: int) -> int:
    return (canvas // 8) ** 2 + (canvas // 16) ** 2 + (canvas // 32) ** 2


def synth_config(canvas=640, persons=20, period=300, clumps=0, occlusion=0, kp_drop_prob=0.05,
                 max_speed=3.0, seed=0x5EEDB200) -> SynthConfig:
    return SynthConfig(num_anchors_for(canvas), canvas, persons, period, clumps, occlusion, kp_drop_prob,
                       max_speed, seed)


def synth_heads(cfg: SynthConfig, stream0: int, nstreams: int, frame0: int, nframes: int,
                frame_major: bool = True, threads: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
    """[F,B,56,N] (frame_major) or [B,F,56,N] fp32 heads."""
    shape = (nframes, nstreams, 56, cfg.num_anchors) if frame_major else (nstreams, nframes, 56, cfg.num_anchors)
    if out is None:
        out = np.empty(shape, dtype=np.float32)
    assert out.shape == shape and out.dtype == np.float32 and out.flags.c_contiguous
    threads = threads or min(os.cpu_count() or 1, 32)
    synth_lib().pb_synth_heads(C.byref(cfg), stream0, nstreams, frame0, nframes, int(frame_major),
                               out.ctypes.data, threads)
    return out


def synth_dets(cfg: SynthConfig, stream: int, frame: int):
    poses = np.zeros((cfg.persons, 51), np.float32)
    scores = np.zeros(cfg.persons, np.float32)
    n = synth_lib().pb_synth_dets(C.byref(cfg), stream, frame, poses.ctypes.data, scores.ctypes.data)
    return poses[:n], scores[:n]


# ---- sharding of streams over ranks (no collective on the hot path) -----------------------
@dataclass
class Shard:
    rank: int
    world: int
    total_streams: int

    @property
    def count(self) -> int:
        base, rem = divmod(self.total_streams, self.world)
        return base + (1 if self.rank < rem else 0)

    @property
    def start(self) -> int:
        base, rem = divmod(self.total_streams, self.world)
        return self.rank * base + min(self.rank, rem)

    def streams(self) -> range:
        return range(self.start, self.start + self.count)


def words_checksum(words_u32: np.ndarray, pos0: int = 0) -> int:
    """Position-weighted 64-bit checksum of 32-bit words: the same function the CPU checker
    applies to its outputs (order sensitive, wraps mod 2^64)."""
    w = np.ascontiguousarray(words_u32).view(np.uint32).astype(np.uint64).ravel()
    idx = np.arange(pos0, pos0 + w.size, dtype=np.uint64)
    with np.errstate(over="ignore"):
        mult = (idx * np.uint64(2) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        return int(((w + np.uint64(0x9E37)) * mult).sum(dtype=np.uint64))
