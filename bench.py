#!/usr/bin/env python
"""bench.py — tracked stream-frames/s of the PoseBYTE post-inference path on N B200s.

One "step" = one pass of the hot path (head decode -> pose-NMS -> tracker update -> TrackOutput
assembly) over one batch of B synthetic stream-frames per GPU.  The headline workload is
BASELINE.json configs[1] (`--config 2`, the default): YOLOv8n-pose 640x640 heads [B,56,8400], 64
concurrent streams per B200, 20 persons per frame, max-age 10.  `--config {1,2,3,4,5a,5b}` selects
any BASELINE.json configuration as the headline; with the default config on one GPU the line also
carries a `configs` array with configurations 1, 3, 4, 5a and 5b measured in the same process
(value, us per batch, roofline fraction, CPU figure, inside the clock-sampling window).  Streams
are sharded over ranks with no data-path collective (weak scaling: B streams per GPU; `--gpus 8
--config 4` is the 1024-stream case, 128 per GPU, max-age 30, occlusion gaps); NCCL only gathers
final statistics.

Contract (see the task statement): W untimed warm-up steps, exactly K timed steps between
barrier + cuda synchronize on both sides, device time from CUDA events, MAX over ranks, rank 0
prints ONE JSON line.  `value` is measured with inputs resident in HBM; `e2e` is the same metric
through the host-buffer entry point with HOST (pinned) buffers, H2D of every step's inputs and D2H
of its track records inside the timed region.  `--impl reference` times the reference side
instead: the reference has no host implementation of this path (its tracker exists only as CUDA
kernels), so the arm runs the scalar C++ transcription of its kernels (oracle/, kind "port",
rebuilt -O3 -march=native on this box) on all host threads, on a bounded sample of the same
workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONF, NMS = 0.30, 0.65

# BASELINE.json configurations.  kind "head": the whole path from a head tensor; kind "tracker": no head,
# tracker.update fed directly (config 5).  B = streams per GPU, F = distinct batches that rotate (working set >> L2).
CONFIGS = {
    "1": dict(kind="head", B=1, canvas=640, persons=20, T=128, Dm=64, max_age=10, F=32,
              workload="yolov8n-pose-640 head [1,56,8400], 1 stream, 20 persons/frame, max-age 10"),
    "2": dict(kind="head", B=64, canvas=640, persons=20, T=128, Dm=64, max_age=10, F=32,
              workload="yolov8n-pose-640 heads [64,56,8400] per GPU, 64 concurrent streams, 20 persons/frame, max-age 10"),
    "3": dict(kind="head", B=64, canvas=1280, persons=100, clumps=10, kp_drop=0.15, T=256, Dm=128, max_age=10, F=4,
              workload="yolo11n-pose-1280 heads [64,56,33600] per GPU, dense crowd of 100 persons/frame in 10 clumps, max_tracks 256 / max_detections 128"),
    "4": dict(kind="head", B=128, canvas=640, persons=20, occlusion=1, T=128, Dm=64, max_age=30, F=16,
              workload="yolov8n-pose-640 heads [128,56,8400] per GPU (1024 streams over 8 GPUs), occlusion gaps, max-age 30 with lost-track recovery"),
    "5a": dict(kind="tracker", B=8, canvas=4096, persons=512, T=512, Dm=512, max_age=10, gating=1, F=8,
               workload="tracker stress: 512 tracks x 512 detections per stream, 8 streams per GPU, OKS cost + auction, spatial gating ON"),
    "5b": dict(kind="tracker", B=8, canvas=4096, persons=512, T=512, Dm=512, max_age=10, gating=0, F=8,
               workload="tracker stress: 512 tracks x 512 detections per stream, 8 streams per GPU, OKS cost + auction, spatial gating OFF"),
}


def num_anchors(canvas):
    return (canvas // 8) ** 2 + (canvas // 16) ** 2 + (canvas // 32) ** 2


def bytes_per_stream_frame(c):
    """algorithmic bytes per tracked stream-frame, dense-read model."""
    if c["kind"] == "head":
        return 224 * num_anchors(c["canvas"]) + 2 * 368 * c["T"] + 228 * c["persons"]
    return 208 * c["Dm"] + 2 * 368 * c["T"] + 228 * c["Dm"]


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram__bytes_read+write per launch from the latest committed ncu --set full capture
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py --traffic); {} if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return {k: int(v) for k, v in json.load(f)["bytes_per_launch"].items()}
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows[1:]:                 # the first sample predates the load
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU side (oracle/ = the scalar C++ transcription of the reference kernels; test infrastructure that
# this file may execute for the cpu_baseline leg and the reference arm only)
# ---------------------------------------------------------------------------------------------------
_orc = None


def oracle():
    """The CPU checker, rebuilt -O3 -march=native for the box it is timed on; the shipped
    portable build is the fallback when no compiler is available."""
    global _orc
    if _orc is None:
        odir = os.path.join(ROOT, "oracle")
        native = os.path.join(odir, "_build", "libposebyte_oracle_native.so")
        flags = "portable build (-O2 -march=x86-64-v3)"
        try:
            r = subprocess.run(["make", "-s", "-C", odir, "native"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=120)
            if r.returncode == 0 and os.path.exists(native):
                os.environ["PB_ORACLE_LIB"] = native
                flags = "g++ -O3 -march=native -ffp-contract=off, built on this box"
        except Exception:
            pass
        sys.path.insert(0, odir)
        import oracle_py
        oracle_py.BUILD_FLAGS = flags
        _orc = oracle_py
    return _orc


def synth_for(pb, c, period=None):
    return pb.synth_config(canvas=c["canvas"], persons=c["persons"], period=period or max(c["F"], 32), clumps=c.get("clumps", 0),
                           occlusion=c.get("occlusion", 0), kp_drop_prob=c.get("kp_drop", 0.05),
                           max_speed=2.0 if c["kind"] == "tracker" else 3.0)


def tracker_dets(pb, c, stream0, B, F):
    """Config 5 input: score-sorted detections [F,B,Dm,51], scores [F,B,Dm], counts [F,B]."""
    scfg = synth_for(pb, c, period=64)
    Dm = c["Dm"]
    poses = np.zeros((F, B, Dm, 51), np.float32); scores = np.zeros((F, B, Dm), np.float32); num = np.zeros((F, B), np.int32)
    for f in range(F):
        for b in range(B):
            p, s = pb.synth_dets(scfg, stream0 + b, f)
            o = np.argsort(-s, kind="stable")[:Dm]
            poses[f, b, : len(o)] = p[o]; scores[f, b, : len(o)] = s[o]; num[f, b] = len(o)
    return poses, scores, num


def cpu_sample(pb, c, threads, frames, streams=None, repeats=1):
    """`streams` (default: one per thread) streams x `frames` frames of configuration c through the port; stream-frames/s."""
    orc = oracle()
    streams = streams or threads
    kw = dict(max_tracks=c["T"], max_detections=c["Dm"], max_age=c["max_age"])
    if c["kind"] == "head":
        heads = pb.synth_heads(synth_for(pb, c), 0, streams, 0, frames, frame_major=False)
        best, stage = None, None
        for _ in range(repeats):
            r = orc.run_streams(heads, False, CONF, NMS, threads=threads, **kw)
            if best is None or r["wall_s"] < best:
                best, stage = r["wall_s"], r["stage_s"]
        return streams * frames / best, stage / (streams * frames) * 1e6
    poses, scores, num = tracker_dets(pb, c, 0, streams, frames)

    def one(b):
        trk = orc.Tracker(gating_enabled=c["gating"], **kw)
        for f in range(frames):
            trk.update(poses[f, b, : num[f, b]], scores[f, b, : num[f, b]], f)
            trk.get_tracks()
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(threads) as ex:          # ctypes releases the GIL: one stream per host thread
            list(ex.map(one, range(streams)))
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return streams * frames / best, None


def cpu_baseline(pb, c, budget="full"):
    """The port on this box's host cores, on a bounded sample of configuration c."""
    cores = os.cpu_count() or 1
    head = c["kind"] == "head"
    big = head and c["canvas"] > 640
    frames_all = (8 if big else 48) if head else 6
    if budget == "short":
        frames_all = (4 if big else 16) if head else 4
    v_all, _ = cpu_sample(pb, c, cores, frames_all, repeats=2 if budget == "full" else 1)
    out = {"value": v_all, "unit": "stream-frames/s", "cores": cores, "kind": "port",
           "sample": f"{cores} streams x {frames_all} frames of this workload, one stream per host thread; scalar C++ transcription of the reference "
                     f"kernels ({oracle().BUILD_FLAGS})"}
    if budget == "full":
        frames_1 = (300 if not big else 30) if head else 12       # config 1 as SURVEY.md 8(d) specifies it: one stream, 300 frames, one core
        v1, st = cpu_sample(pb, c, 1, frames_1, streams=1)
        out["single_core"] = {"value": v1, "sample": f"1 stream x {frames_1} frames on one core"}
        if st is not None:
            out["single_core"]["us_per_frame"] = {"decode": float(st[0]), "nms": float(st[1]), "track": float(st[2])}
    return out


def reference_host_nms(pb, c):
    """NMSCuda::apply COMPILED FROM THE REFERENCE's nms.cu (host code, oracle/_ref) on the candidates of this workload's
    frames: us per frame on one core (BASELINE.md 'CPU-NMS-ref'; reference src/cuda/nms.cu:142-306, src/benchmark.cpp:177-205)."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_py
        if not ref_py.available():
            return None
        orc = oracle()
        heads = pb.synth_heads(synth_for(pb, c), 0, 1, 0, 16, frame_major=False)[0]
        dets = []
        for f in range(16):
            d = orc.decode(heads[f], CONF)
            a = np.zeros(d["num"], ref_py.POSE_DETECTION)
            a["bbox"] = d["bboxes"][: d["num"]]; a["score"] = d["scores"][: d["num"]]; a["keypoints"] = d["poses"][: d["num"]].reshape(-1, 17, 3)
            dets.append(a)
        for a in dets[:2]:
            ref_py.nms_apply(a, 0.65, 0.25)
        t0 = time.perf_counter()
        kept = sum(len(ref_py.nms_apply(a, 0.65, 0.25)) for a in dets)
        dt = time.perf_counter() - t0
        return {"us_per_frame": dt / len(dets) * 1e6, "candidates_per_frame": float(np.mean([len(a) for a in dets])), "kept_per_frame": kept / len(dets),
                "cores": 1, "kind": "reference", "what": "NMSCuda::apply compiled from the reference's src/cuda/nms.cu (host-legacy rule set), 16 frames of this workload"}
    except Exception as e:  # pragma: no cover
        return {"error": str(e)}


def reference_gpu_b1(pb, torch, c, frames=200):
    """The reference's own .cu files compiled unchanged for sm_100a (oracle/_ref), B=1: ms/frame of
    process() + update() (+ getActiveTracks), the launch-bound baseline on the same box."""
    import ctypes as C
    path = os.path.join(ROOT, "oracle", "_ref", "libposebyte_ref.so")
    if not os.path.exists(path) or c["kind"] != "head":
        return None
    try:
        R = C.CDLL(path)
        R.ref_time_frames.restype = C.c_double
        R.ref_time_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
        period = 32
        heads = torch.from_numpy(pb.synth_heads(synth_for(pb, c), 0, 1, 0, period, frame_major=False)[0]).cuda()
        devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)   # the ctor prints a banner
        try:
            n = num_anchors(c["canvas"])
            ms = R.ref_time_frames(heads.data_ptr(), period, n, 50, frames, CONF, NMS, c["T"], c["Dm"], c["max_age"], 0)
            ms_rb = R.ref_time_frames(heads.data_ptr(), period, n, 50, frames, CONF, NMS, c["T"], c["Dm"], c["max_age"], 1)
        finally:
            os.dup2(saved, 1); os.close(devnull); os.close(saved)
        return {"ms_per_frame": ms, "ms_per_frame_with_readback": ms_rb, "stream_frames_per_s": 1000.0 / ms_rb,
                "what": "reference src/cuda/*.cu compiled for sm_100a, 1 stream, GPUPostprocess::process + GPUTracker::update (+ getActiveTracks)"}
    except Exception as e:  # pragma: no cover
        return {"error": str(e)}


def run_reference_arm(args, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import posebyte_b200 as pb
    c = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    head = c["kind"] == "head"
    frames = (32 if c["canvas"] <= 640 else 4) if head else 4
    orc = oracle()
    kw = dict(max_tracks=c["T"], max_detections=c["Dm"], max_age=c["max_age"])
    # one "step" of this arm = one bounded sample: `cores` streams x `frames` frames through the port
    if head:
        heads = pb.synth_heads(synth_for(pb, c), 0, cores, 0, frames, frame_major=False)

        def step():
            orc.run_streams(heads, False, CONF, NMS, threads=cores, **kw)
    else:
        poses, scores, num = tracker_dets(pb, c, 0, cores, frames)

        def one(b):
            trk = orc.Tracker(gating_enabled=c["gating"], **kw)
            for f in range(frames):
                trk.update(poses[f, b, : num[f, b]], scores[f, b, : num[f, b]], f)
                trk.get_tracks()

        def step():
            with ThreadPoolExecutor(cores) as ex:
                list(ex.map(one, range(cores)))
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = cores * frames * args.steps / dt
    sample = (f"{cores} streams x {frames} frames per step; scalar C++ transcription of the reference kernels ({orc.BUILD_FLAGS}; the reference has no "
              "host implementation of the tracker)")
    line = {"impl": "reference", "metric": "tracked stream-frames/sec", "value": value, "unit": "stream-frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": c["workload"], "config_id": args.config,
                       "step": f"{cores} streams x {frames} frames (bounded sample) on {cores} host threads"},
            "cpu_baseline": {"value": value, "unit": "stream-frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "stream-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    out.emit(json.dumps(line))


class OnlyJsonOnStdout:
    """Everything libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr; the one
    JSON line is written to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line: str):
        sys.stdout.flush()
        os.write(self.real, (line + "\n").encode())


# ---------------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------------
class Workload:
    """Device-resident inputs of one configuration on one rank + its handle."""

    def __init__(self, pb, torch, cid, stream0, dev, depth, B=None):
        self.pb, self.torch, self.cid, self.c = pb, torch, cid, dict(CONFIGS[cid])
        c = self.c
        if B:
            c["B"] = B
        self.B, self.F, self.dev = c["B"], c["F"], dev
        self.N = num_anchors(c["canvas"]) if c["kind"] == "head" else 64
        kw = dict(num_streams=self.B, max_tracks=c["T"], max_detections=c["Dm"], max_age=c["max_age"], device=dev.index)
        if c["kind"] == "head":
            self.host = pb.synth_heads(synth_for(pb, c), stream0, self.B, 0, self.F, frame_major=True)       # [F,B,56,N]
            self.d_heads = torch.from_numpy(self.host).to(dev)
            self.pipe = pb.Pipeline(num_anchors=self.N, pipeline_depth=depth, **kw)
        else:
            self.h_poses, self.h_scores, self.h_num = tracker_dets(pb, c, stream0, self.B, self.F)
            self.d_poses, self.d_scores = torch.from_numpy(self.h_poses).to(dev), torch.from_numpy(self.h_scores).to(dev)
            self.d_num = torch.from_numpy(self.h_num).to(dev)
            self.pipe = pb.Pipeline(num_anchors=64, max_candidates=64, max_keep=64, gating_enabled=c["gating"], **kw)
        self.f = 0

    def step(self):
        f, i = self.f, self.f % self.F
        if self.c["kind"] == "head":
            self.pipe.step(self.d_heads[i], f, CONF, NMS)
        else:
            self.pipe.tracker_update(f, self.d_poses[i], self.d_scores[i], self.d_num[i], self.c["Dm"])
        self.f += 1

    def run(self, n):
        if self.c["kind"] == "head":          # n steps through one C call (pb_step_seq == n x pb_step without n ctypes crossings)
            self.pipe.step_seq(self.d_heads, self.f % self.F, n, self.f, CONF, NMS)
            self.f += n
        else:
            for _ in range(n):
                self.step()
        self.pipe.join()

    def timed(self, steps, barrier, stream):
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        t0 = time.perf_counter()
        self.run(steps)
        self.host_enqueue_s = time.perf_counter() - t0      # host time to enqueue the K steps (the device runs behind it)
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1)

    def kernel_us(self, n):
        """Serial path (one stream), every launch bracketed by CUDA events on the launching stream."""
        self.pipe.set_profiling(True)
        for _ in range(3):                     # the serial path's kernels may not have run yet (pb_step_seq takes other kernels: the
            self.step()                        # first launch of a kernel loads its code): warm-up launches, discarded
        self.pipe.kernel_us()
        for _ in range(n):
            self.step()
        k = self.pipe.kernel_us()
        self.pipe.set_profiling(False)
        return k

    def latency_us(self, stream, n=30):
        torch, lat = self.torch, []
        for _ in range(n):
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            self.step(); self.pipe.join()
            b_.record(stream)
            torch.cuda.synchronize()
            lat.append(a.elapsed_time(b_) * 1e3)
        return float(np.median(lat))

    def e2e(self, steps, barrier):
        """The host-buffer entry points: page-locked inputs, results copied back into page-locked memory, host reads every result."""
        torch, pb, pipe, B, c = self.torch, self.pb, self.pipe, self.B, self.c
        Dm = c["Dm"]
        R = 4
        out_p = torch.zeros(R, B * Dm * 228, dtype=torch.uint8).pin_memory()
        cnt_p = torch.zeros(R, B, dtype=torch.int32).pin_memory()
        if c["kind"] == "head":
            nf = min(self.F, 8)
            pinned_np = torch.from_numpy(self.host[:nf]).pin_memory().numpy()
            out_np = [out_p[i].numpy().view(pb.TRACK_OUTPUT).reshape(B, Dm) for i in range(R)]
            cnt_np = [cnt_p[i].numpy() for i in range(R)]

            def run(n, f0):
                tracks = 0
                for i in range(n):
                    pipe.submit_host(pinned_np[i % nf], f0 + i, out_np[i % R], cnt_np[i % R], CONF, NMS)
                    if (i + 1) % R == 0 or i == n - 1:
                        pipe.wait()
                        tracks += int(sum(int(x.sum()) for x in cnt_np[: (i % R) + 1]))
                return tracks
            h2d = None
        else:
            nf = self.F
            pp, ps, pn = (torch.from_numpy(x).pin_memory() for x in (self.h_poses, self.h_scores, self.h_num))
            views = pipe.device_views()
            s = torch.cuda.Stream(device=self.dev)
            dp = [torch.empty_like(self.d_poses[0]) for _ in range(2)]; ds = [torch.empty_like(self.d_scores[0]) for _ in range(2)]
            dn = [torch.empty_like(self.d_num[0]) for _ in range(2)]
            import ctypes as C
            rt = C.CDLL("libcudart.so.12") if os.path.exists("/usr/local/cuda/lib64/libcudart.so.12") else C.CDLL("libcudart.so")
            rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]

            def run(n, f0):
                tracks = 0
                with torch.cuda.stream(s):
                    for i in range(n):
                        k, j = i % nf, i % 2
                        dp[j].copy_(pp[k], non_blocking=True); ds[j].copy_(ps[k], non_blocking=True); dn[j].copy_(pn[k], non_blocking=True)
                        pipe.tracker_update(f0 + i, dp[j], ds[j], dn[j], Dm, stream=s)
                        rt.cudaMemcpyAsync(cnt_p[i % R].data_ptr(), views.num_outputs, B * 4, 2, s.cuda_stream)
                        rt.cudaMemcpyAsync(out_p[i % R].data_ptr(), views.track_outputs, B * Dm * 228, 2, s.cuda_stream)
                        if (i + 1) % 2 == 0 or i == n - 1:
                            s.synchronize()
                            tracks += int(cnt_p[: (i % R) + 1].sum())
                return tracks
            h2d = B * Dm * 52 * 4 + B * 4
        run(4, self.f); self.f += 4
        barrier()
        paths0 = pipe.nms_path_counts() if c["kind"] == "head" else None
        t0 = time.perf_counter()
        tracks = run(steps, self.f); self.f += steps
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        info = {"seconds": dt, "steps": steps, "tracks": tracks, "d2h_bytes_per_step": B * Dm * 228 + B * 4}
        if c["kind"] == "head":
            paths1 = pipe.nms_path_counts()
            kp = (paths1["keypoint_fetches"] - paths0["keypoint_fetches"]) / max(B * steps, 1)
            cand = float(np.mean([pipe.get_kept(b)["num_cand"] for b in range(min(B, 8))]))
            info.update(keypoint_fetches_per_stream_frame=kp, candidates=cand, host_input_bytes_per_step=B * 56 * self.N * 4,
                        h2d_bytes_per_step=int(B * (4 * self.N + 4 * 32 * cand + 51 * 32 * kp)))
        else:
            info.update(h2d_bytes_per_step=h2d, host_input_bytes_per_step=h2d)
        return info

    def close(self):
        self.pipe.close()
        for k in ("d_heads", "host", "d_poses", "d_scores", "d_num", "h_poses", "h_scores", "h_num"):
            if hasattr(self, k):
                delattr(self, k)
        self.torch.cuda.empty_cache()


def kf3_line(pb, torch, peak):
    """3rd-order Kalman filter (north-star list; reference src/cuda/kalman_filter.cu:422-456, micro-benchmark
    src/benchmark.cpp:68-98): predict + update over T tracks, 1 088 B per track read and 1 088 B written per kernel
    (136 means + 136 covariance-diagonal floats), against the measured HBM bandwidth."""
    L = pb.lib()
    st = torch.cuda.current_stream().cuda_stream
    Tn = 1 << 20                                                     # 1 Mi tracks: 1.14 GB of state, far beyond L2
    rng = np.random.default_rng(0)
    means = torch.zeros(Tn, 136, device="cuda"); diag = torch.zeros(Tn, 136, device="cuda")
    dets = torch.from_numpy(rng.uniform(0, 1080, (4096, 51)).astype(np.float32)).cuda().repeat(Tn // 4096, 1)
    slots = torch.arange(Tn, dtype=torch.int32, device="cuda")
    matches = torch.stack([slots, slots], 1).contiguous()
    pb.check(L.pb_kf3_initiate(means.data_ptr(), diag.data_ptr(), dets.data_ptr(), slots.data_ptr(), Tn, st))
    res = {}
    for name, fn, nbytes in (("predict", lambda: L.pb_kf3_predict(means.data_ptr(), diag.data_ptr(), Tn, 0.9, 0.9, st), Tn * 2 * 1088),
                             ("update", lambda: L.pb_kf3_update(means.data_ptr(), diag.data_ptr(), dets.data_ptr(), matches.data_ptr(), Tn, st),
                              Tn * (2 * 1088 + 204 + 8))):
        for _ in range(3):
            pb.check(fn())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            pb.check(fn())
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 10 * 1e3
        res[name] = {"us": us, "tracks_per_s": Tn / us * 1e6, "GB/s": nbytes / us / 1e3, "frac": nbytes / us / 1e3 / peak}
    res["tracks"] = Tn
    res["note"] = "compressed state (mean + covariance diagonal, 1 088 B per track; the reference's 136x136 matrix is identically zero off the diagonal)"
    return res


def main():
    out = OnlyJsonOnStdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="2", choices=sorted(CONFIGS))
    ap.add_argument("--streams-per-gpu", type=int, default=0, help="override the configuration's streams per GPU")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--pipeline-depth", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the `configs` array (other BASELINE.json configurations)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    if args.impl == "reference":
        run_reference_arm(args, out)
        return

    import torch
    import torch.distributed as dist
    import posebyte_b200 as pb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # Rendezvous barriers run over gloo (host side); the NCCL communicator is created for the statistics gather at the end — the run's
    # only collective — so that no NCCL proxy / watchdog thread shares the host cores and the CUDA context with the thread that
    # enqueues the timed steps.  BENCH_EARLY_NCCL=1: NCCL from the start (barriers included), for comparison.
    early_nccl = os.environ.get("BENCH_EARLY_NCCL") == "1"
    gloo = None
    if world > 1:
        if early_nccl:
            dist.init_process_group("nccl", device_id=dev)
        else:
            dist.init_process_group("gloo")
            gloo = dist.group.WORLD

    def barrier():
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
        torch.cuda.synchronize()

    cid = args.config
    c = CONFIGS[cid]
    B = args.streams_per_gpu or c["B"]
    shard = pb.Shard(rank, world, B * world)
    stream = torch.cuda.current_stream()
    peak, peak_src = measured_peak_gbs()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_load0 = time.perf_counter()
    w = Workload(pb, torch, cid, shard.start, dev, args.pipeline_depth, B=B)

    # ---------------- device-resident throughput (`value`) ----------------
    # K steps issued back to back; with pipeline_depth > 1 the library overlaps consecutive steps on
    # internal streams, pb_join makes the timing stream wait for all of it before the closing event.
    w.run(args.warmup)
    barrier()
    launches0 = pb.launch_count()
    ms_total = w.timed(args.steps, barrier, stream)
    host_enqueue_us = w.host_enqueue_s / args.steps * 1e6
    launches = pb.launch_count() - launches0
    outs, counts = w.pipe.get_tracks_all()
    n_out_mean = float(counts.mean())
    kus = w.kernel_us(min(max(args.steps, 20), 100))
    latency_us = w.latency_us(stream)
    e2e = w.e2e(args.e2e_steps, barrier)
    cand_mean = e2e.get("candidates", 0.0)

    # ---------------- the other BASELINE.json configurations, same process (one GPU, default config) ----------------
    extra = []
    if world == 1 and cid == "2" and not args.no_extra_configs:
        w.close()
        xsteps = min(max(args.steps, 100), 200)         # the other configurations: steady-state figures (the headline keeps the caller's K)
        for xid in ("1", "3", "4", "5a", "5b"):
            xc = CONFIGS[xid]
            xw = Workload(pb, torch, xid, 0, dev, args.pipeline_depth)
            xw.run(max(args.warmup, 8))
            torch.cuda.synchronize()
            xms = xw.timed(xsteps, barrier, stream)
            xk = xw.kernel_us(min(xsteps, 40))
            xlat = xw.latency_us(stream, 10)
            xe = xw.e2e(min(args.e2e_steps, 40 if xc["kind"] == "head" and xc["canvas"] <= 640 else 12), barrier)
            _, xcounts = xw.pipe.get_tracks_all()
            bsf = bytes_per_stream_frame(xc)
            v = xw.B * xsteps / (xms / 1e3)
            rec = {"config_id": xid, "workload": xc["workload"], "value": v, "unit": "stream-frames/s", "steps": xsteps,
                   "us_per_batch": xms / xsteps * 1e3, "us_per_stream_frame": xms / xsteps * 1e3 / xw.B, "us_per_batch_latency": xlat,
                   "streams_per_gpu": xw.B, "step_path": xw.pipe.step_path(), "kernel_us": {k: v_ for k, v_ in xk.items() if k != "launches"},
                   "roofline_step": {"bound": "hbm", "achieved": v * bsf / 1e9, "peak": peak, "unit": "GB/s", "frac": v * bsf / 1e9 / peak,
                                     "bytes_per_stream_frame": bsf},
                   "e2e": {"value": xw.B * xe["steps"] / xe["seconds"], "unit": "stream-frames/s", "steps": xe["steps"],
                           "h2d_bytes_per_step": xe["h2d_bytes_per_step"], "d2h_bytes_per_step": xe["d2h_bytes_per_step"]},
                   "tracks_per_stream_frame": float(xcounts.mean())}
            if xc["kind"] == "tracker":
                rec["tracker_stage_us"] = xw.pipe.tracker_stage_us()
                rec["work"] = "3 tiers x up to 50 iterations x 2*T*D = 78.6 M cell visits per stream-frame if no tier converges early (SURVEY.md 8d)"
            xw.close()
            if not args.no_cpu_baseline:
                rec["cpu_baseline"] = cpu_baseline(pb, xc, budget="short")
            extra.append(rec)
        # the clock record has to cover the headline and the extra configurations: keep the GPU under this load to the end
        w = Workload(pb, torch, cid, shard.start, dev, args.pipeline_depth, B=B)
        w.run(args.warmup)
    # nvidia-smi samples every 50 ms; a short timed region is followed by a continuation of the
    # same loop so that the clock record covers at least ~0.6 s of this load
    while rank == 0 and time.perf_counter() - t_load0 < 0.6:
        w.run(50); torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = ("data generation + warm-up + timed region + kernel/latency/e2e legs" + (" + the `configs` array" if extra else "") +
                            ", continued with the same step loop to >= 0.6 s")

    # ---------------- max over ranks + final statistics gather (the only collective) ----------------
    # per-GPU result counters travel with the timings: tracks emitted, active tracks, a hash of the final records
    na = w.pipe.get_num_active()
    h64 = pb.words_checksum(np.frombuffer(outs.tobytes(), np.uint32)[: 1 << 16]) & ((1 << 52) - 1)
    stats = torch.tensor([ms_total, e2e["seconds"], float(launches), float(counts.sum()), kus["gather_us"], kus["nms_us"], kus["track_us"],
                          float(e2e["tracks"]), latency_us, float(na.sum()), float(h64), host_enqueue_us], dtype=torch.float64, device=dev)
    if world > 1:
        nccl = None if early_nccl else dist.new_group(backend="nccl", device_id=dev)      # one rank per GPU over NCCL
        gathered = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats, group=nccl)
        allstats = torch.stack(gathered).cpu().numpy()
    else:
        allstats = stats.cpu().numpy()[None]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_max, e2e_max = float(allstats[:, 0].max()), float(allstats[:, 1].max())
    total_streams = B * world
    value = total_streams * args.steps / (ms_max / 1e3)
    e2e_value = total_streams * e2e["steps"] / e2e_max
    gather_us, nms_us, track_us = (float(allstats[:, i].max()) for i in (4, 5, 6))

    bsf = bytes_per_stream_frame(c)
    traffic = ncu_traffic()

    def roof(name, bytes_per_launch, us, note):
        ach = bytes_per_launch / (us / 1e6) / 1e9 if us > 0 else 0.0
        tkey = name + "[512x512]" if c["kind"] != "head" and name == "pb_tracker_kernel" else name      # (ncu capture of the matching table size)
        return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic.get(tkey), "avg_launch_us": us, "algorithmic_bytes_per_launch": bytes_per_launch,
                "peak_source": peak_src, "note": note}
    kernels = []
    if c["kind"] == "head":
        bh = 224 * w.N
        kernels.append(roof("pb_decode_gather_kernel", B * bh, gather_us,
                            "dense-read model 224*N B per stream-frame (SURVEY.md 8d); the kernel reads the confidence row and 32 B sectors at "
                            "candidate anchors only, so this is an EFFECTIVE bandwidth; `traffic` is ncu dram bytes per launch"))
        npl = w.pipe.nms_plan()
        kernels.append(roof("pb_nms_kernel", B * int(cand_mean) * 224, nms_us,
                            "candidate records (224 B each) read once; latency/issue-bound, working set in shared memory; launched as "
                            + ("pb_nms_tier_kernel<%d,%d> with %d B of shared memory (working set of %d candidates, spill path beyond: the "
                               "shared-memory configuration of the tracker kernel)" % (npl["threads"], npl["ctas_per_sm"], npl["smem_bytes"], npl["tier_candidates"])
                               if npl["tier_candidates"] < w.pipe.cfg.max_candidates else "pb_nms_kernel (full candidate cap in shared memory)")))
    kernels.append(roof("pb_tracker_kernel", B * (bsf - (224 * w.N if c["kind"] == "head" else 0)), track_us,
                        "track state read+written once + TrackOutput records (+ detections for config 5); latency-bound (up to 150 dependent auction "
                        "iterations per stream-frame), HBM is not its limit"))
    # `roofline`: the dominant kernel (longest launch) as the task statement defines it — SURVEY.md 8(d)'s per-unit figure (the
    # algorithmic bytes of a tracked stream-frame through the WHOLE path) x the stream-frames one launch processes, over that
    # kernel's own launch duration.  The kernel's own bytes and bandwidth are its entry in `roofline_kernels`; the step-level
    # figure (same bytes over the measured step time) is `roofline_step`.
    dk = max(kernels, key=lambda r: r["avg_launch_us"])
    dom_ach = B * bsf / (dk["avg_launch_us"] / 1e6) / 1e9 if dk["avg_launch_us"] > 0 else 0.0
    dominant = {"kernel": dk["kernel"], "bound": "hbm", "achieved": dom_ach, "peak": peak, "unit": "GB/s", "frac": dom_ach / peak,
                "traffic": dk["traffic"], "avg_launch_us": dk["avg_launch_us"], "algorithmic_bytes_per_launch": B * bsf,
                "bytes_per_stream_frame": bsf, "stream_frames_per_launch": B, "peak_source": peak_src,
                "own_algorithmic_bytes_per_launch": dk["algorithmic_bytes_per_launch"], "own_achieved": dk["achieved"],
                "note": "SURVEY.md 8(d) bytes per stream-frame (dense-read model of the whole path) x streams per launch over the duration of the "
                        "longest kernel of the step (CUDA events, serial path); this kernel itself moves `own_algorithmic_bytes_per_launch` "
                        "(`traffic`: ncu dram bytes per launch) and is bound by the latency of its dependent stages, not by HBM — see "
                        "`roofline_kernels` and `roofline_step`"}
    step_ach = total_streams * bsf * args.steps / (ms_max / 1e3) / 1e9 / world
    per_rank_ms = [float(x) / args.steps for x in allstats[:, 0]]
    line = {
        "metric": "tracked stream-frames/sec", "value": value, "unit": "stream-frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "us_per_batch": ms_max / args.steps * 1e3,
        "us_per_batch_latency": float(allstats[:, 8].max()),
        "ms_per_step_ranks": {"min": min(per_rank_ms), "median": float(np.median(per_rank_ms)), "max": max(per_rank_ms)},
        "host_enqueue_us_per_step_ranks": {"min": float(allstats[:, 11].min()), "median": float(np.median(allstats[:, 11])), "max": float(allstats[:, 11].max())},
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": c["workload"], "config_id": cid, "streams_per_gpu": B, "total_streams": total_streams,
                   "parallelism": f"stream-sharded x{world}, no data-path collective",
                   "l2": f"{w.F} distinct input batches rotate ({w.F * B * (56 * w.N * 4 if c['kind'] == 'head' else c['Dm'] * 208) / 1e9:.2f} GB per GPU"
                         + (" > 126 MB L2)" if c["kind"] == "head" else "; the tracker's own state is the working set)"),
                   "conf": CONF, "nms": NMS, "max_tracks": c["T"], "max_detections": c["Dm"], "max_age": c["max_age"],
                   "pipeline_depth": args.pipeline_depth, "step_path": w.pipe.step_path(),
                   "timing": "value: K steps back to back + pb_join between CUDA events (steps overlap on internal streams); "
                             "us_per_batch_latency: median of single isolated steps; roofline: per-launch CUDA events on the serial path"},
        "roofline": dominant,
        "roofline_kernels": kernels,
        "roofline_step": {"bound": "hbm", "achieved": step_ach, "peak": peak, "unit": "GB/s", "frac": step_ach / peak,
                          "bytes_per_stream_frame": bsf, "note": "whole step per GPU, SURVEY.md 8(d) dense-read figure"},
        "e2e": {"value": e2e_value, "unit": "stream-frames/s", "h2d_bytes_per_step": e2e["h2d_bytes_per_step"],
                "keypoint_fetches_per_stream_frame": e2e.get("keypoint_fetches_per_stream_frame"),
                "host_input_bytes_per_step": e2e["host_input_bytes_per_step"], "d2h_bytes_per_step": e2e["d2h_bytes_per_step"], "steps": e2e["steps"],
                "api": ("pb_submit_host / pb_wait (results read on the host every 4 steps): page-locked host heads [B,56,N] read in place over PCIe — "
                        "confidence rows, 32 B sectors of the 4 box rows at every candidate anchor and of the 51 keypoint rows at the candidates the "
                        "lazy NMS sweep has to test (= h2d_bytes_per_step; the buffer itself is host_input_bytes_per_step); TrackOutput records "
                        "copied back into page-locked memory") if c["kind"] == "head" else
                       "page-locked detections copied to the device, pb_tracker_update, TrackOutput records copied back into page-locked memory, "
                       "host reads the counts every 2 steps"},
        "gpu_launches": int(allstats[0, 2]),
        "clocks": clocks,
        "tracks_per_stream_frame": n_out_mean,
        "candidates_per_stream_frame": cand_mean,
        "gathered_counters": {"tracks_emitted_last_step": [int(x) for x in allstats[:, 3]], "active_tracks": [int(x) for x in allstats[:, 9]],
                              "records_hash52": [int(x) for x in allstats[:, 10]], "e2e_tracks": [int(x) for x in allstats[:, 7]],
                              "note": "one all_gather of this vector is the run's only collective"},
    }
    if extra:
        line["configs"] = extra
    if world == 1 and not args.no_cpu_baseline:
        w.close()
        cb = cpu_baseline(pb, c)
        ref_gpu = reference_gpu_b1(pb, torch, c)
        if ref_gpu:
            cb["reference_gpu_b1"] = ref_gpu
        if c["kind"] == "head":
            rn = reference_host_nms(pb, c)
            if rn:
                cb["reference_host_nms"] = rn
        line["cpu_baseline"] = cb
        try:
            line["kf3"] = kf3_line(pb, torch, peak)
        except Exception as e:  # pragma: no cover
            line["kf3"] = {"error": str(e)}
    out.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
