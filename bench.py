#!/usr/bin/env python
"""bench.py — tracked stream-frames/s of the PoseBYTE post-inference path on N B200s.

One "step" = one pass of the hot path (head decode -> pose-NMS -> tracker update -> TrackOutput
assembly) over one batch of B synthetic stream-frames per GPU.  Workload = BASELINE.json
configs[1]: YOLOv8n-pose 640x640 heads [B,56,8400], 64 concurrent streams per B200, 20 persons
per frame, max-age 10.  Streams are sharded over ranks with no data-path collective (weak
scaling: 64 streams per GPU); NCCL only gathers final statistics.

Contract (see the task statement): W untimed warm-up steps, exactly K timed steps between
barrier + cuda synchronize on both sides, device time from CUDA events, MAX over ranks, rank 0
prints ONE JSON line.  `value` is measured with inputs resident in HBM; `e2e` is the same metric
through pb_step_host with HOST (pinned) buffers, H2D of every step's heads and D2H of its track
records inside the timed region.  `--impl reference` times the reference side instead: the
reference has no host implementation of this path (its tracker exists only as CUDA kernels), so
the arm runs the scalar C++ transcription of its kernels (oracle/, kind "port") on all host
threads, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = "yolov8n-pose-640 heads [64,56,8400] per GPU, 64 concurrent streams, 20 persons/frame, max-age 10"
STREAMS_PER_GPU = 64
CANVAS, PERSONS, PERIOD = 640, 20, 32          # 32 distinct frames per stream, periodic motion
CONF, NMS = 0.30, 0.65
T, DM, MAX_AGE = 128, 64, 10
N_ANCHORS = 8400
# algorithmic bytes per tracked stream-frame (dense-read model)
BYTES_HEAD = 224 * N_ANCHORS                   # [56,N] fp32 read once
BYTES_TRACK = 2 * 368 * T                      # track state read + written
BYTES_OUT = 228 * PERSONS                      # TrackOutput records
BYTES_PER_STREAM_FRAME = BYTES_HEAD + BYTES_TRACK + BYTES_OUT   # 1 980 368


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


def cpu_baseline(pb, sample_frames=32, repeats=2, threads=None):
    """The scalar C++ transcription of the reference kernels (oracle/, kind 'port') on the host
    cores of this box, on a bounded sample of the bench workload: one stream per thread."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as orc
    cores = threads or (os.cpu_count() or 1)
    scfg = pb.synth_config(canvas=CANVAS, persons=PERSONS, period=PERIOD)
    nstreams = cores
    heads = pb.synth_heads(scfg, 0, nstreams, 0, sample_frames, frame_major=False)
    # single core first (the reference is single-threaded per stream)
    r1 = orc.run_streams(heads[:1], False, CONF, NMS, threads=1, max_tracks=T, max_detections=DM, max_age=MAX_AGE)
    best = None
    for _ in range(repeats):
        r = orc.run_streams(heads, False, CONF, NMS, threads=cores, max_tracks=T, max_detections=DM, max_age=MAX_AGE)
        if best is None or r["wall_s"] < best["wall_s"]:
            best = r
    sf = nstreams * sample_frames
    st = r1["stage_s"] / sample_frames * 1e6
    return {"value": sf / best["wall_s"], "unit": "stream-frames/s", "cores": cores, "kind": "port",
            "sample": f"{nstreams} streams x {sample_frames} frames of the bench workload, one stream per thread, best of {repeats}",
            "single_core": {"value": sample_frames / r1["wall_s"], "us_per_frame": {"decode": float(st[0]), "nms": float(st[1]), "track": float(st[2])}}}


def reference_gpu_b1(pb, torch, frames=200):
    """The reference's own .cu files compiled unchanged for sm_100a (oracle/_ref), B=1: ms/frame of
    process() + update() (+ getActiveTracks), the launch-bound baseline on the same box."""
    import ctypes as C
    path = os.path.join(ROOT, "oracle", "_ref", "libposebyte_ref.so")
    if not os.path.exists(path):
        return None
    try:
        R = C.CDLL(path)
        R.ref_time_frames.restype = C.c_double
        R.ref_time_frames.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]
        scfg = pb.synth_config(canvas=CANVAS, persons=PERSONS, period=PERIOD)
        heads = torch.from_numpy(pb.synth_heads(scfg, 0, 1, 0, PERIOD, frame_major=False)[0]).cuda()
        devnull = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(devnull, 1)   # the ctor prints a banner
        try:
            ms = R.ref_time_frames(heads.data_ptr(), PERIOD, N_ANCHORS, 50, frames, CONF, NMS, T, DM, MAX_AGE, 0)
            ms_rb = R.ref_time_frames(heads.data_ptr(), PERIOD, N_ANCHORS, 50, frames, CONF, NMS, T, DM, MAX_AGE, 1)
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
    cores = os.cpu_count() or 1
    sample_frames = 32
    # one "step" of this arm = one bounded sample: `cores` streams x 32 frames through the port
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as orc
    scfg = pb.synth_config(canvas=CANVAS, persons=PERSONS, period=PERIOD)
    heads = pb.synth_heads(scfg, 0, cores, 0, sample_frames, frame_major=False)
    kw = dict(threads=cores, max_tracks=T, max_detections=DM, max_age=MAX_AGE)
    for _ in range(args.warmup):
        orc.run_streams(heads, False, CONF, NMS, **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.run_streams(heads, False, CONF, NMS, **kw)
    dt = time.perf_counter() - t0
    sf = cores * sample_frames * args.steps
    value = sf / dt
    line = {"impl": "reference", "metric": "tracked stream-frames/sec", "value": value, "unit": "stream-frames/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "step": f"{cores} streams x {sample_frames} frames (bounded sample) on {cores} host threads"},
            "cpu_baseline": {"value": value, "unit": "stream-frames/s", "cores": cores, "kind": "port",
                             "sample": f"{cores} streams x {sample_frames} frames per step; scalar C++ transcription of the reference kernels "
                                       "(the reference has no host implementation of the tracker)"},
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


def main():
    out = OnlyJsonOnStdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams-per-gpu", type=int, default=STREAMS_PER_GPU)
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--pipeline-depth", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.streams_per_gpu
    shard = pb.Shard(rank, world, B * world)
    scfg = pb.synth_config(canvas=CANVAS, persons=PERSONS, period=PERIOD)
    host_heads = pb.synth_heads(scfg, shard.start, B, 0, PERIOD, frame_major=True)       # [F,B,56,N]
    d_heads = torch.from_numpy(host_heads).to(dev)
    kw = dict(num_streams=B, num_anchors=N_ANCHORS, max_tracks=T, max_detections=DM, max_age=MAX_AGE, device=local)
    pipe = pb.Pipeline(pipeline_depth=args.pipeline_depth, **kw)
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`) ----------------
    # K steps issued back to back through pb_step; with pipeline_depth > 1 the library overlaps the
    # decode+gather / NMS / tracker kernels of consecutive steps on internal streams, pb_join makes
    # the timing stream wait for all of it before the closing event.
    f = 0
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_load0 = time.perf_counter()
    for _ in range(args.warmup):
        pipe.step(d_heads[f % PERIOD], f, CONF, NMS); f += 1
    pipe.join()
    barrier()
    launches0 = pb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        pipe.step(d_heads[f % PERIOD], f, CONF, NMS); f += 1
    pipe.join()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = pb.launch_count() - launches0
    # nvidia-smi samples every 50 ms; a short timed region is followed by a continuation of the
    # same loop so that the clock record covers at least ~0.6 s of this load
    while rank == 0 and time.perf_counter() - t_load0 < 0.6:
        for _ in range(50):
            pipe.step(d_heads[f % PERIOD], f, CONF, NMS); f += 1
        pipe.join(); torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed region + continuation of the same step loop to >= 0.6 s"
    outs, counts = pipe.get_tracks_all()
    n_out_mean = float(counts.mean())

    # ---------------- per-kernel device time (roofline) and single-step latency ----------------
    # serial path (one stream), every launch bracketed by CUDA events on the launching stream
    pipe.set_profiling(True)
    nprof = min(args.steps, 100)
    for _ in range(nprof):
        pipe.step(d_heads[f % PERIOD], f, CONF, NMS); f += 1
    kus = pipe.kernel_us()
    pipe.set_profiling(False)
    lat = []
    for _ in range(30):
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        pipe.step(d_heads[f % PERIOD], f, CONF, NMS); f += 1
        pipe.join()
        b_.record(stream)
        torch.cuda.synchronize()
        lat.append(a.elapsed_time(b_) * 1e3)
    latency_us = float(np.median(lat))

    # ---------------- end to end through the host-buffer entry point (`e2e`) ----------------
    # page-locked host heads in (read in place over PCIe), TrackOutput records out (page-locked)
    nf = min(PERIOD, 8)
    pinned = torch.from_numpy(host_heads[:nf]).pin_memory()
    pinned_np = pinned.numpy()
    R = 4                                         # ring of page-locked result buffers, results consumed every R steps
    out_p = torch.zeros(R, B * DM * 228, dtype=torch.uint8).pin_memory()
    cnt_p = torch.zeros(R, B, dtype=torch.int32).pin_memory()
    out_np = [out_p[i].numpy().view(pb.TRACK_OUTPUT).reshape(B, DM) for i in range(R)]
    cnt_np = [cnt_p[i].numpy() for i in range(R)]

    def e2e_run(nsteps, f0):
        """pb_submit_host per step (in-place PCIe read of the heads, records copied back), pb_wait and a
        host-side read of every step's counts every R steps."""
        tracks = 0
        for i in range(nsteps):
            pipe.submit_host(pinned_np[i % nf], f0 + i, out_np[i % R], cnt_np[i % R], CONF, NMS)
            if (i + 1) % R == 0 or i == nsteps - 1:
                pipe.wait()
                tracks += int(sum(int(c.sum()) for c in cnt_np[: (i % R) + 1]))
        return tracks

    e2e_run(4, f); f += 4
    barrier()
    paths0 = pipe.nms_path_counts()
    t0 = time.perf_counter()
    e2e_tracks = e2e_run(args.e2e_steps, f); f += args.e2e_steps
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    paths1 = pipe.nms_path_counts()
    kp_fetch_mean = (paths1["keypoint_fetches"] - paths0["keypoint_fetches"]) / max(B * args.e2e_steps, 1)
    cand_mean = float(np.mean([pipe.get_kept(b)["num_cand"] for b in range(min(B, 8))]))

    # ---------------- max over ranks + final statistics gather (the only collective) ----------------
    stats = torch.tensor([ms_total, e2e_s, float(launches), float(counts.sum()), kus["gather_us"], kus["nms_us"], kus["track_us"],
                          float(e2e_tracks), latency_us], dtype=torch.float64, device=dev)
    if world > 1:
        gathered = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats)
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
    e2e_value = total_streams * args.e2e_steps / e2e_max
    gather_us, nms_us, track_us = (float(allstats[:, i].max()) for i in (4, 5, 6))

    peak, peak_src = measured_peak_gbs()
    traffic = ncu_traffic()
    def roof(name, bytes_per_launch, us, note):
        ach = bytes_per_launch / (us / 1e6) / 1e9 if us > 0 else 0.0
        return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic.get(name), "avg_launch_us": us, "algorithmic_bytes_per_launch": bytes_per_launch,
                "peak_source": peak_src, "note": note}
    r_gather = roof("pb_decode_gather_kernel", B * BYTES_HEAD, gather_us,
                    "dense-read model 224*N B per stream-frame (SURVEY.md 8d); the kernel reads the confidence row and 32 B sectors at "
                    "candidate anchors only, so this is an EFFECTIVE bandwidth; `traffic` is ncu dram bytes per launch")
    r_nms = roof("pb_nms_kernel", B * int(cand_mean) * 224, nms_us,
                 "candidate records (224 B each) read once; latency/issue-bound, working set in shared memory")
    r_track = roof("pb_tracker_kernel", B * (BYTES_TRACK + BYTES_OUT), track_us,
                   "track state read+written once + TrackOutput records; latency-bound (up to 150 dependent auction iterations per "
                   "stream-frame), HBM is not its limit")
    kernels = [r_gather, r_nms, r_track]
    dominant = max(kernels, key=lambda r: r["avg_launch_us"])
    step_ach = total_streams * BYTES_PER_STREAM_FRAME * args.steps / (ms_max / 1e3) / 1e9 / world
    line = {
        "metric": "tracked stream-frames/sec", "value": value, "unit": "stream-frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "us_per_batch": ms_max / args.steps * 1e3,
        "us_per_batch_latency": float(allstats[:, 8].max()),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "streams_per_gpu": B, "total_streams": total_streams, "parallelism": f"stream-sharded x{world}, no data-path collective",
                   "l2": f"{PERIOD} distinct head batches rotate ({PERIOD * B * 56 * N_ANCHORS * 4 / 1e9:.2f} GB per GPU > 126 MB L2)",
                   "conf": CONF, "nms": NMS, "max_tracks": T, "max_detections": DM, "max_age": MAX_AGE,
                   "pipeline_depth": args.pipeline_depth,
                   "timing": "value: K pb_step calls back to back + pb_join between CUDA events (steps overlap on internal streams); "
                             "us_per_batch_latency: median of single isolated steps; roofline: per-launch CUDA events on the serial path"},
        "roofline": dominant,
        "roofline_kernels": kernels,
        "roofline_step": {"bound": "hbm", "achieved": step_ach, "peak": peak, "unit": "GB/s", "frac": step_ach / peak,
                          "bytes_per_stream_frame": BYTES_PER_STREAM_FRAME, "note": "whole step per GPU, SURVEY.md 8(d) dense-read figure"},
        "e2e": {"value": e2e_value, "unit": "stream-frames/s",
                "h2d_bytes_per_step": int(B * (4 * N_ANCHORS + 4 * 32 * cand_mean + 51 * 32 * kp_fetch_mean)),
                "keypoint_fetches_per_stream_frame": kp_fetch_mean,
                "host_input_bytes_per_step": B * 56 * N_ANCHORS * 4,
                "d2h_bytes_per_step": B * DM * 228 + B * 4, "steps": args.e2e_steps,
                "api": "pb_submit_host / pb_wait (results read on the host every 4 steps): page-locked host heads [B,56,N] read in place over PCIe — confidence rows, 32 B sectors of the 4 box rows at every "
                       "candidate anchor and of the 51 keypoint rows at the candidates the lazy NMS sweep has to test (= h2d_bytes_per_step; the buffer "
                       "itself is host_input_bytes_per_step); TrackOutput records copied back into page-locked memory"},
        "gpu_launches": int(allstats[0, 2]),
        "clocks": clocks,
        "tracks_per_stream_frame": n_out_mean,
        "candidates_per_stream_frame": cand_mean,
    }
    if world == 1 and not args.no_cpu_baseline:
        cb = cpu_baseline(pb)
        ref_gpu = reference_gpu_b1(pb, torch)
        if ref_gpu:
            cb["reference_gpu_b1"] = ref_gpu
        line["cpu_baseline"] = cb
    out.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
