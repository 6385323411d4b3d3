"""Throughput of the BASELINE.json configurations 2-5 on one GPU (device-resident heads)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb

def run(name, B, canvas, persons, T, Dm, max_age, clumps=0, occlusion=0, steps=200, depth=5, F=8):
    scfg = pb.synth_config(canvas=canvas, persons=persons, period=64, clumps=clumps, occlusion=occlusion,
                           kp_drop_prob=0.15 if clumps else 0.05)
    d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)).cuda()
    res = {}
    for mode, dep in (("serial", 1), ("pipelined", depth)):
        pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, max_tracks=T, max_detections=Dm, max_age=max_age, pipeline_depth=dep)
        for f in range(20): pipe.step(d[f % F], f)
        pipe.join(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for f in range(20, 20 + steps): pipe.step(d[f % F], f)
        pipe.join(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / steps * 1e3
        res[mode] = us
        if mode == "serial":
            pipe.set_profiling(True)
            for f in range(50): pipe.step(d[f % F], f)
            k = pipe.kernel_us(); pipe.set_profiling(False)
            kept = pipe.get_kept(0)
            na = pipe.get_num_active()
    bytes_sf = 224 * scfg.num_anchors + 736 * T + 228 * persons
    print(json.dumps({"config": name, "B": B, "N": scfg.num_anchors, "T": T, "Dm": Dm, "us_per_batch_serial": round(res["serial"], 1),
                      "us_per_batch_pipelined": round(res["pipelined"], 1), "stream_frames_per_s": round(B / res["pipelined"] * 1e6),
                      "hbm_roofline_frac": round(B * bytes_sf / (res["pipelined"] * 1e-6) / 6545.9e9, 3),
                      "kernel_us": {k_: round(v, 1) for k_, v in k.items() if k_ != "launches"},
                      "cand": kept["num_cand"], "kept": kept["num_keep"], "active_tracks": int(na.mean())}), flush=True)

def run_tracker_only(name, B, T, Dm, gating, steps=40):
    scfg = pb.synth_config(canvas=4096, persons=Dm, period=64, max_speed=2.0)
    import ctypes as C
    poses = np.zeros((8, B, Dm, 51), np.float32); scores = np.zeros((8, B, Dm), np.float32)
    for f in range(8):
        for b in range(B):
            p, s = pb.synth_dets(scfg, b, f)
            n = len(s); o = np.argsort(-s, kind="stable")
            poses[f, b, :n] = p[o]; scores[f, b, :n] = s[o]
    dp, ds = torch.from_numpy(poses).cuda(), torch.from_numpy(scores).cuda()
    num = torch.full((B,), Dm, dtype=torch.int32, device="cuda")
    pipe = pb.Pipeline(num_streams=B, num_anchors=64, max_candidates=64, max_keep=64, max_tracks=T, max_detections=Dm, gating_enabled=gating)
    for f in range(6): pipe.tracker_update(f, dp[f % 8], ds[f % 8], num, Dm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for f in range(6, 6 + steps): pipe.tracker_update(f, dp[f % 8], ds[f % 8], num, Dm)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / steps * 1e3
    print(json.dumps({"config": name, "B": B, "T": T, "Dm": Dm, "gating": gating, "us_per_batch": round(us, 1), "us_per_stream_frame": round(us / B, 1),
                      "stream_frames_per_s": round(B / us * 1e6), "active_tracks": int(pipe.get_num_active().mean()), "stages_us": pipe.tracker_stage_us()}), flush=True)

if __name__ == "__main__":
    run("cfg2 640 x64 streams", 64, 640, 20, 128, 64, 10)
    run("cfg2 640 x148 streams", 148, 640, 20, 128, 64, 10)
    run("cfg4 640 x128 streams max-age 30 occlusion", 128, 640, 20, 128, 64, 30, occlusion=1)
    run("cfg3 1280 crowd x16", 16, 1280, 100, 256, 128, 10, clumps=10, steps=60)
    run("cfg3 1280 crowd x64", 64, 1280, 100, 256, 128, 10, clumps=10, steps=40, F=4)
    run_tracker_only("cfg5 512x512 gating on x8", 8, 512, 512, 1)
    run_tracker_only("cfg5 512x512 gating off x8", 8, 512, 512, 0)
