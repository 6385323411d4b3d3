"""Resident-tracker path of pb_step_seq against the per-step paths: us per batch (steady: one call of `nstep` steps; bursts of 20
as the driver times them) and the final state compared with the per-step path's (one gpurun call)."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb

KEYS = ("PB_SEQ", "PB_SEQ_COMPACT", "PB_SEQ_THREADS", "PB_SEQ_CHUNK", "PB_NO_FUSED", "PB_SEQ_NMS_TIER", "PB_SEQ_RESIDENT_STATE", "PB_SEQ_LANES")

def run(name, B, occlusion, max_age, depth, env, F=16, nstep=320, persons=20, canvas=640, T=128, Dm=64):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)
    scfg = pb.synth_config(canvas=canvas, persons=persons, period=32, occlusion=occlusion)
    key = (B, occlusion, canvas, persons)
    if key not in cache:
        cache.clear(); torch.cuda.empty_cache()
        cache[key] = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)).cuda()
    d = cache[key]
    pp = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, pipeline_depth=depth, max_age=max_age, max_tracks=T, max_detections=Dm)
    f = [0]
    host = [0.0, 0]
    def go(n):
        t0 = time.perf_counter()
        pp.step_seq(d, f[0] % F, n, f[0]); f[0] += n
        host[0] += time.perf_counter() - t0; host[1] += n
        pp.join()
    go(40); torch.cuda.synchronize()
    res = []
    ns0 = pp.stream_stage_ns().astype(np.float64)
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); go(nstep); e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / nstep * 1e3)
    ns1 = pp.stream_stage_ns().astype(np.float64)
    dn = ns1 - ns0
    fr = np.maximum(dn[:, 11], 1)
    tele = {"trk_us_per_frame": round(float((dn[:, 10] / fr).mean() / 1e3), 2), "trk_us_per_frame_max_stream": round(float((dn[:, 10] / fr).max() / 1e3), 2),
            "wait_us_per_frame": round(float((dn[:, 15] / fr).mean() / 1e3), 2)}
    bursts = []
    host[0], host[1] = 0.0, 0
    for rep in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); go(20); e1.record(); torch.cuda.synchronize()
        bursts.append(e0.elapsed_time(e1) / 20 * 1e3)
    tele["post_stage_us"] = pp.post_stage_us()
    state = pp.state_save()[24:]
    o, c = pp.get_tracks_all()
    same = None
    if key + (max_age,) in ref:
        same = bool(ref[key + (max_age,)] == state)
    else:
        ref[key + (max_age,)] = state
    print(json.dumps({"case": name, "B": B, "depth": depth, "env": env, "us_per_batch": [round(x, 1) for x in res], "burst20": round(float(np.median(bursts)), 1),
                      "burst20_min": round(float(np.min(bursts)), 1), "Msf_per_s": round(B / min(res), 3), "tracks": int(c.sum()), "state_equals_first_case": same, "host_us_per_step": round(host[0] / host[1] * 1e6, 1), **tele}), flush=True)
    pp.close()

cache, ref = {}, {}
which = sys.argv[1:] or ["64"]
for w in which:
    B = int(w)
    occ, age = (1, 30) if B == 128 else (0, 10)
    name = f"B{B}"
    run(name, B, occ, age, 5, {"PB_SEQ": "0", "PB_NO_FUSED": "1"})
    run(name, B, occ, age, 5, {"PB_SEQ": "1"})
    if os.environ.get("SEQ_PROBE_OLD"): run(name, B, occ, age, 5, {"PB_SEQ": "1", "PB_SEQ_COMPACT": "1", "PB_SEQ_NMS_TIER": "1"})
    if os.environ.get("SEQ_PROBE_TIERS"):
        for tier in os.environ["SEQ_PROBE_TIERS"].split(","):
            for lanes in os.environ.get("SEQ_PROBE_LANES", "3,4").split(","):
                run(name, B, occ, age, 5, {"PB_SEQ": "1", "PB_SEQ_NMS_TIER": tier, "PB_SEQ_LANES": lanes})
        continue
    if os.environ.get("SEQ_PROBE_QUICK"):
        for lanes in ("3", "4"):
            for tier in ("0", "1"):
                run(name, B, occ, age, 5, {"PB_SEQ": "1", "PB_SEQ_NMS_TIER": tier, "PB_SEQ_LANES": lanes})
        continue
    # resident tracker CTAs of 512 / 256 threads leave registers and thread slots of their SM to decode CTAs
    for thr in ("512", "256"):
        for tier in ("0", "1", "2"):
            for lanes in ("3", "4"):
                run(name, B, occ, age, 5, {"PB_SEQ": "1", "PB_SEQ_THREADS": thr, "PB_SEQ_NMS_TIER": tier, "PB_SEQ_LANES": lanes})
