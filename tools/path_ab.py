"""A/B of the fused and the three-kernel step on BASELINE configs (one gpurun call): us per batch, pipelined and 20-step bursts."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb

def run(name, B, occlusion, max_age, depth, env, F=16, nstep=300):
    for k in ("PB_NO_FUSED", "PB_FUSED_THREADS", "PB_FUSED_COMPACT", "PB_FUSED_WAIT", "PB_DECODE_ON_LANE", "PB_FUSED_TIER"):
        os.environ.pop(k, None)
    os.environ.update(env)
    scfg = pb.synth_config(canvas=640, persons=20, period=32, occlusion=occlusion)
    key = (B, occlusion)
    if key not in cache:
        cache.clear(); torch.cuda.empty_cache()
        cache[key] = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)).cuda()
    d = cache[key]
    pp = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, pipeline_depth=depth, max_age=max_age)
    def go(n, f0):
        for i in range(f0, f0 + n): pp.step(d[i % F], i)
        pp.join()
    go(40, 0); torch.cuda.synchronize()
    res = []
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); go(nstep, 40 + nstep * rep); e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / nstep * 1e3)
    bursts = []
    f0 = 40 + 2 * nstep
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); go(20, f0 + 20 * rep); e1.record(); torch.cuda.synchronize()
        bursts.append(e0.elapsed_time(e1) / 20 * 1e3)
    print(json.dumps({"case": name, "B": B, "depth": depth, "env": env, "us_per_batch": [round(x, 1) for x in res], "burst20": round(float(np.median(bursts)), 1),
                      "Msf_per_s": round(B / min(res), 3)}), flush=True)
    del pp

cache = {}
for B, occ, age in ((64, 0, 10), (128, 1, 30), (1, 0, 10), (148, 0, 10), (32, 0, 10)):
    name = f"B{B}"
    run(name, B, occ, age, 5, {"PB_NO_FUSED": "1"})
    run(name, B, occ, age, 5, {})
    run(name, B, occ, age, 4, {"PB_FUSED_COMPACT": "0"})
    run(name, B, occ, age, 5, {"PB_FUSED_COMPACT": "0"})
    run(name, B, occ, age, 8, {"PB_FUSED_WAIT": "3"})
