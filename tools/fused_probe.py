"""A/B of the fused per-stream kernel inside one gpurun call: pipelined us/step for several CTA sizes / lane counts / tiers
(environment knobs of pb_create), each compared with the serial three-kernel path for identical TrackOutput records."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
B, F = int(os.environ.get("PB_B", "64")), 32
scfg = pb.synth_config(canvas=640, persons=20, period=32)
d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)).cuda()
NSTEP = int(os.environ.get("PB_STEPS", "400"))

def measure(env, depth, reps=3, nstep=NSTEP):
    for k in ("PB_NO_FUSED", "PB_FUSED_THREADS", "PB_LANES", "PB_FUSED_LANES", "PB_FUSED_TIER", "PB_FUSED_AGE", "PB_FUSED_COMPACT", "PB_FUSED_WAIT", "PB_DECODE_ON_LANE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    pp = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, pipeline_depth=depth)
    def run(n, f0):
        for i in range(f0, f0 + n): pp.step(d[i % F], i)
        pp.join()
    run(40, 0); torch.cuda.synchronize()
    res = []
    for rep in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(nstep, 40 + nstep * rep); e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / nstep * 1e3)
    # short bursts as the driver times them: 20 steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0 = 40 + nstep * reps
    e0.record(); run(20, f0); e1.record(); torch.cuda.synchronize()
    burst = e0.elapsed_time(e1) / 20 * 1e3
    lat = []
    for i in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); run(1, f0 + 20 + i); e1.record(); torch.cuda.synchronize()
        lat.append(e0.elapsed_time(e1) * 1e3)
    out = [pp.get_tracks(b).tobytes() for b in range(B)]
    del pp
    return res, burst, float(np.median(lat)), out

ref_t, _, ref_lat, ref = measure({"PB_NO_FUSED": "1"}, 1, reps=3, nstep=NSTEP)
print(f"serial three-kernel path: us/step {[round(x, 1) for x in ref_t]} latency {ref_lat:.1f}", flush=True)
variants = [({"PB_NO_FUSED": "1"}, 5), ({}, 3), ({}, 4), ({}, 5), ({}, 8), ({"PB_FUSED_WAIT": "1"}, 3), ({"PB_FUSED_WAIT": "1"}, 4), ({"PB_FUSED_WAIT": "1"}, 5),
            ({"PB_FUSED_WAIT": "0"}, 5), ({"PB_DECODE_ON_LANE": "1"}, 4), ({"PB_DECODE_ON_LANE": "1"}, 5), ({"PB_FUSED_WAIT": "3"}, 5),
            ({"PB_FUSED_COMPACT": "0"}, 4), ({"PB_FUSED_COMPACT": "0", "PB_FUSED_THREADS": "1024"}, 4)]
if len(sys.argv) > 1:
    variants = [eval(a) for a in sys.argv[1:]]
for env, depth in variants:
    t, burst, lat, out = measure(env, depth)
    bad = [b for b in range(B) if out[b] != ref[b]]
    print(f"{env} depth {depth} B {B}: us/step {[round(x, 1) for x in t]} 20-step burst {burst:.1f} latency {lat:.1f} | streams differing from serial: {len(bad)} {bad[:8]}", flush=True)
