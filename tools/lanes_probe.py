"""A/B of the step pipeline's lane count (PB_LANES) and ring depth inside one gpurun call:
pipelined us/step, and every stream's TrackOutput records compared with the serial path (depth 1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
B, F = int(os.environ.get("PB_B", "64")), 32
scfg = pb.synth_config(canvas=640, persons=20, period=32)
d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)).cuda()
NSTEP = int(os.environ.get("PB_STEPS", "400"))
def measure(lanes, depth, reps=3):
    os.environ["PB_LANES"] = str(lanes)
    pp = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, pipeline_depth=depth, max_candidates=int(os.environ.get("PB_MAXCAND", "1024")))
    def run(n, f0):
        for i in range(f0, f0 + n): pp.step(d[i % F], i)
        pp.join()
    run(40, 0); torch.cuda.synchronize()
    res = []
    for rep in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(NSTEP, 40 + NSTEP * rep); e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / NSTEP * 1e3)
    out = [pp.get_tracks(b).tobytes() for b in range(B)]
    del pp
    return res, out
ref_t, ref = measure(0, 1, reps=3)
print(f"serial depth 1: us/step {[round(x, 1) for x in ref_t]}")
for lanes, depth in ([(3, 5)] if os.environ.get("PB_QUICK") else [(0, 3), (0, 3), (2, 3), (2, 4), (3, 4), (3, 5)]):
    t, out = measure(lanes, depth)
    bad = [b for b in range(B) if out[b] != ref[b]]
    print(f"lanes {lanes} depth {depth} B {B}: pipelined us/step {[round(x, 1) for x in t]}  streams differing from the serial path: {len(bad)} {bad[:8]}")
