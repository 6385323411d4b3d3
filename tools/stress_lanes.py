"""Long run of the pipelined three-kernel step (three lanes: overlapping tracker grids that hand a stream's state from CTA to
CTA — release / acquire flags, bulk-copy staging behind a proxy fence) against the serial path: the complete tracker state
must stay bit-identical.  PB_STEPS steps (default 20000), compared every PB_EVERY steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
B, F = int(os.environ.get("PB_B", "64")), 32
N, EVERY = int(os.environ.get("PB_STEPS", "20000")), int(os.environ.get("PB_EVERY", "2000"))
scfg = pb.synth_config(canvas=640, persons=20, period=F, occlusion=int(os.environ.get("PB_OCC", "1")))
d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)).cuda()
kw = dict(num_streams=B, num_anchors=scfg.num_anchors, max_age=int(os.environ.get("PB_MAX_AGE", "10")))
serial = pb.Pipeline(pipeline_depth=1, fuse_stages=0, **kw)
piped = pb.Pipeline(pipeline_depth=5, **kw)
print("paths", serial.step_path(), piped.step_path(), piped.nms_plan())
bad = 0
for f0 in range(0, N, EVERY):
    for f in range(f0, f0 + EVERY):
        serial.step(d[f % F], f)
    piped.step_seq(d, f0 % F, EVERY, f0)
    piped.join(); torch.cuda.synchronize()
    a, b = serial.state_save()[24:], piped.state_save()[24:]
    o1, c1 = serial.get_tracks_all(); o2, c2 = piped.get_tracks_all()
    same = a == b and np.array_equal(c1, c2) and all(o1[s, :c1[s]].tobytes() == o2[s, :c2[s]].tobytes() for s in range(B))
    print(f"after {f0 + EVERY} steps: state and records {'identical' if same else 'DIFFER'} (tracks {int(c1.sum())})", flush=True)
    bad += 0 if same else 1
    if bad: break
print("stress", "FAILED" if bad else "ok", N, "steps", B, "streams")
sys.exit(1 if bad else 0)
