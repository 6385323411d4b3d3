"""Short bench-like run for ncu captures: B=64 streams, 640x640 heads, 24 steps."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import posebyte_b200 as pb
B = int(os.environ.get("PB_B", "64")); steps = int(os.environ.get("PB_STEPS", "24"))
scfg = pb.synth_config(canvas=640, persons=20, period=32)
d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, 8, frame_major=True)).cuda()
pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors)
for f in range(steps):
    pipe.step(d[f % 8], f)
torch.cuda.synchronize()
print("ok", pipe.get_num_active()[:4])
