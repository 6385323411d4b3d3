"""Timeline of the last launches (PB_TIMELINE=1): when do tracker / NMS CTAs of consecutive steps begin and end?"""
import os, sys
os.environ["PB_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
B, F = int(os.environ.get("PB_B", "64")), 32
scfg = pb.synth_config(canvas=640, persons=20, period=32)
d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)).cuda()
pp = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, pipeline_depth=int(os.environ.get("PB_DEPTH", "3")))
n = 200
for i in range(n): pp.step(d[i % F], i)
pp.join(); torch.cuda.synchronize()
t = pp.debug_timeline().astype(np.float64)          # [64, B, 6]
seqs = [(n - k) for k in range(12, 0, -1)]          # the last 12 launches, oldest first (seq = step + 1)
t0 = t[seqs[0] & 63, :, 0].min()
print("step | tracker: first CTA begin, last begin, first acquired, last acquired, first end, last end | NMS: first begin, last begin, first end, last end   (us, relative)")
for s in seqs:
    q = (t[s & 63] - t0) / 1e3
    print(f"{s:4d} | {q[:,0].min():7.1f} {q[:,0].max():7.1f} {q[:,1].min():7.1f} {q[:,1].max():7.1f} {q[:,2].min():7.1f} {q[:,2].max():7.1f} | {q[:,3].min():7.1f} {q[:,3].max():7.1f} {q[:,4].min():7.1f} {q[:,4].max():7.1f}")
