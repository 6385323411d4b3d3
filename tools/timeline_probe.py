"""Timeline of the last steps (PB_TIMELINE=1): when do the NMS and tracker stages of consecutive steps begin and end, and
(fused path) which step's kernel ran a frame's tracker stage?"""
import os, sys
os.environ["PB_TIMELINE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
B, F = int(os.environ.get("PB_B", "64")), 32
scfg = pb.synth_config(canvas=640, persons=20, period=32)
d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)).cuda()
pp = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, pipeline_depth=int(os.environ.get("PB_DEPTH", "5")))
n = 200
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(40): pp.step(d[i % F], i)
pp.join(); torch.cuda.synchronize()
e0.record()
for i in range(40, 40 + n): pp.step(d[i % F], i)
pp.join(); e1.record(); torch.cuda.synchronize()
print("env", {k: v for k, v in os.environ.items() if k.startswith("PB_")}, "us/step", e0.elapsed_time(e1) / n * 1e3)
t = pp.debug_timeline().astype(np.float64)          # [64, B, 6]
last = 40 + n                                        # seq of the last step (seq = step + 1)
seqs = [last - k for k in range(24, 4, -1)]
t0 = t[seqs[0] & 63, :, 3].min()
print("seq | NMS begin (min,max) end (min,max) | tracker begin (min,mean,max) end (min,mean,max) | dur mean | ran by own kernel: streams | mean lag of runner")
for s in seqs:
    q = t[s & 63]
    r = (q[:, [0, 2, 3, 4]] - t0) / 1e3
    own = int((q[:, 1] == s).sum())
    lag = float((s - q[:, 1]).mean())
    print(f"{s:4d} | {r[:,2].min():7.1f} {r[:,2].max():7.1f}  {r[:,3].min():7.1f} {r[:,3].max():7.1f} | {r[:,0].min():7.1f} {r[:,0].mean():7.1f} {r[:,0].max():7.1f}  "
          f"{r[:,1].min():7.1f} {r[:,1].mean():7.1f} {r[:,1].max():7.1f} | {(r[:,1]-r[:,0]).mean():6.1f} | {own:3d} | {lag:4.2f}")
b = 0
print("stream 0:", [(int(s), round((t[s & 63, b, 3] - t0) / 1e3, 1), round((t[s & 63, b, 4] - t0) / 1e3, 1), round((t[s & 63, b, 0] - t0) / 1e3, 1),
                    round((t[s & 63, b, 2] - t0) / 1e3, 1), int(t[s & 63, b, 1])) for s in seqs[:10]])
