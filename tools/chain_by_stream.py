"""Bench workload (period-32 seamless motion): per-stream mean tracker chain and its stage split for the slowest streams."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
B, F = 64, 32
scfg = pb.synth_config(canvas=640, persons=20, period=32)
d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)).cuda()
pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors)
for f in range(64): pipe.step(d[f % F], f)
a0 = pipe.stream_stage_ns().astype(np.int64)
n = 128
for f in range(64, 64 + n): pipe.step(d[f % F], f)
a = (pipe.stream_stage_ns().astype(np.int64) - a0) / n / 1e3      # [B, 20] us per frame
names = {0: "prologue", 1: "predict", 2: "gate", 12: "t1cost", 13: "t1auction", 14: "t1lock", 4: "tier2", 5: "tier3", 6: "update", 7: "age", 8: "newtracks", 9: "dedup", 10: "total"}
order = np.argsort(-a[:, 10])
print("mean chain over streams", a[:, 10].mean().round(2), "slowest 5 streams:", a[order[:5], 10].round(2), "fastest:", a[order[-1], 10].round(2))
for b in order[:3]:
    print("stream", b, {n_: round(float(a[b, i]), 2) for i, n_ in names.items()}, "active", pipe.get_num_active()[b])
print("all-stream mean", {n_: round(float(a[:, i].mean()), 2) for i, n_ in names.items()})
