"""Distribution of the number of bidders per auction iteration (telemetry slots 15-19 of the hybrid solve)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
B, F = 64, 16
scfg = pb.synth_config(canvas=640, persons=20, period=32)
d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)).cuda()
pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors)
for f in range(32): pipe.step(d[f % F], f)
prev = pipe.stream_stage_ns().astype(np.int64); rows = []
for f in range(32, 96):
    pipe.step(d[f % F], f); cur = pipe.stream_stage_ns().astype(np.int64); rows.append(cur - prev); prev = cur
a = np.stack(rows).astype(np.float64)      # [frames, B, 20]
it = a[:, :, 15:18] / 1000.0
cyc = a[:, :, 18:20]
tot = a[:, :, 10] / 1e3
print("iterations per stream-frame (3 solves) by bidder count [1, 2, >2]: mean", it.mean((0, 1)).round(2))
slow = tot >= np.percentile(tot, 90)
print("slowest 10% of stream-frames:", it[slow].mean(0).round(2), " chain us", tot[slow].mean().round(1))
w = tot.argmax(1)
print("per-frame slowest stream:", np.stack([it[f, w[f]] for f in range(it.shape[0])]).mean(0).round(2), " chain us", tot.max(1).mean().round(1))
print("t1auction us: mean", (a[:, :, 13] / 1e3).mean().round(1), "max-per-frame mean", (a[:, :, 13] / 1e3).max(1).mean().round(1))
print("cycles per stream-frame in the iteration loops (3 solves): mean", cyc[:, :, 0].mean().round(0), " whole solves:", cyc[:, :, 1].mean().round(0),
      " cycles per iteration:", (cyc[:, :, 0].sum() / it.sum()).round(0))
print("slowest stream per frame: loop cycles", np.mean([cyc[f, w[f], 0] for f in range(len(w))]).round(0), "iterations", np.mean([it[f, w[f]].sum() for f in range(len(w))]).round(1))
