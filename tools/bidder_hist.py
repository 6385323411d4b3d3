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
it = a[:, :, 15:20] / 1000.0
tot = a[:, :, 10] / 1e3
print("iterations per stream-frame (3 solves) by path [chain(1), 2, 3-4, 5-8, row scan]: mean", it.mean((0, 1)).round(2))
slow = tot >= np.percentile(tot, 90)
print("slowest 10% of stream-frames:", it[slow].mean(0).round(2), " chain us", tot[slow].mean().round(1))
w = tot.argmax(1)
print("per-frame slowest stream:", np.stack([it[f, w[f]] for f in range(it.shape[0])]).mean(0).round(2), " chain us", tot.max(1).mean().round(1))
print("t1auction us: mean", (a[:, :, 13] / 1e3).mean().round(1), "max-per-frame mean", (a[:, :, 13] / 1e3).max(1).mean().round(1))
t1 = a[:, :, 13] / 1e3
wf, ws = np.unravel_index(t1.argmax(), t1.shape)
print("worst tier-1 auction:", t1[wf, ws].round(1), "us, frame", wf, "stream", ws, "iterations by path (3 solves):", it[wf, ws])
for q in (50, 90, 99):
    sel = t1 >= np.percentile(t1, q)
    print(f"stream-frames with t1auction >= p{q} ({np.percentile(t1, q):.1f} us): iterations by path", it[sel].mean(0).round(2))
