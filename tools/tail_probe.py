"""Per-stream, per-frame tracker chain time: how heavy is the tail that sets the kernel time?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
B, F = 64, 16
scfg = pb.synth_config(canvas=640, persons=20, period=32)
d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)).cuda()
pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors)
print("env", {k: v for k, v in os.environ.items() if k.startswith("PB_")})
for f in range(32): pipe.step(d[f % F], f)
prev = pipe.stream_stage_ns().astype(np.int64)
rows = []
for f in range(32, 96):
    pipe.step(d[f % F], f)
    cur = pipe.stream_stage_ns().astype(np.int64)
    rows.append(cur - prev); prev = cur
a = np.stack(rows) / 1e3            # [frames, B, 20] us
names = {15: "flagwait", 16: "detprolog", 0: "prologue", 1: "predict", 2: "gate", 12: "t1cost", 13: "t1auction", 14: "t1lock", 3: "tier1", 4: "tier2", 5: "tier3", 6: "update", 7: "age", 8: "newtracks", 9: "dedup", 10: "total"}
for i, n in names.items():
    x = a[:, :, i]
    print(f"{n:10s} mean {x.mean():6.2f}  p50 {np.percentile(x, 50):6.2f}  p90 {np.percentile(x, 90):6.2f}  p99 {np.percentile(x, 99):6.2f}  max {x.max():6.2f} | mean over frames of max over streams {x.max(1).mean():6.2f}")
tot = a[:, :, 10]
worst = np.unravel_index(tot.argmax(), tot.shape)
print("worst stream-frame:", worst, {n: round(float(a[worst[0], worst[1], i]), 2) for i, n in names.items()})
print("post stages us:", pipe.post_stage_us())
na = pipe.get_num_active()
print("num_active per stream: min", na.min(), "max", na.max())
tot = a[:, :, 10]                       # [frames, B] us
for K in (1, 2, 4, 8, 16):
    n = (tot.shape[0] // K) * K
    g = tot[:n].reshape(n // K, K, -1).sum(1)            # [groups, B]
    print(f"tracker launch over K={K:2d} frames: mean over launches of (max over streams of summed chain)/K = {g.max(1).mean() / K:6.2f} us per step   (mean chain {tot.mean():.2f})")
