"""A/B timing of tracker build variants inside one gpurun call (run it once per variant, selected by an
environment variable the build under test reads): pipelined step time + tier-1 auction tail.  Single runs on
different boxes differ by more than the effects being measured; runs inside one call repeat to 0.2 %."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
B, F = int(os.environ.get("PB_B", "64")), 16
scfg = pb.synth_config(canvas=640, persons=20, period=32)
d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)).cuda()
pp = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, pipeline_depth=3)
def run(n, f0):
    for i in range(f0, f0 + n): pp.step(d[i % F], i)
    pp.join()
run(40, 0); torch.cuda.synchronize()
res = []
for rep in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(400, 40 + 400 * rep); e1.record(); torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 400 * 1e3)
pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors)
for f in range(32): pipe.step(d[f % F], f)
prev = pipe.stream_stage_ns().astype(np.int64); rows = []
for f in range(32, 96):
    pipe.step(d[f % F], f); cur = pipe.stream_stage_ns().astype(np.int64); rows.append(cur - prev); prev = cur
a = np.stack(rows) / 1e3
print(f"variant {os.environ.get('PB_VARIANT', '0')}: pipelined us/step {[round(x, 1) for x in res]} | t1auction mean {a[:, :, 13].mean():.1f} p90 {np.percentile(a[:, :, 13], 90):.1f} max {a[:, :, 13].max():.1f} | chain: mean {a[:, :, 10].mean():.1f}, mean of per-frame max {a[:, :, 10].max(1).mean():.1f}")
