"""Do consecutive tracker launches overlap?  Pipelined steps back to back, then the per-stream time spent waiting on
the predecessor's sequence flag (stage slot 15) and the kernel-level chain."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
B, F = int(os.environ.get("PB_B", "64")), int(os.environ.get("PB_F", "32"))
scfg = pb.synth_config(canvas=640, persons=20, period=32)
d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)).cuda()
pp = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, pipeline_depth=int(os.environ.get("PB_DEPTH", "3")))
for i in range(40): pp.step(d[i % F], i)
pp.join(); torch.cuda.synchronize()
a0 = pp.stream_stage_ns().astype(np.int64); a0[:, 18] = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 400
import time
ptrs = [d[i] for i in range(F)]
e0.record(); t0 = time.perf_counter()
for i in range(40, 40 + n): pp.step(ptrs[i % F], i)
t1 = time.perf_counter()
pp.join(); e1.record(); torch.cuda.synchronize(); t2 = time.perf_counter()
a = pp.stream_stage_ns().astype(np.int64); a[:, 18] = 0; a = (a - a0) / n / 1e3
print(f"B={B}: {e0.elapsed_time(e1) / n * 1e3:.1f} us/step (host enqueue {(t1 - t0) / n * 1e6:.1f} us/step, host total {(t2 - t0) / n * 1e6:.1f}) | per stream-frame: flag wait mean {a[:, 15].mean():.2f} max {a[:, 15].max():.2f} us, "
      f"det prologue {a[:, 16].mean():.2f}, release->go-on {a[:, 17].mean():.2f} (CTA not started yet: {a[:, 19].mean():.2f}), chain total mean {a[:, 10].mean():.2f}, slowest stream's mean chain {a[:, 10].max():.2f}")
