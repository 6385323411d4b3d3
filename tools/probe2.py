"""Development probe: where the step time goes (no per-kernel events), and zero-copy input."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb

B = int(os.environ.get("PB_B", "64"))
scfg = pb.synth_config(canvas=640, persons=20, period=32)
F = 16
host = pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)
d = torch.from_numpy(host).cuda()
pinned = torch.from_numpy(host).pin_memory()
pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors)

def timed(fn, n=200, warm=20):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(warm, warm + n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

print(f"B={B}")
print(f"step (device heads):      {timed(lambda i: pipe.step(d[i % F], i)):8.1f} us")
print(f"postprocess only:         {timed(lambda i: pipe.postprocess(d[i % F])):8.1f} us")
print(f"tracker only (own dets):  {timed(lambda i: pipe.tracker_update(i)):8.1f} us")
print("  tracker stages:", pipe.tracker_stage_us())
for depth in (2, 3, 4):
    pp = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, pipeline_depth=depth)
    def run(n, f0):
        for i in range(f0, f0 + n): pp.step(d[i % F], i)
        pp.join()
    run(20, 0); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(200, 20); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    print(f"pipelined step depth {depth}:   {us:8.1f} us  -> {B / us * 1e6:.0f} stream-frames/s")
    del pp
t0 = time.perf_counter()
us = timed(lambda i: pipe.step(pinned[i % F].data_ptr(), i), n=50, warm=5)
print(f"step (zero-copy pinned host heads, same kernels): {us:8.1f} us  -> {B / us * 1e6:.0f} stream-frames/s")
us = timed(lambda i: pipe.postprocess(pinned[i % F].data_ptr()), n=50, warm=5)
print(f"postprocess only, zero-copy: {us:8.1f} us")
hn = pinned.numpy()
t0 = time.perf_counter()
for i in range(30): pipe.step_host(hn[i % F], i)
print(f"step_host (memcpy staging): {(time.perf_counter() - t0) / 30 * 1e6:8.1f} us")
