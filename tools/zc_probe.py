"""Zero-copy (page-locked host input) decode+gather timing under different L2 fetch granularities."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
B = 64
scfg = pb.synth_config(canvas=640, persons=20, period=32)
F = 8
host = pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)
pinned = torch.from_numpy(host).pin_memory()
pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors)
def timed(fn, n=40, warm=5):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(warm, warm + n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("PB_L2_FETCH", os.environ.get("PB_L2_FETCH"), f"postprocess zero-copy {timed(lambda i: pipe.postprocess(pinned[i % F].data_ptr())):8.1f} us")
