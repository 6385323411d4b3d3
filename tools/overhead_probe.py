"""Kernel time with empty inputs (no candidates, no tracks): fixed per-launch overhead probe."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import posebyte_b200 as pb
for B in (1, 64, 148):
    pipe = pb.Pipeline(num_streams=B)
    heads = torch.zeros(B, 56, 8400, device="cuda")
    for f in range(10): pipe.step(heads, f)
    torch.cuda.synchronize()
    pipe.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for f in range(100): pipe.step(heads, f)
    e1.record(); torch.cuda.synchronize()
    k = pipe.kernel_ms()
    print(f"B={B}: empty step {e0.elapsed_time(e1)/100*1e3:.1f} us | post {k['post_ms']/k['post_launches']*1e3:.1f} us track {k['track_ms']/k['track_launches']*1e3:.1f} us",
          pipe.post_stage_us(), pipe.tracker_stage_us()["total_us"])
    pipe.set_profiling(False)
    e0.record()
    for f in range(100): pipe.step(heads, f)
    e1.record(); torch.cuda.synchronize()
    print(f"   without per-kernel events: {e0.elapsed_time(e1)/100*1e3:.1f} us/step")
