"""pb_step_host timing with page-locked buffers (zero-copy input, lazy keypoints)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
B = 64
scfg = pb.synth_config(canvas=640, persons=20, period=32)
host = pb.synth_heads(scfg, 0, B, 0, 8, frame_major=True)
pinned = torch.from_numpy(host).pin_memory(); hn = pinned.numpy()
for mode in (0, 2):
    pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, keypoint_fetch=mode)
    out_p = torch.zeros(B * pipe.Dm * 228, dtype=torch.uint8).pin_memory(); cnt_p = torch.zeros(B, dtype=torch.int32).pin_memory()
    out_np = out_p.numpy().view(pb.TRACK_OUTPUT).reshape(B, pipe.Dm); cnt_np = cnt_p.numpy()
    for i in range(5): pipe.step_host(hn[i % 8], i, out=out_np, counts=cnt_np)
    res = []
    for rep in range(3):
        t0 = time.perf_counter()
        for i in range(40): pipe.step_host(hn[i % 8], 5 + i, out=out_np, counts=cnt_np)
        res.append((time.perf_counter() - t0) / 40 * 1e6)
    print(f"keypoint_fetch={mode}: step_host us/step {[round(x, 1) for x in res]} -> {B / min(res) * 1e6:.0f} stream-frames/s", pipe.nms_path_counts(), pipe.post_stage_us())

for depth in (1, 3):
    pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, pipeline_depth=depth)
    R = 4
    outs = torch.zeros(R, B * pipe.Dm * 228, dtype=torch.uint8).pin_memory(); cnts = torch.zeros(R, B, dtype=torch.int32).pin_memory()
    def run(n, f0):
        tracks = 0
        for i in range(n):
            if i >= R: pass
            pipe.submit_host(hn[i % 8], f0 + i, outs[i % R].numpy(), cnts[i % R].numpy())
            if (i + 1) % R == 0:
                pipe.wait(); tracks += int(cnts.sum())
        pipe.wait()
        return tracks
    run(8, 0)
    res = []
    for rep in range(3):
        t0 = time.perf_counter(); run(40, 8 + 40 * rep); res.append((time.perf_counter() - t0) / 40 * 1e6)
    print(f"submit_host depth {depth}, wait every {R} steps: us/step {[round(x, 1) for x in res]} -> {B / min(res) * 1e6:.0f} stream-frames/s")
