"""Per-launch CUDA-event times of the NMS kernel (full: 1024 threads, one CTA per SM) against the tiered kernels (PB_NMS_TIER = 1: 2 x 512
threads per SM, 2: 3 x 384, 3: 3 x 256, 4: 4 x 256); 592 streams = 4 CTAs on every SM, SM-time per stream-frame = kernel time * 148 / B."""
import os, sys, json, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    import posebyte_b200 as pb
    B = int(sys.argv[2])
    scfg = pb.synth_config(canvas=640, persons=20, period=32)
    d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, 8, frame_major=True)).cuda()
    pp = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors)
    for i in range(10): pp.step(d[i % 8], i)
    pp.set_profiling(True)
    for i in range(10, 60): pp.step(d[i % 8], i)
    k = pp.kernel_us()
    print(json.dumps({"B": B, "tier": os.environ.get("PB_NMS_TIER"), "kernel_us": k, "post_stage_us": pp.post_stage_us()}))
else:
    for B in (64, 296, 592):
        for tier in (None, "1", "2", "3", "4"):
            env = dict(os.environ)
            if tier: env["PB_NMS_TIER"] = tier
            subprocess.run([sys.executable, __file__, "child", str(B)], env=env)
