"""Short fixed workloads for ncu captures.  PB_CASE = cfg2 (64 streams x [56,8400], the three-kernel step; default),
cfg4 (128 streams, occlusion, max-age 30: the fused per-stream kernel), cfg5 (512 x 512 tracker tables, 8 streams:
row-sliced pre-kernel + per-stream kernel), res32 / res64 (32 / 64 streams through pb_step_seq: the resident tracker kernel and the paired NMS CTAs; under ncu, which
replays kernels one at a time, run it with PB_SEQ_TRACKER_LAST=1 PB_SEQ_CHUNK=8)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
case = os.environ.get("PB_CASE", "cfg2"); steps = int(os.environ.get("PB_STEPS", "24"))
if case in ("res32", "res64"):
    B = 32 if case == "res32" else 64
    scfg = pb.synth_config(canvas=640, persons=20, period=32)
    d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, 8, frame_major=True)).cuda()
    pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, pipeline_depth=5)
    pipe.step_seq(d, 0, steps, 0)
    pipe.join()
elif case in ("cfg2", "cfg4"):
    B = 64 if case == "cfg2" else 128
    scfg = pb.synth_config(canvas=640, persons=20, period=32, occlusion=int(case == "cfg4"))
    d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, 8, frame_major=True)).cuda()
    pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, max_age=10 if case == "cfg2" else 30,
                       pipeline_depth=int(os.environ.get("PB_DEPTH", "1" if case == "cfg2" else "5")))
    for f in range(steps):
        pipe.step(d[f % 8], f)
    pipe.join()
else:
    B, T = 8, 512
    scfg = pb.synth_config(canvas=4096, persons=T, period=64, max_speed=2.0)
    poses = np.zeros((4, B, T, 51), np.float32); scores = np.zeros((4, B, T), np.float32)
    for f in range(4):
        for b in range(B):
            p, s = pb.synth_dets(scfg, b, f); o = np.argsort(-s, kind="stable")
            poses[f, b, :len(o)] = p[o]; scores[f, b, :len(o)] = s[o]
    dp, ds = torch.from_numpy(poses).cuda(), torch.from_numpy(scores).cuda()
    num = torch.full((B,), T, dtype=torch.int32, device="cuda")
    pipe = pb.Pipeline(num_streams=B, num_anchors=64, max_candidates=64, max_keep=64, max_tracks=T, max_detections=T,
                       gating_enabled=int(os.environ.get("PB_GATING", "1")))
    for f in range(steps):
        pipe.tracker_update(f, dp[f % 4], ds[f % 4], num, T)
torch.cuda.synchronize()
print("ok", case, pipe.get_num_active()[:4])
