import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
for canvas, persons, clumps, B, T, Dm in ((640, 20, 0, 64, 128, 64), (1280, 100, 10, 16, 256, 128)):
    scfg = pb.synth_config(canvas=canvas, persons=persons, period=32, clumps=clumps, kp_drop_prob=0.15 if clumps else 0.05)
    d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, 8, frame_major=True)).cuda()
    pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, max_tracks=T, max_detections=Dm)
    for f in range(8): pipe.postprocess(d[f])
    torch.cuda.synchronize()
    k = pipe.get_kept(0)
    print(canvas, persons, "paths:", pipe.nms_path_counts(), "cand", k["num_cand"], "kept", k["num_keep"], pipe.post_stage_us())
