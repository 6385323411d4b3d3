"""Small run of every product kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
B, F = 3, 6
scfg = pb.synth_config(canvas=640, persons=8, period=16, occlusion=1)
host = pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)
d = torch.from_numpy(host).cuda()
for kw in (dict(), dict(keypoint_fetch=1), dict(pipeline_depth=3), dict(max_tracks=512, max_detections=512, max_keep=512)):
    pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, **kw)
    for f in range(F): pipe.step(d[f], f)
    pipe.join(); torch.cuda.synchronize()
    print(kw, pipe.get_num_active(), flush=True)
pinned = torch.from_numpy(host).pin_memory()
pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors)
for f in range(F): o, c = pipe.step_host(pinned[f].numpy(), f)
print("host", c, flush=True)
# stand-alone entry points
import ctypes as C
L = pb.lib()
cost = torch.rand(4, 40, 24, device="cuda"); row = torch.zeros(4, 40, dtype=torch.int32, device="cuda"); col = torch.zeros(4, 24, dtype=torch.int32, device="cuda")
pb.check(L.pb_auction_solve(cost.data_ptr(), 4, 40, 24, row.data_ptr(), col.data_ptr(), None, None))
means = torch.zeros(8, 136, device="cuda"); diag = torch.zeros(8, 136, device="cuda"); dets = torch.rand(8, 51, device="cuda") * 100
slots = torch.arange(8, dtype=torch.int32, device="cuda"); m = torch.stack([slots, slots], 1).contiguous()
pb.check(L.pb_kf3_initiate(means.data_ptr(), diag.data_ptr(), dets.data_ptr(), slots.data_ptr(), 8, None))
pb.check(L.pb_kf3_predict(means.data_ptr(), diag.data_ptr(), 8, 0.9, 0.9, None))
pb.check(L.pb_kf3_update(means.data_ptr(), diag.data_ptr(), dets.data_ptr(), m.data_ptr(), 8, None))
torch.cuda.synchronize()
print("aux ok", flush=True)
