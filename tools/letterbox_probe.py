"""HBM-bound check of pb_letterbox_batch: B frames of WxH BGR u8 -> [B,3,640,640] fp32; effective GB/s over
source bytes read + target bytes written, against the measured copy bandwidth."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
peak = 6545.9
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
for (B, w, h, tw, th) in [(64, 1920, 1080, 640, 640), (64, 1280, 720, 640, 640), (256, 640, 480, 640, 640), (64, 1920, 1080, 1280, 1280)]:
    frames = torch.randint(0, 256, (B, w * h * 3), dtype=torch.uint8, device="cuda")
    sizes = torch.tensor([[w, h]] * B, dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for it in range(12):
        flush.fill_(it)                                   # evict the frames from the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out, xf = pb.letterbox_batch(frames, sizes, tw, th); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = float(np.median(ts[2:]))
    scale = min(tw / w, th / h); nw, nh = int(w * scale), int(h * scale)
    bytes_ = B * (w * h * 3 + 3 * tw * th * 4)
    print(json.dumps({"kernel": "letterbox_batch_kernel", "frames": B, "source": [w, h], "target": [tw, th], "us": round(t, 1),
                      "frames_per_s": round(B / t * 1e6), "algorithmic_bytes": bytes_, "achieved_gbs": round(bytes_ / t / 1e3, 1),
                      "peak_gbs": peak, "frac": round(bytes_ / t / 1e3 / peak, 3), "l2": "256 MB flush between launches"}))
