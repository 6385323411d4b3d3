"""Per-stage device time of both kernels on the bench workload (development aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb

def main(B=64, canvas=640, persons=20, period=32, T=128, Dm=64, steps=128, occlusion=0, clumps=0):
    scfg = pb.synth_config(canvas=canvas, persons=persons, period=period, occlusion=occlusion, clumps=clumps,
                           kp_drop_prob=0.15 if clumps else 0.05)
    F = min(period, 16)
    d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)).cuda()
    pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, max_tracks=T, max_detections=Dm)
    for f in range(32): pipe.step(d[f % F], f)
    torch.cuda.synchronize()
    pipe.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for f in range(32, 32 + steps): pipe.step(d[f % F], f)
    e1.record(); torch.cuda.synchronize()
    k = pipe.kernel_ms()
    print(f"B={B} canvas={canvas} P={persons} T={T} Dm={Dm}: step {e0.elapsed_time(e1)/steps*1e3:.1f} us | "
          f"post {k['post_ms']/k['post_launches']*1e3:.1f} us  track {k['track_ms']/k['track_launches']*1e3:.1f} us")
    print("  post stages us:", pipe.post_stage_us())
    print("  track stages us:", pipe.tracker_stage_us())
    kept = pipe.get_kept(0)
    print("  stream0: cand", kept["num_cand"], "kept", kept["num_keep"], "active", pipe.get_num_active()[:4])

if __name__ == "__main__":
    main()
    main(B=16, canvas=1280, persons=100, period=16, T=256, Dm=128, steps=32, clumps=10)
