"""Stage-by-stage GPU-vs-checker diagnostics (development aid; run under gpurun)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import posebyte_b200 as pb, oracle_py as orc
from helpers import compare_state

def run(B, F, canvas, persons, period, occlusion=0, clumps=0, T=128, Dm=64, max_age=10, verbose=True):
    scfg = pb.synth_config(canvas=canvas, persons=persons, period=period, occlusion=occlusion, clumps=clumps,
                           kp_drop_prob=0.15 if clumps else 0.05)
    heads = pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)
    pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, max_tracks=T, max_detections=Dm, max_age=max_age)
    trk = [orc.Tracker(max_tracks=T, max_detections=Dm, max_age=max_age) for _ in range(B)]
    d_heads = torch.from_numpy(heads).cuda()
    nbad = 0
    for f in range(F):
        pipe.step(d_heads[f], f)
        torch.cuda.synchronize()
        for b in range(B):
            ref = orc.postprocess(heads[f, b])
            got = pipe.get_kept(b)
            ok_post = (got["num_cand"] == ref["num_cand"] and np.array_equal(got["keep_anchors"], ref["keep_anchors"])
                       and np.array_equal(got["keep_slots"], ref["keep_slots"])
                       and got["poses"].tobytes() == ref["poses"].tobytes() and got["bboxes"].tobytes() == ref["bboxes"].tobytes()
                       and got["scores"].tobytes() == ref["scores"].tobytes())
            if not ok_post:
                nbad += 1
                if verbose and nbad < 6:
                    print(f"POST MISMATCH f={f} b={b} cand {got['num_cand']}/{ref['num_cand']} keep {got['num_keep']}/{ref['num_keep']}")
                    print("  got", got["keep_anchors"][:24]); print("  ref", ref["keep_anchors"][:24])
            trk[b].update(ref["poses"], ref["scores"], f)
            rs, gs = trk[b].get_state(), pipe.get_state(b)
            D = int(rs["scalars"][2])
            bad = compare_state(gs, rs, D, T, where=f"f={f} b={b}")
            rt, gt = trk[b].get_tracks(), pipe.get_tracks(b)
            if rt.tobytes() != gt.tobytes():
                bad.append(f"f={f} b={b} outputs differ n {len(gt)}/{len(rt)} ids {gt['track_id'][:10]} / {rt['track_id'][:10]}")
            if bad:
                nbad += 1
                if verbose and nbad < 8:
                    for s in bad[:6]: print("TRK", s)
    print(f"run B={B} F={F} canvas={canvas} P={persons} occl={occlusion} clumps={clumps} T={T} Dm={Dm}: mismatching (frame,stream) = {nbad}")
    t = pipe.timing()
    print("  timing us/frame:", {n: getattr(t, n) / max(t.frame_count, 1) for n, _ in t._fields_ if n != "frame_count"})
    return nbad

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    tot = 0
    tot += run(2, 30, 640, 20, 64)
    tot += run(2, 60, 640, 20, 64, occlusion=1, max_age=5)
    tot += run(1, 12, 1280, 100, 32, clumps=10, T=256, Dm=128)
    print("TOTAL MISMATCH", tot)
    # quick timing of the step at B=64
    B = 64
    scfg = pb.synth_config(canvas=640, persons=20, period=32)
    heads = pb.synth_heads(scfg, 0, B, 0, 8, frame_major=True)
    d = torch.from_numpy(heads).cuda()
    pipe = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors)
    for f in range(16): pipe.step(d[f % 8], f)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 64
    for f in range(16, 16 + n): pipe.step(d[f % 8], f)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"B=64 step: {ms*1000:.1f} us/step -> {B/ms*1000:.0f} stream-frames/s")
    e0.record()
    for f in range(n): pipe.postprocess(d[f % 8])
    e1.record(); torch.cuda.synchronize()
    print(f"  postprocess only: {e0.elapsed_time(e1)/n*1000:.1f} us")
    e0.record()
    for f in range(n): pipe.tracker_update(100 + f)
    e1.record(); torch.cuda.synchronize()
    print(f"  tracker only: {e0.elapsed_time(e1)/n*1000:.1f} us")
    t = pipe.timing()
    print("  timing us/frame:", {nm: round(getattr(t, nm) / max(t.frame_count, 1), 2) for nm, _ in t._fields_ if nm != "frame_count"})
