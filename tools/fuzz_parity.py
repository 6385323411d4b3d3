"""Randomised parity run: random workloads (persons, clumps, occlusion, table sizes, life-cycle parameters, gating) through pb_step
on the default path and through the checker, every frame compared (records, active counts, full state every few frames).
PB_FUZZ_CASES cases (default 40), seed PB_FUZZ_SEED."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import posebyte_b200 as pb
import oracle_py as orc
from helpers import compare_state
rng = np.random.default_rng(int(os.environ.get("PB_FUZZ_SEED", "1")))
N = int(os.environ.get("PB_FUZZ_CASES", "40"))
bad = 0
t0 = time.time()
for case in range(N):
    crowd = rng.random() < 0.4
    canvas = 1280 if crowd else 640
    persons = int(rng.integers(30, 130)) if crowd else int(rng.integers(1, 45))
    clumps = int(rng.integers(2, 12)) if crowd and rng.random() < 0.8 else 0
    T, Dm = ((256, 128) if crowd else [(128, 64), (64, 32), (32, 32), (128, 128)][int(rng.integers(0, 4))])
    kw = dict(max_age=int(rng.integers(1, 12)), min_hits=int(rng.integers(1, 4)), gating_enabled=int(rng.random() < 0.8))
    occ = int(rng.random() < 0.5)
    B, F = int(rng.integers(1, 4)), int(rng.integers(6, 20))
    fuse = int(rng.integers(0, 2)) if not crowd else 0
    depth = int(rng.choice([1, 5]))
    cfg = pb.synth_config(canvas=canvas, persons=persons, period=int(rng.integers(8, 64)), occlusion=occ, clumps=clumps,
                          kp_drop_prob=float(rng.choice([0.05, 0.15, 0.3])), seed=int(rng.integers(1, 1 << 30)))
    heads = pb.synth_heads(cfg, int(rng.integers(0, 100)), B, 0, F, frame_major=True)
    pipe = pb.Pipeline(num_streams=B, num_anchors=cfg.num_anchors, max_tracks=T, max_detections=Dm, fuse_stages=fuse, pipeline_depth=depth, **kw)
    trk = [orc.Tracker(max_tracks=T, max_detections=Dm, **kw) for _ in range(B)]
    d = torch.from_numpy(heads).cuda()
    ok = True
    seq = depth > 1 and rng.random() < 0.5          # pb_step_seq in calls of random length (the resident-tracker path where it applies)
    f = 0
    while f < F:
        n = int(rng.integers(1, 8)) if seq else 1
        n = min(n, F - f)
        if seq:
            pipe.step_seq(d, f, n, f)
        else:
            pipe.step(d[f], f)
        pipe.join(); torch.cuda.synchronize()
        for g in range(f, f + n - 1):               # the checker catches up to the last frame of the call
            for b in range(B):
                ref = orc.postprocess(heads[g, b]); trk[b].update(ref["poses"], ref["scores"], g)
        f += n - 1
        na = pipe.get_num_active()
        for b in range(B):
            ref = orc.postprocess(heads[f, b])
            ra = trk[b].update(ref["poses"], ref["scores"], f)
            rt, gt = trk[b].get_tracks(), pipe.get_tracks(b)
            if na[b] != ra or gt.tobytes() != rt.tobytes():
                ok = False; print("MISMATCH records", case, f, b); break
            if f % 3 == 2 or f == F - 1:
                rs = trk[b].get_state()
                diff = compare_state(pipe.get_state(b), rs, int(rs["scalars"][2]), T, f"f{f} b{b}")
                if diff:
                    ok = False; print("MISMATCH state", case, diff[:3]); break
        if not ok: break
        f += 1
    print(f"case {case}: seq {int(seq)} canvas {canvas} persons {persons} clumps {clumps} T {T} Dm {Dm} B {B} F {F} occ {occ} fuse {fuse} depth {depth} {kw} -> {'ok' if ok else 'FAIL'}", flush=True)
    bad += 0 if ok else 1
    pipe.close()
print("fuzz", "FAILED" if bad else "ok", N, "cases", f"{time.time() - t0:.0f} s")
sys.exit(1 if bad else 0)
