"""Golden vectors from the REFERENCE ITSELF, run on a B200.

Drives the reference's own classes (GPUPostprocess, GPUTracker, LinearAssignmentCUDA,
KalmanFilterCUDA — its unmodified src/cuda/*.cu compiled for sm_100a into
oracle/_ref/libposebyte_ref.so) on seeded synthetic inputs and stores what they return.  The
fixtures are what pins the CPU restatement (oracle/): tests/test_golden.py replays the same
inputs through the restatement and compares (discrete outputs equal, floats within 1e-4
relative).  Inputs are not stored: they are regenerated from the recorded generator settings
(the generator is counter-based and bit-reproducible, tests/test_abi_and_host.py).

Run on a GPU box:   python tools/make_golden.py gpurun_out/golden
then copy gpurun_out/golden/*.npz into tests/golden/ and commit.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402

SCENARIOS = {
    # name: generator settings, streams, frames, tracker settings
    "std640": dict(synth=dict(canvas=640, persons=12, period=48), streams=[0, 3], frames=36,
                   trk=dict(max_tracks=128, max_detections=64, max_age=10, min_hits=3), conf=0.30, nms=0.65),
    "occl640": dict(synth=dict(canvas=640, persons=6, period=90, occlusion=1), streams=[7], frames=90,
                    trk=dict(max_tracks=128, max_detections=64, max_age=4, min_hits=3), conf=0.30, nms=0.65),
    "crowd1280": dict(synth=dict(canvas=1280, persons=40, period=16, clumps=4, kp_drop_prob=0.15), streams=[2], frames=8,
                      trk=dict(max_tracks=256, max_detections=128, max_age=10, min_hits=3), conf=0.30, nms=0.65),
}
FRAME_KEYS = ["states", "ids", "hits", "ages", "active", "row_assign", "col_assign"]   # per-frame discrete state
KMAX = 160         # padded kept-detection rows per frame
STATE_KEYS = ["poses", "vel", "scores", "states", "ids", "hits", "ages", "last_frame", "active", "row_assign",
              "col_assign", "scalars"]


def run_scenario(name, sc, pb, ref, torch):
    scfg = pb.synth_config(**sc["synth"])
    out = {}
    for s in sc["streams"]:
        heads = pb.synth_heads(scfg, s, 1, 0, sc["frames"], frame_major=False)[0]       # [F,56,N]
        d_heads = torch.from_numpy(heads).cuda()
        post = ref.Postprocess(1024, scfg.num_anchors)
        trk = ref.Tracker(new_track_thresh=sc["conf"], high_thresh=sc["conf"], low_thresh=sc["conf"] / 2, **sc["trk"])
        F = sc["frames"]
        Dm = sc["trk"]["max_detections"]
        nkeep = np.zeros(F, np.int32); nact = np.zeros(F, np.int32); ntrk = np.zeros(F, np.int32)
        kscores = np.zeros((F, KMAX), np.float32); kposes = np.zeros((F, KMAX, 51), np.float32)
        kboxes = np.zeros((F, KMAX, 4), np.float32)
        tracks = np.zeros((F, Dm), ref.TRACK_OUTPUT)
        T = sc["trk"]["max_tracks"]
        per_frame = {k: np.zeros((F, Dm if k == "col_assign" else T), np.int32) for k in FRAME_KEYS}
        for f in range(F):
            r = post.process(d_heads[f].data_ptr(), sc["conf"], sc["nms"])
            n = r["num_keep"]
            assert n <= KMAX, (name, s, f, n)
            nkeep[f] = n
            kscores[f, :n] = r["scores"]; kposes[f, :n] = r["poses"]; kboxes[f, :n] = r["bboxes"]
            nact[f] = trk.update(post.poses_dev(), post.scores_dev(), n, f)
            t = trk.get_tracks()
            ntrk[f] = len(t); tracks[f, :len(t)] = t
            st = trk.get_state()
            for k in FRAME_KEYS:
                per_frame[k][f] = st[k]
        st = trk.get_state()
        pre = f"{name}/s{s}/"
        out.update({pre + "num_keep": nkeep, pre + "num_active": nact, pre + "num_tracks": ntrk, pre + "kept_scores": kscores,
                    pre + "kept_poses": kposes, pre + "kept_bboxes": kboxes, pre + "tracks": tracks.view(np.uint8).reshape(F, Dm, 228)})
        for k in STATE_KEYS:
            out[pre + "state_" + k] = st[k]
        for k in FRAME_KEYS:
            out[pre + "frame_" + k] = per_frame[k]
        del post, trk
        print(f"{name} stream {s}: kept/frame {nkeep.mean():.1f}, tracks/frame {ntrk.mean():.1f}, final active {nact[-1]}, "
              f"max id {st['ids'].max()}", flush=True)
    return out


def run_auction(ref):
    out = {}
    rng = np.random.default_rng(20261018)
    shapes = [(5, 5), (20, 20), (50, 50), (30, 12), (12, 30), (128, 64), (64, 64)]
    for i, (R, Cc) in enumerate(shapes):
        cost = rng.uniform(0.0, 1.0, (R, Cc)).astype(np.float32)
        if i % 2 == 1:                         # sparse gating look-alike: most cells far away / locked
            far = rng.uniform(0, 1, (R, Cc)) < 0.7
            cost[far] = 1.0
            cost[rng.uniform(0, 1, (R, Cc)) < 0.05] = 1e9
        active = (rng.uniform(0, 1, R) < 0.8).astype(np.int32) if i >= 2 else None
        row, col = ref.auction(cost, active)
        out[f"auction/{i}/cost"] = cost
        out[f"auction/{i}/active"] = active if active is not None else np.ones(R, np.int32)
        out[f"auction/{i}/has_active"] = np.array([active is not None], np.int32)
        out[f"auction/{i}/row"] = row; out[f"auction/{i}/col"] = col
    return out


def run_kf3(ref):
    out = {}
    rng = np.random.default_rng(7)
    T = 16
    k = ref.KF3(T)
    dets0 = np.zeros((10, 17, 3), np.float32)
    dets0[:, :, 0] = rng.uniform(50, 600, (10, 17)); dets0[:, :, 1] = rng.uniform(50, 600, (10, 17))
    dets0[:, :, 2] = rng.uniform(0, 1, (10, 17)); dets0[0, :3, 2] = 0.0
    slots0 = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 12], np.int32)
    k.initiate(dets0.reshape(10, 51), slots0)
    steps = []
    snaps = []
    for it in range(6):
        k.predict(13, 0.9, 0.9)
        dets = dets0.copy()
        dets[:, :, :2] += rng.normal(0, 3.0, (10, 17, 2)).astype(np.float32) * (it + 1)
        dets[:, :, 2] = rng.uniform(0, 1, (10, 17)); dets[1, 5:9, 2] = 0.05
        m = np.array([[0, 0], [1, 1], [2, 3], [5, 4], [12, 9], [7, 7]][: 3 + it % 4], np.int32)
        k.update(dets.reshape(10, 51), m)
        means, diag, off = k.state()
        steps.append((dets.reshape(10, 51).copy(), m.copy()))
        snaps.append((means.copy(), diag.copy(), off))
    out["kf3/T"] = np.array([T], np.int32)
    out["kf3/dets0"] = dets0.reshape(10, 51); out["kf3/slots0"] = slots0
    for i, ((d, m), (mn, dg, off)) in enumerate(zip(steps, snaps)):
        out[f"kf3/{i}/dets"] = d; out[f"kf3/{i}/matches"] = m
        out[f"kf3/{i}/means"] = mn; out[f"kf3/{i}/diag"] = dg; out[f"kf3/{i}/max_offdiag"] = np.array([off], np.float32)
    return out


def main():
    outdir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(outdir, exist_ok=True)
    import torch
    import posebyte_b200 as pb
    import ref_py as ref
    assert torch.cuda.is_available() and ref.available()
    torch.zeros(1).cuda()
    allv = {}
    for name, sc in SCENARIOS.items():
        allv.update(run_scenario(name, sc, pb, ref, torch))
    allv.update(run_auction(ref))
    allv.update(run_kf3(ref))
    np.savez_compressed(os.path.join(outdir, "ref_b200.npz"), **allv)
    meta = {"scenarios": SCENARIOS, "kmax": KMAX, "device": torch.cuda.get_device_name(0),
            "what": "outputs of the reference's own src/cuda/*.cu (unmodified, compiled for sm_100a) on seeded synthetic inputs",
            "generator_seed": hex(pb.synth_config().seed)}
    with open(os.path.join(outdir, "ref_b200.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", outdir, len(allv), "arrays")
    selfcheck(os.path.join(outdir, "ref_b200.npz"), pb)


def selfcheck(path, pb):
    """Replay through the CPU checker right away so a generation run also says whether it pins."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_util as gu
    import oracle_py as orc
    G = np.load(path)
    for name, sc in SCENARIOS.items():
        for s in sc["streams"]:
            for replay in (False, True):
                bad, ident = gu.check_stream_against(pb, orc, sc, name, s, gu.golden_stream(G, name, s), replay=replay)
                print(f"selfcheck {name} stream {s} replay={replay}: {'OK' if not bad else 'MISMATCH'} ids_identical={ident}")
                for b in bad[:8]:
                    print("   ", b)
    i = 0
    while f"auction/{i}/cost" in G.files:
        act = G[f"auction/{i}/active"] if G[f"auction/{i}/has_active"][0] else None
        row, col = orc.auction(G[f"auction/{i}/cost"], act)
        print(f"selfcheck auction {i} {G[f'auction/{i}/cost'].shape}: rows_equal={np.array_equal(row, G[f'auction/{i}/row'])} "
              f"cols_equal={np.array_equal(col, G[f'auction/{i}/col'])} matched={(row >= 0).sum()}")
        i += 1
    k = orc.KF3(int(G["kf3/T"][0]))
    k.initiate(G["kf3/dets0"], G["kf3/slots0"])
    i = 0
    while f"kf3/{i}/dets" in G.files:
        k.predict(13, 0.9, 0.9)
        k.update(G[f"kf3/{i}/dets"], G[f"kf3/{i}/matches"])
        m, d = k.state()
        print(f"selfcheck kf3 step {i}: means rel {gu.max_rel(m, G[f'kf3/{i}/means']):.2e} diag rel {gu.max_rel(d, G[f'kf3/{i}/diag']):.2e} "
              f"ref max offdiag {G[f'kf3/{i}/max_offdiag'][0]}")
        i += 1


if __name__ == "__main__":
    main()
