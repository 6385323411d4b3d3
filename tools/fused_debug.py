"""Which path deviates?  non-fused serial vs fused serial vs fused pipelined on the failing test scenario, frame by frame."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
B, F, depth = int(os.environ.get("PB_B", "64")), 48, int(os.environ.get("PB_DEPTH", "3"))
scfg = pb.synth_config(canvas=640, persons=12, period=48, occlusion=1)
heads = torch.from_numpy(pb.synth_heads(scfg, 7, B, 0, F, frame_major=True)).cuda()
kw = dict(num_streams=B, num_anchors=scfg.num_anchors, max_age=4)
ref = pb.Pipeline(fuse_stages=0, **kw)
fs = pb.Pipeline(**kw)
fp = pb.Pipeline(pipeline_depth=depth, **kw)
SYNC = int(os.environ.get("PB_SYNC_EVERY", "1"))
for f in range(70):
    ref.step(heads[f % F], f); fs.step(heads[f % F], f); fp.step(heads[f % F], f)
    if (f + 1) % SYNC == 0 or f == 69:
        o0, c0 = ref.get_tracks_all(); o1, c1 = fs.get_tracks_all(); o2, c2 = fp.get_tracks_all()
        bad1 = [b for b in range(B) if c0[b] != c1[b] or o0[b, :c0[b]].tobytes() != o1[b, :c1[b]].tobytes()]
        bad2 = [b for b in range(B) if c0[b] != c2[b] or o0[b, :c0[b]].tobytes() != o2[b, :c2[b]].tobytes()]
        if bad1 or bad2:
            print(f"frame {f}: fused serial differs on {bad1[:10]} ({len(bad1)}), fused pipelined on {bad2[:10]} ({len(bad2)})")
            b = (bad1 + bad2)[0]
            k0, k1, k2 = ref.get_kept(b), fs.get_kept(b), fp.get_kept(b)
            print("  kept", k0["num_keep"], k1["num_keep"], k2["num_keep"], "cand", k0["num_cand"], k1["num_cand"], k2["num_cand"],
                  "anchors equal", np.array_equal(k0["keep_anchors"], k1["keep_anchors"]), np.array_equal(k0["keep_anchors"], k2["keep_anchors"]))
            s0, s1, s2 = ref.get_state(b), fs.get_state(b), fp.get_state(b)
            for k in s0:
                if s0[k].tobytes() != s1[k].tobytes() or s0[k].tobytes() != s2[k].tobytes():
                    print("   state", k, "serial-fused equal:", s0[k].tobytes() == s1[k].tobytes(), "piped equal:", s0[k].tobytes() == s2[k].tobytes())
            break
else:
    print("all 70 frames equal (sync every", SYNC, ")")
