"""Two handles fed the same heads on the serial path: after which frame, in which slab of the tracker
state (or in the kept detections) do they first differ?  (development aid: hunts data races)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb
B, F = int(os.environ.get("PB_B", "64")), int(os.environ.get("PB_PERIOD", "32"))
NSTEP = int(os.environ.get("PB_STEPS", "400"))
CANVAS, PERSONS = int(os.environ.get("PB_CANVAS", "640")), int(os.environ.get("PB_PERSONS", "20"))
kw = dict(max_tracks=int(os.environ.get("PB_T", "128")), max_detections=int(os.environ.get("PB_DM", "64")),
          max_age=int(os.environ.get("PB_MAX_AGE", "10")), keypoint_fetch=int(os.environ.get("PB_KPF", "0")))
scfg = pb.synth_config(canvas=CANVAS, persons=PERSONS, period=F, occlusion=int(os.environ.get("PB_OCC", "0")),
                       clumps=int(os.environ.get("PB_CLUMPS", "0")), kp_drop_prob=0.15 if os.environ.get("PB_CLUMPS") else 0.05)
d = torch.from_numpy(pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)).cuda()
A = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, **kw)
Bp = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, **kw)
print("config", dict(B=B, canvas=CANVAS, persons=PERSONS, period=F, steps=NSTEP, **kw))
T, Dm = A.T, A.Dm
slabs = [("poses", T * 51), ("vel", T * 34), ("scores", T), ("predicted", T * 51), ("tcent", T * 4), ("dcent", Dm * 4),
         ("cost", T * Dm), ("det_scores", Dm), ("states", T), ("ids", T), ("hits", T), ("ages", T), ("last_frame", T),
         ("active", T), ("pred_dirty", T), ("row_assign", T), ("col_assign", Dm), ("scalars", 4)]
def split(blob):
    a = np.frombuffer(blob, np.uint8)[24:].view(np.uint32)
    out, off = {}, 0
    for name, n in slabs:
        out[name] = a[off: off + B * n].reshape(B, n); off += B * n
    return out
found = 0
for f in range(NSTEP):
    A.postprocess(d[f % F]); Bp.postprocess(d[f % F])
    torch.cuda.synchronize()
    va, vb = A.device_views(), Bp.device_views()
    for b in range(B):
        ka, kb = A.get_kept(b), Bp.get_kept(b)
        for k in ka:
            if np.asarray(ka[k]).tobytes() != np.asarray(kb[k]).tobytes():
                print(f"frame {f} stream {b}: kept detections differ in {k}: {np.asarray(ka[k]).ravel()[:8]} vs {np.asarray(kb[k]).ravel()[:8]}"); found += 1
    A.tracker_update(f); Bp.tracker_update(f)
    sa, sb = split(A.state_save()), split(Bp.state_save())
    for name, _ in slabs:
        if not np.array_equal(sa[name], sb[name]):
            rows = np.nonzero((sa[name] != sb[name]).any(1))[0]
            b0 = rows[0]; idx = np.nonzero(sa[name][b0] != sb[name][b0])[0]
            print(f"frame {f}: slab {name} differs in streams {rows[:8].tolist()}; stream {b0} elements {idx[:8].tolist()} "
                  f"A {sa[name][b0][idx[:4]].tolist()} B {sb[name][b0][idx[:4]].tolist()}")
            found += 1
    if found:
        b0 = rows[0]
        print("stream", b0, "A ids", sa["ids"][b0][sa["active"][b0] == 1].tolist(), "\n          B ids", sb["ids"][b0][sb["active"][b0] == 1].tolist())
        print("A row_assign", sa["row_assign"][b0].view(np.int32)[:40].tolist(), "\nB row_assign", sb["row_assign"][b0].view(np.int32)[:40].tolist())
        print("A col_assign", sa["col_assign"][b0].view(np.int32)[:30].tolist(), "\nB col_assign", sb["col_assign"][b0].view(np.int32)[:30].tolist())
        print("A states", sa["states"][b0][:40].tolist(), "\nB states", sb["states"][b0][:40].tolist())
        print("A scalars", sa["scalars"][b0].tolist(), "B scalars", sb["scalars"][b0].tolist())
        break
print("frames compared", f + 1, "differences", found)
