"""Find the first frame / state field where the resident-tracker path of pb_step_seq differs from the serial path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import posebyte_b200 as pb

B, F = 6, 24
n_per_call = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scfg = pb.synth_config(canvas=640, persons=12, period=48, occlusion=1)
host = pb.synth_heads(scfg, 11, B, 0, F, frame_major=True)
heads = torch.from_numpy(host).cuda()
serial = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, max_age=4, fuse_stages=0)
res = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, max_age=4, pipeline_depth=5)
f = 0
bad = 0
while f < 60 and bad < 3:
    for i in range(n_per_call):
        serial.step(heads[(f + i) % F], f + i)
    res.step_seq(heads, f % F, n_per_call, f)
    f += n_per_call
    for b in range(B):
        s1, s2 = serial.get_state(b), res.get_state(b)
        diff = [k for k in s1 if s1[k].tobytes() != s2[k].tobytes()]
        if diff:
            bad += 1
            print("after frame", f - 1, "stream", b, "differs in", diff)
            for k in diff[:4]:
                a, c = np.asarray(s1[k]).ravel(), np.asarray(s2[k]).ravel()
                idx = np.nonzero(a != c)[0]
                print("   ", k, "first idx", idx[:8], "serial", a[idx[:4]], "resident", c[idx[:4]], "of", a.size)
            break
print("done, frames", f, "bad", bad)
