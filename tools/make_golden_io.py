"""Golden vectors for the rows either side of the path, recorded from the REFERENCE's own code on a B200
(oracle/_ref/libposebyte_ref.so = its unmodified preprocess.cu / hungarian.cu compiled for sm_100a):
PreprocessorCUDA::preprocess on small frames and LinearAssignmentCUDA::solve on seeded cost matrices.
Run on the GPU box:  python tools/make_golden_io.py gpurun_out/golden/ref_io_rows.npz
then copy the file to tests/golden/ (about 0.2 MB)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_py
from test_oracle_io_rows import make_frame

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden", "ref_io_rows.npz")
os.makedirs(os.path.dirname(out_path), exist_ok=True)
rng = np.random.default_rng(20261018)
d = {}
frames = [(96, 54, 64, 64), (54, 96, 64, 64), (64, 48, 64, 64), (37, 53, 48, 40), (20, 20, 64, 64), (64, 64, 64, 64)]
for i, (w, h, tw, th) in enumerate(frames):
    img = make_frame(rng, w, h)
    with ref_py.quiet():
        out, xf = ref_py.preprocess(img, tw, th)
    d[f"img{i}"] = img; d[f"target{i}"] = np.array([tw, th], np.int32); d[f"out{i}"] = out; d[f"xf{i}"] = xf
d["n_frames"] = np.int32(len(frames))
costs = [(3, 3, 0.5), (9, 11, 0.5), (10, 10, 0.5), (10, 10, 1.0), (50, 50, 0.5), (30, 12, 0.7), (12, 30, 0.3), (64, 20, 0.5), (100, 100, 1.0)]
for i, (R, C, thr) in enumerate(costs):
    cost = rng.uniform(0, 1, (R, C)).astype(np.float32)
    with ref_py.quiet():
        row, col, n = ref_py.assign_solve(cost, thr)
    d[f"cost{i}"] = cost; d[f"thr{i}"] = np.float32(thr); d[f"row{i}"] = row; d[f"col{i}"] = col; d[f"cnt{i}"] = np.int32(n)
d["n_costs"] = np.int32(len(costs))
np.savez_compressed(out_path, **d)
print("wrote", out_path, os.path.getsize(out_path), "bytes")
