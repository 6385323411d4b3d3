"""The reference's C++ class API (include/cuda/*.h, header-only views of the C ABI) used the way
the reference's main.cpp:132-140,207-224 uses it: tests/cpp/shim_frame_loop.cpp is compiled with
g++ (host C++17, no nvcc) against the shared library and run."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "build", "shim_frame_loop")


def build_demo():
    import __graft_entry__ as g
    g.build_cpp_demo()
    return EXE


def test_shims_compile_as_plain_cpp17_and_fail_loudly_without_gpu(tmp_path):
    import torch
    exe = build_demo()
    assert os.path.exists(exe)
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    heads = tmp_path / "h.bin"
    np.zeros((1, 56, 64), np.float32).tofile(heads)
    r = subprocess.run([exe, str(heads), "1", "64", "0.3", "10"], capture_output=True, text=True)
    assert r.returncode == 1
    assert "no CPU path" in r.stderr           # pb_create refuses: there is no host fallback


@pytest.mark.gpu
def test_reference_frame_loop_through_shims_equals_checker(pb, orc, cuda, tmp_path):
    exe = build_demo() if not os.path.exists(EXE) else EXE
    F, conf, max_age = 24, 0.30, 10
    scfg = pb.synth_config(canvas=640, persons=14, period=64)
    heads = pb.synth_heads(scfg, 5, 1, 0, F, frame_major=False)[0]       # [F,56,N]
    path = tmp_path / "heads.bin"
    heads.tofile(path)
    r = subprocess.run([exe, str(path), str(F), str(scfg.num_anchors), str(conf), str(max_age)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    trk = orc.Tracker(new_track_thresh=conf, high_thresh=conf, low_thresh=conf / 2, max_age=max_age)
    total = 0
    for f in range(F):
        k = orc.postprocess(heads[f], conf, 0.65)
        na = trk.update(k["poses"], k["scores"], f)
        t = trk.get_tracks()
        want = f"frame {f}: kept {k['num_keep']} active {na} tracks {len(t)} :" + "".join(
            f" {int(x['track_id'])}:{x['score'].view(np.uint32):08x}:{x['keypoints'][0, 0].view(np.uint32):08x}" for x in t)
        assert lines[f].strip() == want.strip(), (f, lines[f], want)
        total += len(t)
    assert total > 100
    rest = {ln.split()[0]: ln for ln in lines[F:]}
    assert rest["raw"].startswith(f"raw {k['num_keep']} first_score {k['scores'][0].view(np.uint32):08x}")
    # NMSCuda::apply on the last frame's detections, each duplicated with a 2 px shift (built
    # exactly as the C++ side builds them), against the checker's host-NMS restatement
    n = k["num_keep"]
    dets = np.zeros(2 * n, orc.POSE_DETECTION)
    for i in range(n):
        for dup in range(2):
            d = dets[2 * i + dup]
            d["bbox"] = k["bboxes"][i] + np.float32(2.0 * dup)
            d["score"] = k["scores"][i] - np.float32(0.01) * np.float32(dup)
            kp = k["poses"][i].reshape(17, 3).copy()
            kp[:, 0] += np.float32(2.0 * dup)
            d["keypoints"] = kp
    want_keep = orc.nms_legacy(dets, 0.65, 0.25)
    assert rest["nms_apply"].strip() == (f"nms_apply {len(want_keep)} of {2 * n} :" + "".join(f" {i}" for i in want_keep)).strip()
    assert all(i % 2 == 0 for i in want_keep) and len(want_keep) >= n - 2      # every shifted copy loses to its original
    assert rest["auction"] == "auction 0 1 2"
    # LinearAssignmentCUDA::solve on the same 3x3 table, threshold 0.25: greedy rows (9 cells < 100)
    cost3 = np.array([[0.1, 0.9, 0.8], [0.7, 0.2, 0.9], [0.9, 0.8, 0.3]], np.float32)
    wr, _, wn = orc.assign_legacy(cost3, 0.25)
    assert rest["solve"] == f"solve {wn} : {wr[0]} {wr[1]} {wr[2]}", rest["solve"]
    assert rest["accessors"] == "accessors ok"        # getCostMatrixDevice / getRow..Device / getPricesDevice / getRowMatchedDevice / getSigmasDevice ...
    # PreprocessorCUDA on a 4x2 frame: scale 2, bars of 2 rows above and below
    img = (10 * np.arange(24)).astype(np.uint8).reshape(2, 4, 3)
    lb, xf = orc.letterbox(img, 8, 8)
    assert rest["letterbox"] == f"letterbox scale {xf[0]:.3f} pad {int(xf[2])} {int(xf[3])} corner {lb[0, 0, 0]:.4f} centre_r {lb[0, 2, 0]:.4f}", rest["letterbox"]
    # KF3: initiate at (100, 200), one predict with zero velocity keeps the position; confidence 1.0;
    # variance 10 (conf > 0) + process noise 1; off-diagonal 0
    assert rest["kf3"] == "kf3 nose 100.000 200.000 conf 1.0 var_x 11.000 offdiag 0.000", rest["kf3"]
