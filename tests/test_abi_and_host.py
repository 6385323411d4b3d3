"""The C-ABI library: loads on a GPU-less host, exports every symbol the header declares,
fails loudly (no CPU path) when asked to compute without a device; host-side logic."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    txt = open(os.path.join(ROOT, "include", "posebyte_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+|launchPoseNMS)\s*\(", txt)))


def test_library_exports_every_declared_symbol(pb):
    L = pb.lib()
    names = header_functions()
    assert "pb_step" in names and "launchPoseNMS" in names and len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/posebyte_b200.h but not exported"
    assert sorted(pb.ABI_SYMBOLS) == names


def test_struct_layouts_match_header(pb):
    assert C.sizeof(pb.PbConfig) == 18 * 4
    assert pb.default_config().pipeline_depth == 1
    assert C.sizeof(pb.PbTiming) == 10 * 8 + 8
    assert pb.TRACK_OUTPUT.itemsize == 228 and pb.POSE_DETECTION.itemsize == 224
    cfg = pb.default_config()
    assert (cfg.num_anchors, cfg.max_candidates, cfg.max_keep, cfg.max_tracks, cfg.max_detections) == (8400, 1024, 256, 128, 64)
    assert (cfg.max_age, cfg.min_hits, cfg.gating_enabled) == (10, 3, 1)
    assert abs(cfg.new_track_thresh - 0.30) < 1e-7 and abs(cfg.match_threshold - 0.5) < 1e-7


def test_no_cpu_fallback(pb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pb.PbError) as e:
        pb.Pipeline(num_streams=2)
    assert e.value.status == pb.PB_ERR_NO_DEVICE
    assert "no CPU path" in str(e.value)


def test_invalid_config_is_rejected(pb):
    with pytest.raises(pb.PbError) as e:
        pb.Pipeline(num_streams=0)
    assert e.value.status == pb.PB_ERR_INVALID
    with pytest.raises(pb.PbError) as e:
        pb.Pipeline(num_anchors=100000)
    assert e.value.status == pb.PB_ERR_UNSUPPORTED


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "yolo-pose-cpp_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_py" not in txt and "posebyte_oracle" not in txt and "orc_" not in txt, os.path.join(dp, f)
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "orc_" not in open(p).read().replace("orc_tracker_get_state", ""), p


def test_shard_partition(pb):
    for total, world in [(64, 1), (64, 8), (1024, 8), (10, 4), (3, 8)]:
        shards = [pb.Shard(r, world, total) for r in range(world)]
        covered = [s for sh in shards for s in sh.streams()]
        assert covered == list(range(total))
        assert max(sh.count for sh in shards) - min(sh.count for sh in shards) <= 1


def test_words_checksum_matches_oracle(pb, orc):
    cfg = pb.synth_config(canvas=640, persons=6, period=16)
    heads = pb.synth_heads(cfg, 0, 1, 0, 6, frame_major=False)
    r = orc.run_streams(heads, False)
    t = orc.Tracker()
    h, pos = 0, 0
    for f in range(6):
        p = orc.postprocess(heads[0, f])
        t.update(p["poses"], p["scores"], f)
        out = t.get_tracks()
        words = np.concatenate([np.array([p["num_keep"], len(out)], np.uint32), p["keep_anchors"].view(np.uint32),
                                np.frombuffer(out.tobytes(), np.uint32)])
        h = (h + pb.words_checksum(words, pos)) % (1 << 64)
        pos += words.size
    assert h == int(r["hashes"][0])


def test_synth_is_deterministic_and_plausible(pb, orc):
    cfg = pb.synth_config(canvas=640, persons=20, period=32)
    a = pb.synth_heads(cfg, 3, 2, 5, 2, frame_major=True, threads=1)
    b = pb.synth_heads(cfg, 3, 2, 5, 2, frame_major=True, threads=4)
    assert a.tobytes() == b.tobytes() and a.shape == (2, 2, 56, 8400)
    c = pb.synth_heads(cfg, 4, 1, 5 + 32, 1)          # period wrap: frame f+period == frame f
    assert c[0, 0].tobytes() == a[0, 1].tobytes()
    n = orc.decode(a[0, 0], 0.30)["num"]
    assert 20 * 4 <= n <= 20 * 9
    assert pb.num_anchors_for(1280) == 33600
