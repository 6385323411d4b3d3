"""Whole path at BASELINE config sizes: pb_step over many streams against the CPU checker's
stream runner through size-independent properties (a checksum of every frame's kept anchors and
TrackOutput records), shard invariance, and the host-buffer entry point."""
import numpy as np
import pytest

from helpers import valid_records

pytestmark = pytest.mark.gpu


def gpu_hashes(pb, torch, heads, **cfg):
    F, B, _, N = heads.shape
    pipe = pb.Pipeline(num_streams=B, num_anchors=N, **cfg)
    h = [0] * B
    pos = [0] * B
    total = 0
    for f in range(F):
        pipe.step(torch.from_numpy(heads[f]).cuda(), f)
        outs, counts = pipe.get_tracks_all()
        for b in range(B):
            k = pipe.get_kept(b)
            words = np.concatenate([np.array([k["num_keep"], counts[b]], np.uint32), k["keep_anchors"].view(np.uint32),
                                    np.frombuffer(outs[b, : counts[b]].tobytes(), np.uint32)])
            h[b] = (h[b] + pb.words_checksum(words, pos[b])) % (1 << 64)
            pos[b] += words.size
            total += int(counts[b])
    return np.array(h, np.uint64), total


def test_config2_64_streams_checksum(pb, orc, cuda):
    """Config 2 (64 concurrent 640x640 streams, 20 persons, max-age 10), 24 frames."""
    cfg = pb.synth_config(canvas=640, persons=20, period=24)
    heads = pb.synth_heads(cfg, 0, 64, 0, 24, frame_major=True)
    got, total = gpu_hashes(pb, cuda, heads)
    ref = orc.run_streams(heads, True, threads=8)
    assert np.array_equal(got, ref["hashes"]) and total == ref["tracks_total"] > 10000


def test_config4_max_age_30_with_occlusion_and_shard_invariance(pb, orc, cuda):
    """Config 4 semantics (max-age 30, occlusion gaps) on one shard of 16 streams; a second handle
    holding only streams 8..15 must reproduce their results (streams are independent)."""
    cfg = pb.synth_config(canvas=640, persons=20, period=120, occlusion=1)
    heads = pb.synth_heads(cfg, 0, 16, 0, 120, frame_major=True)
    got, total = gpu_hashes(pb, cuda, heads, max_age=30)
    ref = orc.run_streams(heads, True, threads=8, max_age=30)
    assert np.array_equal(got, ref["hashes"]) and total == ref["tracks_total"]
    shard, _ = gpu_hashes(pb, cuda, np.ascontiguousarray(heads[:, 8:]), max_age=30)
    assert np.array_equal(shard, got[8:])


def test_config3_crowd_checksum(pb, orc, cuda):
    cfg = pb.synth_config(canvas=1280, persons=100, period=12, clumps=10, kp_drop_prob=0.15)
    heads = pb.synth_heads(cfg, 0, 4, 0, 12, frame_major=True)
    got, total = gpu_hashes(pb, cuda, heads, max_tracks=256, max_detections=128)
    ref = orc.run_streams(heads, True, threads=4, max_tracks=256, max_detections=128)
    assert np.array_equal(got, ref["hashes"]) and total == ref["tracks_total"] > 1000


def test_step_host_equals_device_path(pb, orc, cuda):
    torch = cuda
    cfg = pb.synth_config(canvas=640, persons=10, period=16)
    heads = pb.synth_heads(cfg, 0, 4, 0, 10, frame_major=True)
    a, b = pb.Pipeline(num_streams=4), pb.Pipeline(num_streams=4)
    for f in range(10):
        out_h, cnt_h = a.step_host(heads[f], f)
        b.step(torch.from_numpy(heads[f]).cuda(), f)
        out_d, cnt_d = b.get_tracks_all()
        assert np.array_equal(cnt_h, cnt_d)
        for s in range(4):
            assert out_h[s, : cnt_h[s]].tobytes() == out_d[s, : cnt_d[s]].tobytes()
    assert cnt_h.sum() > 20
    # page-locked buffers: the input is read in place over PCIe, the results land in `out`/`counts`
    c = pb.Pipeline(num_streams=4)
    pinned = torch.from_numpy(heads).pin_memory()
    out_p = torch.zeros(4 * c.Dm * 228, dtype=torch.uint8).pin_memory()
    cnt_p = torch.zeros(4, dtype=torch.int32).pin_memory()
    out_np = out_p.numpy().view(pb.TRACK_OUTPUT).reshape(4, c.Dm)
    d = pb.Pipeline(num_streams=4)
    for f in range(10):
        c.step_host(pinned[f].numpy(), f, out=out_np, counts=cnt_p.numpy())
        d.step(torch.from_numpy(heads[f]).cuda(), f)
        out_d, cnt_d = d.get_tracks_all()
        assert np.array_equal(cnt_p.numpy(), cnt_d)
        for s in range(4):
            assert out_np[s, : cnt_d[s]].tobytes() == out_d[s, : cnt_d[s]].tobytes()


def test_launch_counter_counts_real_kernels(pb, cuda):
    torch = cuda
    heads = torch.zeros(2, 56, 8400, device="cuda")
    for fuse, per_step in ((1, 2), (0, 3)):         # decode+gather and the fused NMS+tracker kernel / decode+gather, NMS, tracker
        pipe = pb.Pipeline(num_streams=2, fuse_stages=fuse)
        before = pb.launch_count()
        for f in range(5):
            pipe.step(heads, f)
        torch.cuda.synchronize()
        assert pb.launch_count() - before == 5 * per_step


@pytest.mark.gpu
@pytest.mark.parametrize("depth,fuse", [(2, 0), (3, 0), (2, 1), (3, 1), (5, 1), (8, 1)])
def test_pipelined_steps_equal_serial_steps(pb, orc, cuda, depth, fuse):
    """pipeline_depth > 1 overlaps consecutive steps on internal streams; every result must be
    the one the serial path (and hence the checker) produces."""
    torch = cuda
    B, F = 8, 40
    scfg = pb.synth_config(canvas=640, persons=14, period=64, occlusion=1)
    host = pb.synth_heads(scfg, 100, B, 0, F, frame_major=True)
    heads = torch.from_numpy(host).cuda()
    serial = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, max_age=5, fuse_stages=0)
    piped = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, max_age=5, pipeline_depth=depth, fuse_stages=fuse)
    for f in range(F):                       # no synchronisation between steps: real overlap
        serial.step(heads[f], f)
        piped.step(heads[f], f)
    piped.join()
    torch.cuda.synchronize()
    o1, c1 = serial.get_tracks_all(); o2, c2 = piped.get_tracks_all()
    assert np.array_equal(c1, c2) and c1.sum() > 0
    for b in range(B):
        assert o1[b, :c1[b]].tobytes() == o2[b, :c2[b]].tobytes()
        s1, s2 = serial.get_state(b), piped.get_state(b)
        for k in s1:
            assert s1[k].tobytes() == s2[k].tobytes(), (b, k)
        k1, k2 = serial.get_kept(b), piped.get_kept(b)
        assert np.array_equal(k1["keep_anchors"], k2["keep_anchors"])
    # and with a read-back after every step against the checker
    piped.reset()
    trk = [orc.Tracker(max_age=5) for _ in range(2)]
    for f in range(12):
        piped.step(heads[f], f)
        for b in range(2):
            ref = orc.postprocess(host[f, b])
            trk[b].update(ref["poses"], ref["scores"], f)
            assert np.array_equal(piped.get_kept(b)["keep_anchors"], ref["keep_anchors"])
            assert piped.get_tracks(b).tobytes() == trk[b].get_tracks().tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("B,depth,fuse", [(64, 3, 0), (74, 6, 0), (96, 3, 0), (64, 4, 1), (64, 8, 1), (96, 5, 1), (150, 5, 1)])
def test_overlapping_tracker_launches_keep_every_stream_in_frame_order(pb, cuda, B, depth, fuse):
    """Consecutive tracker launches overlap (two streams, per-stream sequence flags) when two grids fit the
    device (2*B <= 148); streams with occlusions run into the auction's iteration limit and fall behind the
    others.  120 steps without a join, a reset and a stand-alone tracker update in between: states and
    records must equal the serial handle's bit for bit (B = 96: the non-overlapping fallback)."""
    torch = cuda
    F = 48
    scfg = pb.synth_config(canvas=640, persons=12, period=48, occlusion=1)
    heads = torch.from_numpy(pb.synth_heads(scfg, 7, B, 0, F, frame_major=True)).cuda()
    serial = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, max_age=4, fuse_stages=0)
    piped = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, max_age=4, pipeline_depth=depth, fuse_stages=fuse)
    def compare(tag):
        o1, c1 = serial.get_tracks_all(); o2, c2 = piped.get_tracks_all()
        assert np.array_equal(c1, c2) and c1.sum() > 0, tag
        assert valid_records(o1, c1) == valid_records(o2, c2), tag
        for b in (0, B // 2, B - 1):
            s1, s2 = serial.get_state(b), piped.get_state(b)
            for k in s1:
                assert s1[k].tobytes() == s2[k].tobytes(), (tag, b, k)
    for f in range(70):
        serial.step(heads[f % F], f); piped.step(heads[f % F], f)
    compare("after 70 steps")
    serial.reset(); piped.reset()
    for f in range(50):
        serial.step(heads[(f + 5) % F], f); piped.step(heads[(f + 5) % F], f)
        if f == 20:                                   # a stage-level call in the middle of the pipelined steps
            serial.postprocess(heads[3]); piped.postprocess(heads[3])
            serial.tracker_update(100); piped.tracker_update(100)
    compare("after reset + 50 steps")


@pytest.mark.gpu
@pytest.mark.parametrize("fuse", [0, 1])
def test_submit_host_pipelined_equals_serial_device_path(pb, cuda, fuse):
    """pb_submit_host / pb_wait: page-locked heads read in place, lazy NMS sweep, records copied into
    per-step page-locked buffers, consecutive steps overlapping; results = the plain device path."""
    torch = cuda
    B, F = 6, 24
    scfg = pb.synth_config(canvas=640, persons=12, period=64)
    host = pb.synth_heads(scfg, 40, B, 0, F, frame_major=True)
    pinned = torch.from_numpy(host).pin_memory()
    ref = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, fuse_stages=0)
    piped = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, pipeline_depth=4, fuse_stages=fuse)
    outs = torch.zeros(F, B * piped.Dm * 228, dtype=torch.uint8).pin_memory()
    cnts = torch.zeros(F, B, dtype=torch.int32).pin_memory()
    for f in range(F):
        piped.submit_host(pinned[f].numpy(), f, outs[f].numpy(), cnts[f].numpy())
    piped.wait()
    d = torch.from_numpy(host).cuda()
    total = 0
    for f in range(F):
        ref.step(d[f], f)
        o, c = ref.get_tracks_all()
        assert np.array_equal(c, cnts[f].numpy()), f
        got = outs[f].numpy().view(pb.TRACK_OUTPUT).reshape(B, piped.Dm)
        for b in range(B):
            assert got[b, : c[b]].tobytes() == o[b, : c[b]].tobytes(), (f, b)
        total += int(c.sum())
    assert total > 500
    with pytest.raises(pb.PbError):                       # pageable buffers are refused
        piped.submit_host(host[0], 0, outs[0].numpy(), cnts[0].numpy())


@pytest.mark.gpu
def test_long_run_equals_checker_and_pipelined_lanes(pb, orc, cuda):
    """Long runs (32 streams x 300 frames, heads rotating with period 20).  Regression test for a race
    that needed ~5000 stream-frames to show: the tracker's active-row list was compacted through an
    atomic counter, so its order — which breaks ties between equal auction bids (lowest row) — depended
    on which warp arrived first.  Serial path, a second serial handle, the pipelined path with three
    lanes (pipeline_depth 5) and the CPU checker must agree bit for bit on every stream's final state."""
    torch = cuda
    B, F, STEPS = 32, 20, 300
    scfg = pb.synth_config(canvas=640, persons=20, period=F)
    host = pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)
    heads = torch.from_numpy(host).cuda()
    a = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, fuse_stages=0)
    a2 = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, fuse_stages=1)
    p = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, pipeline_depth=5, fuse_stages=0)
    pf = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, pipeline_depth=6, fuse_stages=1)
    for f in range(STEPS):
        a.step(heads[f % F], f); a2.step(heads[f % F], f); p.step(heads[f % F], f); pf.step(heads[f % F], f)
    p.join(); pf.join(); torch.cuda.synchronize()
    oa, ca = a.get_tracks_all(); o2, c2 = a2.get_tracks_all(); op, cp = p.get_tracks_all(); of, cf = pf.get_tracks_all()
    assert np.array_equal(ca, c2) and valid_records(oa, ca) == valid_records(o2, c2), "serial handles (separate kernels, fused kernel) differ"
    assert np.array_equal(ca, cp) and valid_records(oa, ca) == valid_records(op, cp), "pipelined (lanes) differs from serial"
    assert np.array_equal(ca, cf) and valid_records(oa, ca) == valid_records(of, cf), "pipelined fused kernel (chain hand-off) differs from serial"
    assert a.state_save()[24:] == a2.state_save()[24:] == p.state_save()[24:] == pf.state_save()[24:]
    # the checker on a quarter of the streams (kept detections repeat with the head period)
    for b in range(0, B, 4):
        dets = [orc.postprocess(host[f, b]) for f in range(F)]
        trk = orc.Tracker()
        for f in range(STEPS):
            trk.update(dets[f % F]["poses"], dets[f % F]["scores"], f)
        assert trk.get_tracks().tobytes() == oa[b, :ca[b]].tobytes(), b
        s1, s2 = trk.get_state(), a.get_state(b)
        for k in ("ids", "hits", "ages", "states", "active", "row_assign", "col_assign", "poses", "vel"):
            assert s1[k].tobytes() == s2[k].tobytes(), (b, k)


@pytest.mark.gpu
def test_config4_full_shape_128_streams_per_gpu(pb, orc, cuda):
    """BASELINE config 4 per-GPU shape: 128 streams (1024 over 8 GPUs), max-age 30, occlusion gaps.  128 one-CTA-per-SM
    tracker grids do not overlap (two would not fit 148 SMs), which is a different launch plan from the 64-stream
    case: pipelined steps without synchronisation against a serial handle (all streams) and the checker (every 8th)."""
    torch = cuda
    B, F, STEPS = 128, 10, 60
    scfg = pb.synth_config(canvas=640, persons=20, period=F, occlusion=1)
    host = pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)
    heads = torch.from_numpy(host).cuda()
    serial = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, max_age=30)
    piped = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, max_age=30, pipeline_depth=5)
    for f in range(STEPS):
        serial.step(heads[f % F], f); piped.step(heads[f % F], f)
    piped.join(); torch.cuda.synchronize()
    o1, c1 = serial.get_tracks_all(); o2, c2 = piped.get_tracks_all()
    assert np.array_equal(c1, c2) and c1.sum() > 500 and valid_records(o1, c1) == valid_records(o2, c2)
    assert serial.state_save()[24:] == piped.state_save()[24:]
    for b in range(0, B, 8):
        dets = [orc.postprocess(host[f, b]) for f in range(F)]
        trk = orc.Tracker(max_age=30)
        for f in range(STEPS):
            trk.update(dets[f % F]["poses"], dets[f % F]["scores"], f)
        assert trk.get_tracks().tobytes() == o2[b, :c2[b]].tobytes(), b
        s1, s2 = trk.get_state(), piped.get_state(b)
        for k in ("ids", "hits", "ages", "states", "active", "row_assign", "poses", "vel"):
            assert s1[k].tobytes() == s2[k].tobytes(), (b, k)


@pytest.mark.gpu
def test_config3_crowd_16_streams(pb, orc, cuda):
    """BASELINE config 3 at B = 16: 1280x1280 heads [16,56,33600], 100 persons in clumps, max_tracks 256 / max_detections 128."""
    cfg = pb.synth_config(canvas=1280, persons=100, period=6, clumps=10, kp_drop_prob=0.15)
    heads = pb.synth_heads(cfg, 0, 16, 0, 6, frame_major=True)
    got, total = gpu_hashes(pb, cuda, heads, max_tracks=256, max_detections=128)
    ref = orc.run_streams(heads, True, threads=8, max_tracks=256, max_detections=128)
    assert np.array_equal(got, ref["hashes"]) and total == ref["tracks_total"] > 2000


@pytest.mark.gpu
def test_step_seq_equals_step_loop(pb, cuda):
    torch = cuda
    B, F = 5, 6
    scfg = pb.synth_config(canvas=640, persons=9, period=16)
    heads = torch.from_numpy(pb.synth_heads(scfg, 3, B, 0, F, frame_major=True)).cuda()
    a = pb.Pipeline(num_streams=B, pipeline_depth=4); b = pb.Pipeline(num_streams=B, pipeline_depth=4)
    for f in range(15):
        a.step(heads[(2 + f) % F], 100 + f)
    b.step_seq(heads, 2, 15, 100)
    a.join(); b.join()
    o1, c1 = a.get_tracks_all(); o2, c2 = b.get_tracks_all()
    assert np.array_equal(c1, c2) and c1.sum() > 0 and valid_records(o1, c1) == valid_records(o2, c2)
    assert a.state_save()[24:] == b.state_save()[24:]


@pytest.mark.gpu
@pytest.mark.parametrize("B,occlusion,max_age,chunk,tier,persons", [(6, 1, 4, 0, 0, 12), (6, 1, 4, 16, 0, 12), (64, 0, 10, 16, 0, 12), (70, 1, 30, 7, 0, 12),
                                                                    (64, 0, 10, 16, 3, 12), (70, 1, 30, 7, 4, 12), (12, 0, 10, 5, 2, 12),
                                                                    (9, 1, 10, 16, 3, 40), (9, 0, 10, 16, 4, 40),
                                                                    (64, 0, 10, 0, -1, 12), (72, 1, 30, 0, -1, 12), (9, 0, 10, 0, -1, 40)])
def test_resident_tracker_sequence_path(pb, orc, cuda, monkeypatch, B, occlusion, max_age, chunk, tier, persons):
    """pb_step_seq on a pipelined handle takes the resident-tracker path (one tracker CTA per video stream stays on its SM for
    a chunk of up to 16 steps here, the stream's state in shared memory, and is fed by the decode and NMS kernels of those
    steps; up to 74 streams): sequences that are not multiples of the chunk, back-to-back calls
    without a join, single steps, a reset and a stage-level call in between — states and records must equal the serial
    handle's bit for bit, and the checker's on some streams."""
    torch = cuda
    if chunk:                                      # (0: the handle's defaults — resident path up to 49 streams, chunks of 32)
        monkeypatch.setenv("PB_SEQ", "1"); monkeypatch.setenv("PB_SEQ_CHUNK", str(chunk))
    # tier > 0: the steps' NMS kernel is a tiered one (2 x 512, 3 x 384, 3 x 256 or 4 x 256 threads per SM); with 40 persons
    # (about 280 candidates) the streams exceed the shared-memory tier of the small CTAs and take their spill path
    # tier -1: the handle's own choice — pairs of 512-thread NMS CTAs with 81 KB each (the working set of 256 candidates; 40 persons
    # exceed it: spill path), four lanes, resident path up to 74 streams
    if tier >= 0:
        monkeypatch.setenv("PB_SEQ_NMS_TIER", str(tier))
    F = 24
    scfg = pb.synth_config(canvas=640, persons=persons, period=48, occlusion=occlusion)
    host = pb.synth_heads(scfg, 11, B, 0, F, frame_major=True)
    heads = torch.from_numpy(host).cuda()
    serial = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, max_age=max_age, fuse_stages=0)
    res = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, max_age=max_age, pipeline_depth=5)
    f = 0
    def both_seq(n):
        nonlocal f
        for i in range(n):
            serial.step(heads[(f + i) % F], f + i)
        res.step_seq(heads, f % F, n, f)
        f += n
    def compare(tag, streams=None):
        o1, c1 = serial.get_tracks_all(); o2, c2 = res.get_tracks_all()
        assert np.array_equal(c1, c2) and c1.sum() > 0, tag
        assert valid_records(o1, c1) == valid_records(o2, c2), tag
        assert serial.state_save()[24:] == res.state_save()[24:], tag
        for b in (streams or (0, B - 1)):
            k1, k2 = serial.get_kept(b), res.get_kept(b)
            assert np.array_equal(k1["keep_anchors"], k2["keep_anchors"]), (tag, b)
    both_seq(45)                                   # 16 + 16 + 13
    both_seq(3); both_seq(2); both_seq(37)         # back to back, no join in between
    compare("after 87 steps in four calls")
    for i in range(5):                             # single steps (the per-step pipelined path) behind the resident path
        serial.step(heads[(f + i) % F], f + i); res.step(heads[(f + i) % F], f + i)
    f += 5
    both_seq(20)
    compare("after single steps + 20")
    serial.reset(); res.reset(); f = 0
    both_seq(18)
    serial.postprocess(heads[3]); res.postprocess(heads[3])
    serial.tracker_update(500); res.tracker_update(500)
    both_seq(33)
    compare("after reset, 18 steps, a stage-level call and 33 steps")
    # the checker on two streams over the last phase
    for b in (0, B - 1):
        trk = orc.Tracker(max_age=max_age)
        for g in range(18):
            d = orc.postprocess(host[g % F, b]); trk.update(d["poses"], d["scores"], g)
        d = orc.postprocess(host[3, b]); trk.update(d["poses"], d["scores"], 500)
        for g in range(18, 51):
            d = orc.postprocess(host[g % F, b]); trk.update(d["poses"], d["scores"], g)
        o2, c2 = res.get_tracks_all()
        assert trk.get_tracks().tobytes() == o2[b, :c2[b]].tobytes(), b
