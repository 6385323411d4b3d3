"""Boundary robustness of the C ABI on the GPU: NaN/inf confidences, detections spilling to the
global scratch under pipelining, a foreign SM-filling kernel beside the pipelined steps, handles on
two devices in one process, argument validation of the Python binding."""
import numpy as np
import pytest

from helpers import valid_records

pytestmark = pytest.mark.gpu


def test_nan_and_inf_confidences_rank_like_the_reference_sort(pb, orc, cuda):
    """NaN passes the confidence filter (`!(conf < thr)`, gpu_postprocess.cu:51) and the reference's insertion
    sort (:192-202) never moves an element across a NaN: the rank order is a permutation with the NaNs
    as fixed points.  +inf / -inf are ordinary scores."""
    rng = np.random.default_rng(11)
    cfg = pb.synth_config(canvas=640, persons=12, period=8)
    heads = pb.synth_heads(cfg, 3, 4, 0, 1, frame_major=True)[0].copy()
    cand = [np.flatnonzero(~(heads[b, 4] < 0.30)) for b in range(4)]
    assert all(len(c) > 40 for c in cand)
    heads[0, 4, cand[0][[3, 17, 18, 40]]] = np.nan            # NaNs inside the candidate list
    heads[1, 4, cand[1][0]] = np.nan; heads[1, 4, cand[1][-1]] = np.nan      # first and last slot
    heads[2, 4, cand[2][5]] = np.inf; heads[2, 4, cand[2][9]] = np.inf; heads[2, 4, cand[2][7]] = np.nan
    heads[3, 4, rng.choice(8400, 30, replace=False)] = np.nan  # NaNs at background anchors become candidates
    for mode in (0, 3, 1):
        pipe = pb.Pipeline(num_streams=4, keypoint_fetch=mode)
        pipe.postprocess(cuda.from_numpy(heads).cuda())
        cuda.cuda.synchronize()
        for b in range(4):
            got, ref = pipe.get_kept(b), orc.postprocess(heads[b])
            assert got["num_cand"] == ref["num_cand"]
            assert np.array_equal(got["keep_anchors"], ref["keep_anchors"]), (mode, b)
            assert np.array_equal(got["keep_slots"], ref["keep_slots"]), (mode, b)
            for k in ("poses", "bboxes", "scores"):
                assert got[k].tobytes() == ref[k].tobytes(), (mode, b, k)


def test_pipelined_steps_with_detections_in_global_scratch(pb, orc, cuda):
    """max_detections so large that the tracker keeps this frame's detections in a global scratch instead of
    shared memory: consecutive tracker grids overlap on different lanes, and the successor of a video stream
    must not overwrite the scratch its predecessor still reads."""
    torch = cuda
    B, F, STEPS = 40, 16, 64
    scfg = pb.synth_config(canvas=640, persons=16, period=F, occlusion=1)
    host = pb.synth_heads(scfg, 21, B, 0, F, frame_major=True)
    heads = torch.from_numpy(host).cuda()
    kw = dict(num_streams=B, num_anchors=scfg.num_anchors, max_tracks=128, max_detections=1024, max_age=4)
    serial = pb.Pipeline(**kw)
    piped = pb.Pipeline(pipeline_depth=5, **kw)
    for f in range(STEPS):
        serial.step(heads[f % F], f); piped.step(heads[f % F], f)
    piped.join(); torch.cuda.synchronize()
    o1, c1 = serial.get_tracks_all(); o2, c2 = piped.get_tracks_all()
    assert np.array_equal(c1, c2) and c1.sum() > 0
    assert valid_records(o1, c1) == valid_records(o2, c2)
    assert serial.state_save()[24:] == piped.state_save()[24:]
    trk = orc.Tracker(max_tracks=128, max_detections=1024, max_age=4)
    dets = [orc.postprocess(host[f, 5]) for f in range(F)]
    for f in range(STEPS):
        trk.update(dets[f % F]["poses"], dets[f % F]["scores"], f)
    assert trk.get_tracks().tobytes() == o2[5, :c2[5]].tobytes()


@pytest.mark.parametrize("fuse", [0, 1])
def test_foreign_kernel_filling_the_sms_beside_pipelined_steps(pb, cuda, fuse):
    """The drop-in site has a TensorRT engine on the same GPU: long foreign kernels that fill every SM run on
    another stream while pipelined steps (tracker CTAs spinning on their predecessors' flags) are in flight.
    Results must equal the serial path's and no time-out may be reported."""
    torch = cuda
    B, F, STEPS = 64, 16, 96
    scfg = pb.synth_config(canvas=640, persons=16, period=F, occlusion=1)
    heads = torch.from_numpy(pb.synth_heads(scfg, 50, B, 0, F, frame_major=True)).cuda()
    serial = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, fuse_stages=0)
    for f in range(STEPS):
        serial.step(heads[f % F], f)
    o1, c1 = serial.get_tracks_all()
    piped = pb.Pipeline(num_streams=B, num_anchors=scfg.num_anchors, pipeline_depth=5, fuse_stages=fuse)
    side = torch.cuda.Stream()
    a = torch.randn(8192, 8192, device="cuda")
    big = torch.empty(1 << 28, device="cuda")
    for f in range(STEPS):
        if f % 3 == 0:
            with torch.cuda.stream(side):
                a = torch.tanh(a @ a * 1e-4)        # GEMM grids + elementwise grids far larger than the device
                big.normal_()
        piped.step(heads[f % F], f)
    piped.join(); torch.cuda.synchronize()
    piped.wait()                                   # raises if a tracker CTA gave up waiting for its predecessor
    o2, c2 = piped.get_tracks_all()
    assert np.array_equal(c1, c2) and valid_records(o1, c1) == valid_records(o2, c2)
    assert serial.state_save()[24:] == piped.state_save()[24:]


def test_two_handles_on_two_devices_in_one_process(pb, cuda):
    """The opt-in to more than 48 KB of dynamic shared memory is a per-device kernel attribute and every entry
    point has to run on its handle's device whatever the caller's current device is."""
    torch = cuda
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    B, F = 6, 12
    scfg = pb.synth_config(canvas=640, persons=10, period=16)
    host = pb.synth_heads(scfg, 0, B, 0, F, frame_major=True)
    h0, h1 = torch.from_numpy(host).to("cuda:0"), torch.from_numpy(host).to("cuda:1")
    torch.cuda.set_device(0)
    p0 = pb.Pipeline(num_streams=B, device=0)
    p1 = pb.Pipeline(num_streams=B, device=1, pipeline_depth=3)
    assert torch.cuda.current_device() == 0                      # pb_create gave the caller's device back
    s0 = torch.cuda.Stream(device=0); s1 = torch.cuda.Stream(device=1)
    for f in range(F):
        torch.cuda.set_device(f & 1)                             # current device alternates: both handles must not care
        p0.step(h0[f], f, stream=s0)
        p1.step(h1[f], f, stream=s1)
    p1.join(stream=s1)
    o0, c0 = p0.get_tracks_all(); o1, c1 = p1.get_tracks_all()
    assert np.array_equal(c0, c1) and c0.sum() > 0 and valid_records(o0, c0) == valid_records(o1, c1)
    assert p0.state_save()[24:] == p1.state_save()[24:]
    # the stand-alone entry points with large shared-memory requests on the second device
    torch.cuda.set_device(0)
    cost = torch.rand(2, 300, 300, device="cuda:1")
    with torch.cuda.device(1):
        r1 = pb.greedy_match(cost, 0.5)
    with torch.cuda.device(0):
        r0 = pb.greedy_match(cost.to("cuda:0"), 0.5)
    assert torch.equal(r0.cpu(), r1.cpu())


def test_binding_rejects_wrong_buffers(pb, cuda):
    torch = cuda
    pipe = pb.Pipeline(num_streams=2)
    good = torch.zeros(2, 56, 8400, device="cuda")
    pipe.step(good, 0)
    with pytest.raises(ValueError):
        pipe.step(good.half(), 1)                                # dtype
    with pytest.raises(ValueError):
        pipe.step(good.transpose(1, 2), 1)                       # not contiguous
    with pytest.raises(ValueError):
        pipe.step(good[:1], 1)                                   # too small
    with pytest.raises(ValueError):
        pipe.step(good.cpu(), 1)                                 # host tensor on the device entry point
    with pytest.raises(ValueError):
        pipe.step_host(np.zeros((2, 56, 8400), np.float64), 1)
    with pytest.raises(ValueError):
        pipe.step_host(np.zeros((2, 8400, 56), np.float32).transpose(0, 2, 1), 1)
    with pytest.raises(pb.PbError):
        pipe.state_load(b"short")
