"""GPU (-m gpu): the round-2 entry points of the C ABI -- float64 camera rows, lens undistortion inside the input stage,
whole-video evaluation from pixels (device and host buffers, flip augmentation), and stream ordering between the
synchronous and the lane-based calls.  Everything here is checked for BIT equality against the already-pinned
paths (ray encode vs the reference's CameraInfoPacket fixtures, forward_rays vs the reference's goldens)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, relerr
from oracle import ray3d_oracle as O
from ray3d_b200 import Lifter, NetSpec, RayCamera, synth

pytestmark = pytest.mark.gpu

SPEC = NetSpec(filter_widths=(3, 3, 3))          # cfg_ray3d_h36m_stage1 architecture with RF = 27 (BASELINE configs[0])
_lf = {}


def lifter(precision="bf16x3", spec=SPEC):
    key = (precision, spec)
    if key not in _lf:
        _lf.clear()
        sp, st = synth.make_state_dicts(spec)
        _lf[key] = Lifter(spec, sp, st, precision=precision)
    return _lf[key]


def cameras():
    c = load_golden("camera")
    plain = [RayCamera(c[f"K{i}"], c[f"R{i}"], c[f"t{i}"], res_w=1000, res_h=1002) for i in range(6)]
    g = load_golden("camera_undistort")
    lens = RayCamera(g["K"], g["R"], g["t"], res_w=1000, res_h=1002, undistort=True, dist_coeff=g["dist"])
    return plain, lens, g


def windows_uv(B, T, J=17, seed=3, res=1000.0):
    rng = np.random.default_rng(seed)
    return (rng.uniform(0.2, 0.8, size=(B, T, J, 2)) * res).astype(np.float32)


@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
def test_forward_uv_float64_camera_rows(precision):
    """R3D_CAM_F64 rows carry the reference's own float64 calibration (camera.py:438-439) and libm's cos/sin of the
    pitch: the fused encode then equals CameraInfoPacket.get_cam_ray_given_uv(...).astype(float32) BIT FOR BIT, so
    the network outputs equal forward_rays on the reference-encoded rays exactly -- also for calibration values that
    are not float32-representable (the golden cameras' K entries are float64 draws)."""
    lf = lifter(precision)
    plain, _, _ = cameras()
    T, B = SPEC.receptive_field, len(plain)
    uv = windows_uv(B, T)
    rays = np.stack([cam.get_cam_ray_given_uv(uv[i].astype(np.float64)) for i, cam in enumerate(plain)]).astype(np.float32)
    param = np.stack([cam.param for cam in plain])
    rows = np.stack([cam.table_row64() for cam in plain])
    assert not all(float(np.float32(r[0])) == r[0] for r in rows)            # genuinely float64 focal lengths
    a = lf.forward_uv(torch.from_numpy(uv).cuda(), torch.from_numpy(rows).cuda())
    b = lf.forward_rays(torch.from_numpy(rays).cuda(), torch.from_numpy(param).cuda())
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    # the same through host buffers
    out = lf.forward_uv_host(torch.from_numpy(uv), torch.from_numpy(rows))
    assert torch.equal(out, a[2].cpu())


def test_lens_undistortion_inside_the_fused_call():
    """camera.py:435-436: encode_uv_with_intrinsic undistorts first when the camera says so.  The standalone float64
    path is pinned bit-exactly to the reference fixture (test_lens_undistortion_bit_exact); the fused input stage must
    produce the very same float32 rays, i.e. identical network outputs."""
    lf = lifter("bf16x3")
    plain, lens, g = cameras()
    T = SPEC.receptive_field
    uv = windows_uv(4, T, seed=5)
    cams = [lens, plain[0], lens, plain[1]]                                      # mixed batch: flag is per row
    rays = np.stack([cam.get_cam_ray_given_uv(uv[i].astype(np.float64)) for i, cam in enumerate(cams)]).astype(np.float32)
    # the standalone encode of the distorted camera really differs from the pinhole one
    pin = RayCamera(lens.K, lens.Rw2c, lens.Tw2c).get_cam_ray_given_uv(uv[0].astype(np.float64)).astype(np.float32)
    assert not np.array_equal(pin, rays[0])
    param = np.stack([cam.param for cam in cams])
    rows = np.stack([cam.table_row64() for cam in cams])
    a = lf.forward_uv(torch.from_numpy(uv).cuda(), torch.from_numpy(rows).cuda())
    b = lf.forward_rays(torch.from_numpy(rays).cuda(), torch.from_numpy(param).cuda())
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    with pytest.raises(ValueError):
        lens.table_row()                                                         # the float32 row has no lens model


@pytest.mark.parametrize("use_lens", [False, True])
def test_video_from_pixels_device_and_host(use_lens):
    """Lifter.forward_video_uv == eval_data_prepare + np.tile + encode + forward (trainer.py:47-58, 297-337) without
    materialising windows and with ONE encode per frame; host-buffer form and flip augmentation included."""
    lf = lifter("bf16x3")
    plain, lens, _ = cameras()
    cam = lens if use_lens else plain[2]
    T, F = SPEC.receptive_field, 150
    rng = np.random.default_rng(11)
    uv_seq = (rng.uniform(0.25, 0.75, size=(F + T - 1, 17, 2)) * 1000).astype(np.float32)
    rays_seq = cam.get_cam_ray_given_uv(uv_seq.astype(np.float64)).astype(np.float32)
    prm = torch.from_numpy(cam.param).cuda()
    want = lf.forward_video(torch.from_numpy(rays_seq).cuda(), prm)
    got = lf.forward_video_uv(torch.from_numpy(uv_seq).cuda(), cam)
    assert all(torch.equal(x, y) for x, y in zip(got, want))
    # materialised windows through the per-window encode give the same bits
    win = np.stack([uv_seq[f:f + T] for f in range(F)])
    rows = np.tile(cam.table_row64(), (F, 1))
    mat = lf.forward_uv(torch.from_numpy(win).cuda(), torch.from_numpy(rows).cuda())
    assert all(torch.equal(x, y) for x, y in zip(got, mat))
    # host buffers: 136 bytes per frame in, F results out
    out = lf.forward_video_uv_host(torch.from_numpy(uv_seq), cam)
    assert torch.equal(out, want[2].cpu())
    pin_uv, pin_row = torch.from_numpy(uv_seq).pin_memory(), torch.from_numpy(cam.table_row64()).pin_memory()
    pin_out = torch.empty((F, 1, 17, 3), dtype=torch.float32).pin_memory()
    tickets = [lf.submit_video_uv_host(pin_uv, pin_row, pin_out) for _ in range(3)]
    for t in tickets:
        lf.wait(t)
    assert torch.equal(pin_out, want[2].cpu())
    # flip augmentation (trainer.py:299-353)
    left, right = [4, 5, 6, 11, 12, 13], [1, 2, 3, 14, 15, 16]
    lf.set_flip(left, right)
    tta_want = lf.forward_video_tta(torch.from_numpy(rays_seq).cuda(), prm)
    tta_got = lf.forward_video_uv(torch.from_numpy(uv_seq).cuda(), cam, tta=True)
    assert all(torch.equal(x, y) for x, y in zip(tta_got, tta_want))
    assert torch.equal(lf.forward_video_uv_host(torch.from_numpy(uv_seq), cam, tta=True), tta_want[2].cpu())
    # a video longer than one staged chunk (1024 windows) splits across the lanes and still matches
    F2 = 2300
    uv2 = (rng.uniform(0.25, 0.75, size=(F2 + T - 1, 17, 2)) * 1000).astype(np.float32)
    dev2 = lf.forward_video_uv(torch.from_numpy(uv2).cuda(), cam)
    assert torch.equal(lf.forward_video_uv_host(torch.from_numpy(uv2), cam), dev2[2].cpu())
    ref = O.lift(*[O.to_torch_state(s, torch.float64) for s in synth.make_state_dicts(SPEC)], SPEC,
                 torch.from_numpy(np.stack([cam.get_cam_ray_given_uv(uv2[f:f + T].astype(np.float64)).astype(np.float32) for f in (0, F2 - 1)])).double(),
                 torch.from_numpy(np.tile(cam.param, (2, 1))).double())[2].numpy()
    assert relerr(dev2[2][[0, F2 - 1]].cpu().numpy(), ref) < 1e-4


def test_sync_and_lane_calls_are_ordered_on_the_device():
    """r3d_forward_* on a caller stream shares lane 0's workspace with r3d_submit_* / *_host.  Interleave
    them on different streams without any host synchronisation; every result must equal the serial one."""
    spec = NetSpec(filter_widths=(3, 3, 3, 3))
    lf = lifter("bf16x3", spec)
    B = 300
    sets = []
    for i in range(4):
        uv, cam = synth.make_inputs(spec, B, seed=70 + i)
        sets.append((torch.from_numpy(uv).cuda(), torch.from_numpy(cam).cuda(), torch.from_numpy(uv).pin_memory(), torch.from_numpy(cam).pin_memory()))
    serial = []
    for uv, cam, _, _ in sets:
        serial.append(lf.forward_uv(uv, cam)[2].clone())
        torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs_h = [torch.empty((B, 1, 17, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    for rep in range(6):
        pend = lf.submit_uv(sets[0][0], sets[0][1])                     # lane 0 (or 1), not joined yet
        with torch.cuda.stream(s1):
            a = lf.forward_uv(sets[1][0], sets[1][1])[2]                # caller stream, lane 0's workspace
        tk = lf.submit_uv_host(sets[2][2], sets[2][3], outs_h[0])       # copy stream + lane compute stream
        with torch.cuda.stream(s2):
            b = lf.forward_uv(sets[3][0], sets[3][1])[2]
        pend2 = lf.submit_uv(sets[1][0], sets[1][1])
        tk2 = lf.submit_uv_host(sets[3][2], sets[3][3], outs_h[1])
        c = lf.join(pend)[2]
        d = lf.join(pend2)[2]
        lf.wait(tk)
        lf.wait(tk2)
        torch.cuda.synchronize()
        assert torch.equal(a, serial[1]) and torch.equal(b, serial[3])
        assert torch.equal(c, serial[0]) and torch.equal(d, serial[1])
        assert torch.equal(outs_h[0], serial[2].cpu()) and torch.equal(outs_h[1], serial[3].cpu())


def test_inputs_on_the_wrong_device_or_of_the_wrong_kind_are_refused():
    lf = lifter("bf16x3")
    T = SPEC.receptive_field
    x = torch.zeros(2, T, 17, 3)
    with pytest.raises(RuntimeError):
        lf.forward_rays(x, torch.zeros(2, 2))                        # CPU tensors
    with pytest.raises(RuntimeError):
        lf.forward_rays(x.cuda(), None)                              # embedding on, param missing
    with pytest.raises(RuntimeError):
        lf.submit_rays(x.cuda(), None)
    with pytest.raises(AssertionError):
        lf.forward_rays(x.cuda(), torch.zeros(3, 2).cuda())          # param rows != windows
    if torch.cuda.device_count() > 1:
        with pytest.raises(RuntimeError):
            lf.forward_rays(x.to("cuda:1"), torch.zeros(2, 2, device="cuda:1"))
        assert torch.cuda.current_device() == 0                      # the plan never leaves the caller on another device


@pytest.mark.parametrize("stage", [1, 3])
def test_chained_tail_launch_is_bit_identical(stage):
    """Chained launches: the GlobalInfo chain (default, on a few CTA pairs of the side stream; rie.py:362) and -- option
    tail_fusion -- the one-row layers of the main chain (top tree level, shrink, FuseBlocks, Integration; rie.py:94-105,
    388-414) as ONE persistent kernel each, whose work units wait on per-(op, problem, row group) completion counters.
    Same tile code, same arithmetic: outputs must equal the one-launch-per-layer form bit for bit (both unit widths,
    ragged last row group, fewer CTA pairs than dependency chains)."""
    spec = NetSpec(filter_widths=(3, 3, 3), stage=stage)
    sp, st = synth.make_state_dicts(spec)
    base = Lifter(spec, sp, st, precision="bf16x3", options={"side_chain": 0})        # every op its own launch
    assert not any(L["tail"] for L in base.plan.describe()["launches"])
    # (side_chain=2 forces the chained GlobalInfo launch: by default it is only planned for long receptive fields)
    for opts in ({"side_chain": 2}, {"side_chain": 2, "side_clusters": 3}, {"tail_fusion": 1, "tail_width": 128, "side_chain": 2},
                 {"tail_fusion": 1, "tail_width": 256, "side_chain": 0}):
        lf = Lifter(spec, sp, st, precision="bf16x3", options=opts)
        width = opts
        assert any(L["tail"] for L in lf.plan.describe()["launches"])
        for B in (700, 256, 300, 40):                     # < 256 windows: the plan falls back to one launch per op
            uv, cam = synth.make_inputs(spec, B, seed=90 + B)
            uvc, camc = torch.from_numpy(uv).cuda(), torch.from_numpy(cam).cuda()
            a, b = lf.forward_uv(uvc, camc), base.forward_uv(uvc, camc)
            assert all(torch.equal(x, y) for x, y in zip(a, b)), (stage, width, B)
        del lf


def test_tile_policy_is_result_neutral():
    """Option tile_policy picks the GEMM tile width of the narrow launches (latency: wave-count model; throughput: widest
    tile; auto switches at 256 windows of capacity).  The K order of every dot product is the same: bit-identical."""
    spec = NetSpec(filter_widths=(3, 3, 3), stage=1)
    sp, st = synth.make_state_dicts(spec)
    lat = Lifter(spec, sp, st, precision="bf16x3", options={"tile_policy": 1})
    thr = Lifter(spec, sp, st, precision="bf16x3", options={"tile_policy": 2})
    auto = Lifter(spec, sp, st, precision="bf16x3")
    for B in (40, 600):
        uv, cam = synth.make_inputs(spec, B, seed=70 + B)
        uvc, camc = torch.from_numpy(uv).cuda(), torch.from_numpy(cam).cuda()
        a, b, c = lat.forward_uv(uvc, camc), thr.forward_uv(uvc, camc), auto.forward_uv(uvc, camc)
        assert all(torch.equal(x, y) and torch.equal(x, z) for x, y, z in zip(a, b, c)), B
    tiles = lambda lf: [o["n_tile"] for o in lf.plan.describe()["ops"]]
    assert tiles(auto) == tiles(thr) and all(x <= y for x, y in zip(tiles(lat), tiles(thr)))   # capacity grew to 600 windows


@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
def test_whole_evaluate_core_step_vs_the_reference(precision):
    """One video through Lifter.forward_video_tta / forward_video_uv(tta=True) + metrics.evaluate, against what the
    reference's own Trainer.evaluate_core(flip_test=True) produced for it (tests/golden/evaluate_core_tta.npz: the
    tensor it hands to normalized2world, computed by the reference modules in float64, and its five metrics)."""
    import ray3d_b200
    g = load_golden("evaluate_core_tta")
    lf = lifter(precision)
    lf.set_flip(g["kps_left"].tolist(), g["kps_right"].tolist())
    tol = {"bf16x3": 1e-4, "fp32": 2e-6}[precision]
    both = lf.forward_video_tta(torch.from_numpy(g["rays"]).cuda(), torch.from_numpy(g["param"]).cuda())[2]
    assert relerr(both.cpu().numpy(), g["pred64"]) < tol
    # from pixels, with the camera the fixture was made with (float64 row): same float32 rays -> same bits
    cam = RayCamera(g["K"], g["R"], g["t"], res_w=1000, res_h=1002)
    both_uv = lf.forward_video_uv(torch.from_numpy(g["uv"].astype(np.float32)).cuda(), cam, tta=True)[2]
    rays32 = cam.get_cam_ray_given_uv(g["uv"].astype(np.float32).astype(np.float64)).astype(np.float32)
    assert torch.equal(both_uv, lf.forward_video_tta(torch.from_numpy(rays32).cuda(), torch.from_numpy(cam.param).cuda())[2])
    # evaluate_core's metrics (millimetres) from the device-side evaluation tail
    m = ray3d_b200.metrics.evaluate(both, torch.from_numpy(g["target"][:, None]).cuda(), g["Rn2w"], g["Tn2w"])
    e1, e2, e3, ev, er = g["metrics64"]
    for want, key in ((e1, "mpjpe"), (e2, "p_mpjpe"), (e3, "n_mpjpe"), (ev, "mpjve"), (er, "mrpe")):
        assert abs(m[key] * 1000 - want) <= 5 * tol * abs(want), (key, m[key] * 1000, want)
