"""CPU: drop-in nn.Module surface (names, order, factory, loud failures) and the multi-rank host logic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ray3d_b200
from ray3d_b200 import dist as rdist, synth
from ray3d_b200.spec import NetSpec

MODEL_CONFIG = {   # the keys lib/model/__init__.py:11-46 reads, values of cfg_ray3d_h36m_stage1.py
    'MODEL': 'RIE', 'ARCHITECTURE': '3,3', 'DROPOUT': 0.2, 'CAUSAL': False, 'CHANNELS': 256, 'DENSE': False, 'NUM_KPTS': 17,
    'INPUT_DIM': 3, 'CAMERA_EMBDDING': True, 'EXTRINSIC_DIM': 2, 'EMBEDD_DIM': 64, 'LATENT_FEATURES_DIM': 256,
    'DISABLE_OPTIMIZATIONS': False, 'STAGE': 1, 'TRAJECTORY_MODEL': True,
}


def _mods(meta, name):
    kw = dict(meta[name]["spec"])
    ctor = dict(filter_widths=kw["filter_widths"], latten_features=256, channels=256, stage=kw.get("stage", 1),
                extrinsic_dim=kw.get("extrinsic_dim", 2), embedd_dim=kw.get("embed_dim", 64))
    J, C = kw.get("num_joints", 17), kw.get("in_features", 3)
    return ray3d_b200.RIEModel(J, C, J, **ctor), ray3d_b200.RIETrajectoryModel(J, C, J, **ctor)


@pytest.mark.parametrize("name", ["h36m_s1_t27", "h36m_s3_t9", "humaneva_s1_t9", "h36mcross_s2_t9", "rie_s1_t9_noembed"])
def test_state_dict_is_identical_to_reference(golden_meta, name):
    pos, trj = _mods(golden_meta, name)
    m = golden_meta[name]
    assert [(k, list(v.shape)) for k, v in pos.state_dict().items()] == [tuple(e) for e in map(tuple, m["keys_pos"])]
    assert [(k, list(v.shape)) for k, v in trj.state_dict().items()] == [tuple(e) for e in map(tuple, m["keys_trj"])]
    assert [k for k, _ in pos.named_parameters()] == m["named_params_pos"]     # main.py:164-168 freezes by index
    assert pos.receptive_field() == m["receptive_field"] == trj.receptive_field()
    # strict load of reference-named weights works
    spec = pos._spec
    sp, st = synth.make_state_dicts(spec)
    pos.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sp.items()}, strict=True)
    trj.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in st.items()}, strict=True)


def test_factory_surface_and_loud_failures():
    m = ray3d_b200.Model(MODEL_CONFIG, None, is_train=False)
    pos, trj = m.get_pos_model(), m.get_trj_model()
    pos_m, trj_m = getattr(pos, "module", pos), getattr(trj, "module", trj)
    assert isinstance(pos_m, ray3d_b200.RIEModel) and isinstance(trj_m, ray3d_b200.RIETrajectoryModel)
    assert pos_m.receptive_field() == 9
    pos_m.set_bn_momentum(0.05); pos_m.set_training_status(False); pos_m.set_augment(False)
    assert pos_m.LocalLayer_Torso.expand_bn.momentum == 0.05
    x = torch.zeros(2, 9, 17, 3)
    prm = torch.zeros(2, 2)
    if not torch.cuda.is_available():
        pos_m.eval()
        with pytest.raises(RuntimeError, match="no CPU path"):
            pos_m(x, prm)
    pos_m.train()
    with pytest.raises(RuntimeError, match="eval-mode"):
        pos_m(x.cuda() if torch.cuda.is_available() else x, prm)
    with pytest.raises(AssertionError):
        pos_m.eval()(torch.zeros(2, 9, 16, 3), prm)          # rie.py:286
    with pytest.raises(ValueError):
        ray3d_b200.Model(dict(MODEL_CONFIG, MODEL='VideoPose3D'), None)
    with pytest.raises(NotImplementedError):
        ray3d_b200.Model(dict(MODEL_CONFIG, CAUSAL=True), None)


def test_shard_bounds_partition():
    for batch in (0, 1, 7, 8, 1024, 8191):
        for world in (1, 2, 3, 8):
            spans = [rdist.shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_main(rank, world, port, batch, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        J = 17
        g = torch.Generator().manual_seed(5)
        uv = torch.rand(batch, 9, J, 2, generator=g)
        cam = torch.rand(batch, 6, generator=g)

        def fake_lift(u, c):     # stands in for Lifter.forward_uv on the CPU: any per-sequence function
            both = (u.mean(dim=1, keepdim=True).repeat(1, 1, 1, 2)[..., :3] + c[:, None, None, :3])
            return both, both[:, :, :1] * 2

        lo, hi = rdist.shard_bounds(batch, rank, world)
        both, trj = rdist.lift_sharded(fake_lift, uv[lo:hi], cam[lo:hi], batch)
        ref_both, ref_trj = fake_lift(uv, cam)
        ok = bool(torch.equal(both, ref_both) and torch.equal(trj, ref_trj))
        if batch % world == 0:     # streaming gather: several collectives in flight, results in submission order
            og = rdist.OverlappedGather(hi - lo, J, "cpu", depth=2)
            tickets = []
            for step in range(5):
                b, t = fake_lift(uv[lo:hi] + step, cam[lo:hi])
                tickets.append(og.submit(b, t))
                if step >= 1:
                    got_b, got_t = rdist.unpack_outputs(og.result(tickets[step - 1]))
                    want_b, want_t = fake_lift(uv + (step - 1), cam)
                    ok = ok and bool(torch.equal(got_b, want_b) and torch.equal(got_t, want_t))
            og.drain()
            try:
                og.result(0)
                ok = False
            except ValueError:
                pass
        out_q.put((rank, ok, tuple(both.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 7])
def test_sharded_lift_all_gather_gloo_world2(batch):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, batch, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), res
    assert all(r[2] == (batch, 1, 17, 3) for r in res)


def test_camera_undistorted_principal_point_host():
    """RayCamera(undistort=True) computes pp_cam on the host with the device kernel's arithmetic; it must match the
    reference's cv2.undistortPoints value bit for bit (camera.py:253-256)."""
    from conftest import load_golden
    g = load_golden("camera_undistort")
    cam = ray3d_b200.RayCamera(g["K"], g["R"], g["t"], undistort=True, dist_coeff=g["dist"])
    assert np.array_equal(cam.pp_cam.reshape(-1), g["pp_cam"].reshape(-1))
    with pytest.raises(ValueError):
        ray3d_b200.RayCamera(g["K"], g["R"], g["t"], undistort=True)
    with pytest.raises(ValueError):
        cam.table_row()


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) needs no GPU and prints exactly one JSON line
    carrying the contract's keys; under torchrun only rank 0 prints."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                          "--cpu-sample", "4"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and "workload" in d["config"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--cpu-sample", "4"],
                         capture_output=True, text=True, timeout=600, cwd=root, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
