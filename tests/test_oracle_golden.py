"""CPU: the oracle restatement vs golden vectors produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from conftest import load_golden, relerr
from oracle import ray3d_oracle as O
from ray3d_b200 import synth
from ray3d_b200.spec import NetSpec, pos_state_entries, trj_state_entries, flops_per_sequence, weight_count

CASES = ["h36m_s1_t27", "h36m_s3_t9", "humaneva_s1_t9", "h36mcross_s2_t9", "rie_s1_t9_noembed", "rie15_s3_t27",
         "h36m_s1_t81", "h36m_s1_t243", "3dhp_s3_t243"]


def spec_of(meta, name):
    kw = dict(meta[name]["spec"])
    kw["filter_widths"] = tuple(kw["filter_widths"])
    return NetSpec(**kw)


@pytest.mark.parametrize("name", CASES)
def test_state_dict_layout_matches_reference(golden_meta, name):
    spec = spec_of(golden_meta, name)
    assert [(k, list(s)) for k, s, _ in pos_state_entries(spec)] == [tuple(e) for e in map(tuple, golden_meta[name]["keys_pos"])]
    assert [(k, list(s)) for k, s, _ in trj_state_entries(spec)] == [tuple(e) for e in map(tuple, golden_meta[name]["keys_trj"])]
    assert spec.receptive_field == golden_meta[name]["receptive_field"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_outputs(golden_meta, name):
    spec = spec_of(golden_meta, name)
    g = load_golden(name)
    sd_pos, sd_trj = synth.make_state_dicts(spec)
    # the regenerated weights are the ones the reference saw
    assert np.allclose([synth.state_digest(sd_pos), synth.state_digest(sd_trj)], g["digest"], rtol=1e-12, atol=0)
    # In the reference's .double() run the stage!=1 mix buffer is torch.zeros(...) == float32
    # (rie.py:389), which rounds the fused features to fp32: allow for it in the fp64 comparison.
    tol64 = 1e-12 if spec.stage == 1 else 5e-8
    for tag, dtype, tol in (("32", torch.float32, 2e-6), ("64", torch.float64, tol64)):
        sp, st = O.to_torch_state(sd_pos, dtype), O.to_torch_state(sd_trj, dtype)
        x = torch.from_numpy(g["x"]).to(dtype)
        p = torch.from_numpy(g["param"]).to(dtype)
        pos, trj, both = O.lift(sp, st, spec, x, p)
        assert pos.shape == g["pos" + tag].shape and trj.shape == g["trj" + tag].shape
        assert relerr(pos.numpy(), g["pos" + tag]) < tol
        assert relerr(trj.numpy(), g["trj" + tag]) < tol
        assert relerr(both.numpy(), g["pos" + tag] + g["trj" + tag]) < tol


def test_ray_encode_bit_exact_vs_reference_class(golden_meta):
    for name in ("h36m_s1_t27", "3dhp_s3_t243", "humaneva_s1_t9"):
        g = load_golden(name)
        cam = g["cam"].astype(np.float64)[:, None, None, :]
        ray = O.ray_encode(g["uv"], cam[..., 0], cam[..., 1], cam[..., 2], cam[..., 3], cam[..., 4])
        assert np.array_equal(ray, g["x64"])
        assert np.array_equal(O.ray_encode_batch(g["uv"], g["cam"]), g["x"])


def test_camera_scalars_and_encodings():
    c = load_golden("camera")
    for i in range(6):
        pitch, height = O.camera_pitch_height(c[f"R{i}"], c[f"t{i}"])
        # libm / BLAS last-ulp behaviour is host dependent; identical on the box that wrote the fixture
        assert abs(pitch - float(c[f"pitch{i}"])) <= 4 * np.spacing(abs(float(c[f"pitch{i}"])))
        assert abs(height - float(c[f"height{i}"])) <= 4 * np.spacing(abs(float(c[f"height{i}"])))
        K = c[f"K{i}"]
        ray = O.ray_encode(c[f"uv{i}"], K[0, 0], K[1, 1], K[0, 2], K[1, 2], float(c[f"pitch{i}"]))
        assert np.array_equal(ray, c[f"ray{i}"])
        assert np.array_equal(O.normalize_screen_coordinates(c[f"uv{i}"], 1000, 1002), c[f"norm{i}"])


def test_eval_windows_match_reference():
    w = load_golden("windows")
    got = O.eval_windows(torch.from_numpy(w["seq"][0]), 27)
    assert np.array_equal(got.numpy(), w["win"])


def test_work_model_matches_survey_table():
    s = NetSpec(filter_widths=(3, 3, 3, 3, 3))
    assert abs(flops_per_sequence(s) / 1e6 - 215.1) < 0.1
    assert abs(flops_per_sequence(NetSpec(filter_widths=(3, 3, 3, 3, 3), stage=3)) / 1e6 - 251.8) < 0.1
    assert abs(flops_per_sequence(NetSpec(filter_widths=(3, 3, 3))) / 1e6 - 68.0) < 0.1
    assert abs(flops_per_sequence(NetSpec(filter_widths=(3, 3, 3, 3))) / 1e6 - 104.8) < 0.1
    # params incl. BN affine + running stats; survey table counts nn.Parameters only
    assert weight_count(s) > 23_772_307 + 8_462_435


def test_eval_tail_matches_reference_losses():
    m = load_golden("metrics")
    pw = O.normalized2world(m["pred"], m["Rn2w"], m["Tn2w"])
    tw = O.normalized2world(m["target"], m["Rn2w"], m["Tn2w"])
    assert np.array_equal(pw, m["pred_world"]) and np.array_equal(tw, m["target_world"])
    got = O.eval_metrics(pw, tw)
    for k in ("mpjpe", "mrpe", "n_mpjpe", "mpjve", "p_mpjpe"):
        assert abs(got[k] - float(m[k])) <= 1e-12 * abs(float(m[k])), k


def test_undistort_matches_reference_cv2():
    """oracle.undistort_points vs the reference's CameraInfoPacket(undistort=True) (cv2.undistortPoints): bit-exact."""
    g = load_golden("camera_undistort")
    und = O.undistort_points(g["uv"], g["K"], g["dist"])
    assert np.array_equal(und, g["und"])
    pp = O.undistort_points(np.array([[g["K"][0, 2], g["K"][1, 2]]]), g["K"], g["dist"])
    assert np.array_equal(pp.reshape(-1), g["pp_cam"].reshape(-1))


def test_flip_tta_matches_reference_evaluate_core():
    """oracle.lift_tta vs the reference's own Trainer.evaluate_core(flip_test=True) (trainer.py:283-405) run on one
    synthetic video by tests/golden/make_golden.py: the tensor evaluate_core hands to normalized2world (pos_tta + trj_tta)
    and the five metrics it returns."""
    g = load_golden("evaluate_core_tta")
    spec = NetSpec(num_joints=17, in_features=3, filter_widths=(3, 3, 3), stage=1)
    sp, st = synth.make_state_dicts(spec)
    win = O.eval_windows(torch.from_numpy(g["rays"]), spec.receptive_field)                    # trainer.py:47-58
    prm = torch.from_numpy(np.tile(g["param"], (win.shape[0], 1)))                             # trainer.py:324
    for tag, dtype, tol in (("32", torch.float32, 2e-6), ("64", torch.float64, 1e-12)):
        got = O.lift_tta(O.to_torch_state(sp, dtype), O.to_torch_state(st, dtype), spec, win.to(dtype), prm.to(dtype),
                         g["kps_left"].tolist(), g["kps_right"].tolist())[2].numpy()
        assert got.shape == g["pred" + tag].shape
        assert relerr(got, g["pred" + tag]) < tol, tag
    # the metrics of evaluate_core (millimetres): normalized2world on prediction and target, then lib/loss/loss.py
    pw = O.normalized2world(g["pred32"], g["Rn2w"], g["Tn2w"])
    tw = O.normalized2world(g["target"][:, None], g["Rn2w"], g["Tn2w"])
    m = O.eval_metrics(pw, tw)
    e1, e2, e3, ev, er = g["metrics32"]
    for want, key in ((e1, "mpjpe"), (e2, "p_mpjpe"), (e3, "n_mpjpe"), (ev, "mpjve"), (er, "mrpe")):
        assert abs(m[key] * 1000 - want) <= 1e-9 * abs(want), key
