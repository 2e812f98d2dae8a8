"""CPU: host side of the C ABI -- symbols, weight folding/packing, launch-graph wiring, error behaviour."""
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, relerr
from replay import replay
from ray3d_b200 import _capi, synth
from ray3d_b200.spec import NetSpec, BN_EPS

CASES = ["h36m_s1_t27", "h36m_s3_t9", "humaneva_s1_t9", "h36mcross_s2_t9", "rie_s1_t9_noembed", "rie15_s3_t27"]


def spec_of(meta, name):
    kw = dict(meta[name]["spec"])
    kw["filter_widths"] = tuple(kw["filter_widths"])
    return NetSpec(**kw)


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ray3d_b200.h")).read()
    declared = set(re.findall(r"R3D_API[^;(]*?\b(r3d_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    assert declared == set(_capi.EXPORTS), declared ^ set(_capi.EXPORTS)
    lib = _capi.lib()                      # loads the .so and resolves every EXPORTS entry
    for name in declared:
        assert hasattr(lib, name)
    assert lib.r3d_abi_version() == 2


def make_plan(spec, nets=3, precision="fp32"):
    sp, st = synth.make_state_dicts(spec)
    p = _capi.Plan(spec, nets, precision)
    if nets & 1:
        p.load_state(_capi.NET_POS, sp)
    if nets & 2:
        p.load_state(_capi.NET_TRJ, st)
    p.finalize()
    return p, sp, st


def test_bn_fold_and_packing_match_numpy():
    spec = NetSpec(filter_widths=(3, 3, 3), stage=3)
    p, sp, st = make_plan(spec)

    def fold(sd, w, bn, bias=None):
        W = sd[w].astype(np.float64)
        n = W.shape[0]
        if bn:
            s = sd[bn + ".weight"].astype(np.float64) / np.sqrt(sd[bn + ".running_var"].astype(np.float64) + BN_EPS)
            sh = sd[bn + ".bias"].astype(np.float64) - sd[bn + ".running_mean"].astype(np.float64) * s
        else:
            s, sh = np.ones(n), np.zeros(n)
        b = (sd[bias].astype(np.float64) * s if bias else 0) + sh
        if W.ndim == 3:      # (out, in, tap) -> column tap*in + c
            W = W.transpose(0, 2, 1).reshape(n, -1)
        return (W * s[:, None]).astype(np.float32), np.asarray(b, np.float32)

    checks = [
        (1, "LocalLayer_LLeg.layers_conv.2", sp, "LocalLayer_LLeg.layers_conv.2.weight", "LocalLayer_LLeg.layers_bn.2", None),
        (1, "LocalLayer_RArm.layers_conv.1", sp, "LocalLayer_RArm.layers_conv.1.weight", "LocalLayer_RArm.layers_bn.1", None),
        (1, "LocalLayer_RArm.shrink", sp, "LocalLayer_RArm.shrink.weight", None, "LocalLayer_RArm.shrink.bias"),
        (1, "GlobalInfo.fc_1", sp, "GlobalInfo.fc_1.weight", "GlobalInfo.bn_1", "GlobalInfo.fc_1.bias"),
        (1, "FuseBlocks.3.layers.0.w2", sp, "FuseBlocks.3.layers.0.w2.weight", "FuseBlocks.3.layers.0.batch_norm2", "FuseBlocks.3.layers.0.w2.bias"),
        (1, "Integration_Torso.fc_2", sp, "Integration_Torso.fc_2.weight", None, "Integration_Torso.fc_2.bias"),
        (2, "Integration.fc_2", st, "Integration.fc_2.weight", None, "Integration.fc_2.bias"),
        (2, "embedder.w2", st, "embedder.w2.weight", "embedder.b2", "embedder.w2.bias"),
    ]
    for net, layer, sd, w, bn, bias in checks:
        W, b = fold(sd, w, bn, bias)
        pw, pb = p.packed_layer(net, layer)
        n, k = W.shape
        assert pw.shape[0] >= n and pw.shape[1] >= k and pw.shape[1] % 64 in (0, k % 64)
        assert np.array_equal(pw[:n, :k], W), layer
        assert np.array_equal(pb[:n], b), layer
        assert not pw[n:].any() and not pw[:, k:].any() and not pb[n:].any()


def test_folded_expand_conv_equals_grouped_conv():
    """expand_conv over [x_g | x_g - root | x_g - x_g[tc]] (rie.py:301-357, :86) == folded weights applied to the
    shared operand [w0 frames | x[tc]]: checked in float64 on random inputs for every group and the trajectory net."""
    from ray3d_b200.spec import GROUP_JOINTS
    spec = NetSpec(filter_widths=(3, 3, 3), stage=1)
    p, sp, st = make_plan(spec)
    rng = np.random.default_rng(0)
    T, J, C = 27, 17, 3
    x = rng.standard_normal((4, T, J * C))
    tc, w0 = T // C, 3
    from replay import first_layer_operand
    g = p.describe()
    A = first_layer_operand(x, g)                       # shared operand, columns grouped per joint group (a0_map)
    assert A.shape[2] == 256 and sorted(m for m in g["a0_map"] if m >= 0 and m >= w0 * J * C) == sorted(
        [w0 * J * C + i for i in range(J * C)] + [w0 * J * C + c for c in range(C)] * 4)   # x[tc] once per joint + 4 root copies
    for net, sd, pre, joints in [(1, sp, "LocalLayer_" + g, GROUP_JOINTS[17][g]) for g in ("Torso", "LArm", "RArm", "LLeg", "RLeg")] + \
                                [(2, st, "LocalLayer", tuple(range(17)))]:
        idx = [j * C + c for j in joints for c in range(C)]
        xg = x[:, :, idx]
        inp = np.concatenate([xg, xg - np.tile(x[:, :, :C], (1, 1, len(joints))), xg - x[:, tc:tc + 1, idx]], axis=2)   # (B,T,Cg)
        W = sd[pre + ".expand_conv.weight"].astype(np.float64)
        s = sd[pre + ".expand_bn.weight"].astype(np.float64) / np.sqrt(sd[pre + ".expand_bn.running_var"].astype(np.float64) + BN_EPS)
        sh = sd[pre + ".expand_bn.bias"].astype(np.float64) - sd[pre + ".expand_bn.running_mean"].astype(np.float64) * s
        ref = np.einsum("bqkc,ock->bqo", inp.reshape(4, T // w0, w0, -1), W) * s + sh
        pw, pb = p.packed_layer(net, pre + ".expand_conv")
        assert pw.shape == (256, 256)
        got = A @ pw.astype(np.float64).T + pb.astype(np.float64)
        # zero-step elision: a limb problem keeps 3 of the 16 K steps, the Torso 4, the trajectory net all 16
        steps = int((np.abs(pw).reshape(256, 16, 16).max(axis=(0, 2)) > 0).sum())
        assert steps == (16 if net == 2 else 4 if pre.endswith("Torso") else 3), (pre, steps)
        assert relerr(got, ref) < 2e-7, pre          # weights are rounded to fp32 after the float64 fold


@pytest.mark.parametrize("name", CASES)
def test_first_layer_layout_and_k_step_masks(golden_meta, name):
    """Column layout of the shared first-layer operand (a0_map) and the zero-step elision, for every joint set
    (17/15/14) and input width (rays / pixels): each window column is used exactly once plus one root copy per limb slot,
    every slot starts on a 16-column K step, and the K-step count the plan reports for a layer is exactly the number of
    16-column steps of its packed weights that hold a non-zero -- limb problems touch 3-4 steps, never the whole row."""
    from ray3d_b200.spec import GROUP_JOINTS
    spec = spec_of(golden_meta, name)
    p, sp, st = make_plan(spec, precision="bf16x3")
    g = p.describe()
    J, C, w0 = g["J"], g["Cin"], g["w0"]
    amap = np.asarray(g["a0_map"])
    assert len(amap) % 64 == 0
    used = amap[amap >= 0]
    want = list(range((w0 + 1) * J * C)) + [t * J * C + c for t in range(w0 + 1) for c in range(C)] * 4   # + 4 root copies
    assert sorted(used.tolist()) == sorted(want)
    slot_starts = [0]
    for grp in ("Torso", "LArm", "RArm", "LLeg", "RLeg"):
        n = (len(GROUP_JOINTS[J][grp]) + (grp != "Torso")) * (w0 + 1) * C
        slot_starts.append((slot_starts[-1] + n + 15) // 16 * 16)
    assert all(s0 % 16 == 0 for s0 in slot_starts) and slot_starts[-1] <= len(amap)
    first = g["ops"][0]
    assert first["name"] == "expand_conv"
    for op in g["ops"]:
        for pr in op["prob"]:
            net, layer = pr["layer"].split(":", 1)
            pw, _ = p.packed_layer(1 << int(net), layer)
            steps = int((np.abs(pw).reshape(pw.shape[0], -1, 16).max(axis=(0, 2)) > 0).sum())
            assert pr["k_steps"] == max(steps, 1), (op["name"], layer)
    limb_steps = [pr["k_steps"] for pr in first["prob"] if "LocalLayer_L" in pr["layer"] or "LocalLayer_R" in pr["layer"]]
    assert limb_steps and max(limb_steps) <= -(-(4 * (w0 + 1) * C) // 16) + 1 < len(amap) // 16


@pytest.mark.parametrize("name", CASES)
def test_launch_graph_replay_matches_reference(golden_meta, name):
    """Wiring + packing + gather tables: numpy replay of the native graph vs the reference's outputs."""
    spec = spec_of(golden_meta, name)
    g = load_golden(name)
    p, _, _ = make_plan(spec)
    pos, trj = replay(p, g["x"], g["param"])
    assert relerr(pos, g["pos64"]) < 5e-6
    assert relerr(trj, g["trj64"]) < 5e-6
    assert p.receptive_field == spec.receptive_field
    assert p.kernel_launches == len(p.describe()["ops"]) + 2


@pytest.mark.parametrize("name", ["h36m_s1_t27", "rie15_s3_t27"])
def test_fused_conv_pair_graph_replay(golden_meta, name):
    """Tensor-core plans run a level's k=w conv + 1x1 conv as one fused launch when the level keeps >= 3 rows per window
    (the single-row top level stays two narrow-tile launches): same function, fewer ops."""
    spec = spec_of(golden_meta, name)
    g = load_golden(name)
    p, _, _ = make_plan(spec, precision="bf16x3")
    ops = p.describe()["ops"]
    fused = [o for o in ops if "+" in o["name"]]
    rows, want = spec.receptive_field, 0
    for i, w in enumerate(spec.filter_widths):
        rows //= w
        want += i >= 1 and rows >= 3
    assert len(fused) == want >= 1 and all("layer2" in q for o in fused for q in o["prob"])
    pos, trj = replay(p, g["x"], g["param"])
    assert relerr(pos, g["pos64"]) < 5e-6 and relerr(trj, g["trj64"]) < 5e-6
    p32, _, _ = make_plan(spec, precision="fp32")
    assert len(p32.describe()["ops"]) == len(ops) + len(fused)


@pytest.mark.parametrize("nets", [1, 2])
def test_single_net_plans_replay(golden_meta, nets):
    spec = spec_of(golden_meta, "h36m_s3_t9")
    g = load_golden("h36m_s3_t9")
    p, _, _ = make_plan(spec, nets=nets)
    pos, trj = replay(p, g["x"], g["param"])
    if nets == 1:
        assert trj is None and relerr(pos, g["pos64"]) < 5e-6
    else:
        assert pos is None and relerr(trj, g["trj64"]) < 5e-6


def test_error_behaviour():
    spec = NetSpec(filter_widths=(3, 3))
    sp, st = synth.make_state_dicts(spec)
    p = _capi.Plan(spec, 3, "fp32")
    with pytest.raises(_capi.R3DError, match="unknown state_dict key"):
        p.set_tensor(_capi.NET_POS, "LocalLayer_Nose.expand_conv.weight", np.zeros((2, 2), np.float32))
    with pytest.raises(_capi.R3DError, match="shape mismatch"):
        p.set_tensor(_capi.NET_POS, "LocalLayer_Torso.expand_conv.weight", np.zeros((256, 45, 5), np.float32))
    p.load_state(_capi.NET_POS, sp)
    with pytest.raises(_capi.R3DError, match="never supplied") as ei:    # trj weights missing
        p.finalize()
    assert ei.value.code == _capi.ERR_MISSING_WEIGHT
    # DataParallel-style prefixes are accepted (trainer.py:232-240)
    p.load_state(_capi.NET_TRJ, {"module." + k: v for k, v in st.items()})
    p.finalize()
    if not torch.cuda.is_available():
        with pytest.raises(_capi.R3DError) as ei:
            p.upload(0)
        assert ei.value.code == _capi.ERR_NO_DEVICE     # loud, no CPU fallback
    for bad in (dict(num_joints=16), dict(in_features=4), dict(filter_widths=(2, 3))):
        with pytest.raises(ValueError):
            NetSpec(**bad)
    cfg = _capi.make_config(spec, 3, "fp32")
    cfg.channels = 100
    h = _capi.C.c_void_p()
    assert _capi.lib().r3d_plan_create(_capi.C.byref(cfg), _capi.C.byref(h)) == _capi.ERR_UNSUPPORTED
    assert b"multiples of 64" in _capi.lib().r3d_last_error()


def test_tensor_precision_plans_finalize():
    spec = NetSpec(filter_widths=(3, 3))
    for prec in ("bf16x3", "bf16"):
        p, _, _ = make_plan(spec, precision=prec)
        assert p.weight_bytes > 0
    p32, _, _ = make_plan(spec, precision="fp32")
    pb, _, _ = make_plan(spec, precision="bf16")
    assert pb.weight_bytes < 0.6 * p32.weight_bytes
