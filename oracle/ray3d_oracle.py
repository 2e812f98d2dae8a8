"""ORACLE -- CPU restatement of the Ray3D lifting forward pass.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs may import it.  Nothing under
``ray3d_b200/`` imports it; the product path fails loudly when the CUDA library is missing.

Pinning status: the reference ships no tests, golden vectors or checkpoints (SURVEY.md section 4,
8c) => parity is *unpinned by the reference's own fixtures*.  It is pinned instead against outputs
of the unmodified reference modules run in the build container on seeded synthetic weights
(``tests/golden/make_golden.py`` imports /root/reference and writes ``tests/golden/*.npz``);
``tests/test_oracle_golden.py`` checks this restatement against those vectors.

Third-party arithmetic: the reference's math lives in PyTorch (pinned torch==1.4.0+cu100,
requirements.txt:78; this image has 2.11.0) -- conv1d / linear / batch_norm / leaky_relu, whose
semantics are stable; summation order is not, hence tolerance-based parity -- and in OpenCV's
cv2.undistortPoints (camera.py:420, only when undistort=True; out of scope).

The restatement is written with the same torch.nn.functional ops the reference's nn.Modules
dispatch to, so timing it on host cores is a faithful CPU baseline for the reference.
Every function cites the reference lines it follows (paths relative to the reference root).
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from ray3d_b200.spec import GROUPS, NetSpec, BN_EPS, SLOPE_EMBED, SLOPE_NET

Tensor = torch.Tensor
State = Mapping[str, Tensor]


# ------------------------------------------------------------------------------------------
# lib/camera/camera.py -- host-side camera arithmetic (float64 numpy, like the reference)
# ------------------------------------------------------------------------------------------
def normalize_screen_coordinates(X: np.ndarray, w: float, h: float) -> np.ndarray:
    """camera.py:11-18 -- map [0,w] to [-1,1] keeping the aspect ratio."""
    assert X.shape[-1] == 2
    return X / w * 2 - np.array([1, h / w])


def camera_pitch_height(R: np.ndarray, t: np.ndarray) -> Tuple[float, float]:
    """camera.py:245-259,285-316 -- pitch = angle(Rc2w @ e_z, e_z) - pi/2, height = (-R^T t)[2].

    ``angle`` (camera.py:198-205) is acos(<a,b> / (|a||b|)) with python-float arithmetic.
    """
    R = np.asarray(R, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64).reshape(3, 1)
    ray_world = (R.T @ np.array([0.0, 0.0, 1.0])).reshape(3)
    dot = float(ray_world[2])                          # <ray_world, (0,0,1)>
    len_ray = math.sqrt(float(sum(c * c for c in ray_world)))
    pitch = math.acos(dot / (len_ray * 1.0)) - np.pi / 2
    height = float((-R.T @ t)[2, 0])
    return pitch, height


def undistort_points(uv: np.ndarray, K: np.ndarray, dist: Sequence[float]) -> np.ndarray:
    """camera.py:412-421: cv2.undistortPoints(points, K, dist_coeff, P=K) for the 5-coefficient model.
    Third-party arithmetic (OpenCV, pinned opencv-python==4.4.0.42 in requirements.txt:40; pinned here against
    4.13.0 of this image through tests/golden/camera_undistort.npz): x=(u-cx)/fx by reciprocal multiply, 5 fixed-point
    iterations x <- (x0 - dx(x,y)) / (1 + k1 r2 + k2 r2^2 + k3 r2^3), then back through P=K."""
    uv = np.asarray(uv, dtype=np.float64)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    k1, k2, p1, p2, k3 = (np.float64(c) for c in dist)
    ifx, ify = 1.0 / fx, 1.0 / fy
    x = (uv[..., 0] - cx) * ifx
    y = (uv[..., 1] - cy) * ify
    x0, y0 = x.copy(), y.copy()
    for _ in range(5):
        r2 = x * x + y * y
        icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2)
        dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x)
        dy = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y
        x = (x0 - dx) * icdist
        y = (y0 - dy) * icdist
    return np.stack([fx * x + cx, fy * y + cy], axis=-1)


def ray_encode(uv: np.ndarray, fx, fy, cx, cy, pitch) -> np.ndarray:
    """camera.py:423-441 + 460-471 (+ Rc2n from 325-345), undistort=False.

    uv (..., J, 2) pixel coordinates; intrinsics/pitch broadcast against uv[..., 0].  Returns
    float64 (..., J, 3) = [xn, yn, 1] @ Rx(pitch)^T, i.e. (xn, c*yn + s, -s*yn + c).
    """
    uv = np.asarray(uv, dtype=np.float64)
    fx, fy, cx, cy, pitch = (np.asarray(a, dtype=np.float64) for a in (fx, fy, cx, cy, pitch))
    xn = (uv[..., 0] - cx) / fx
    yn = (uv[..., 1] - cy) / fy
    c = np.cos(pitch)
    s = np.sin(pitch)
    out = np.empty(uv.shape[:-1] + (3,), dtype=np.float64)
    out[..., 0] = xn
    out[..., 1] = c * yn + s
    out[..., 2] = -s * yn + c
    return out


def ray_encode_batch(uv: np.ndarray, cam: np.ndarray) -> np.ndarray:
    """Batched form used by the benchmark: uv (B,T,J,2), cam (B,6)=[fx,fy,cx,cy,pitch,height].
    Same arithmetic as ``ray_encode`` per sequence; result cast like trainer.py:298."""
    c = np.asarray(cam, dtype=np.float64)[:, None, None, :]
    return ray_encode(uv, c[..., 0], c[..., 1], c[..., 2], c[..., 3], c[..., 4]).astype(np.float32)


# ------------------------------------------------------------------------------------------
# lib/model/rie.py + lib/model/embedding.py -- eval-mode forward (Dropout == identity)
# ------------------------------------------------------------------------------------------
def _bn(sd: State, p: str, x: Tensor) -> Tensor:
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.1, BN_EPS)


def _act(x: Tensor, slope: float = SLOPE_NET) -> Tensor:
    return F.leaky_relu(x, slope)


def temporal_block(sd: State, p: str, x: Tensor, widths: Sequence[int]) -> Tensor:
    """rie.py:85-105 with Optimize1f=True, causal=False: (B, Cin, T) -> (B*T_out, 1, latent)."""
    x = _act(_bn(sd, p + ".expand_bn", F.conv1d(x, sd[p + ".expand_conv.weight"], None, stride=widths[0])))
    for i in range(len(widths) - 1):
        w = widths[i + 1]
        res = x[:, :, w // 2::w]                                                       # rie.py:94
        x = _act(_bn(sd, p + f".layers_bn.{2 * i}", F.conv1d(x, sd[p + f".layers_conv.{2 * i}.weight"], None, stride=w)))
        x = res + _act(_bn(sd, p + f".layers_bn.{2 * i + 1}", F.conv1d(x, sd[p + f".layers_conv.{2 * i + 1}.weight"])))
    x = F.conv1d(x, sd[p + ".shrink.weight"], sd[p + ".shrink.bias"])                   # rie.py:99
    x = x.permute(0, 2, 1)
    return x.reshape(x.shape[0] * x.shape[1], x.shape[2]).unsqueeze(1)


def fc_block(sd: State, p: str, x: Tensor, nblocks: int) -> Tensor:
    """rie.py:159-169 (FCBlock) with rie.py:122-135 (Linear residual blocks)."""
    x = _act(_bn(sd, p + ".bn_1", F.linear(x, sd[p + ".fc_1.weight"], sd[p + ".fc_1.bias"])))
    for i in range(nblocks):
        q = p + f".layers.{i}"
        y = _act(_bn(sd, q + ".batch_norm1", F.linear(x, sd[q + ".w1.weight"], sd[q + ".w1.bias"])))
        y = _act(_bn(sd, q + ".batch_norm2", F.linear(y, sd[q + ".w2.weight"], sd[q + ".w2.bias"])))
        x = x + y
    return F.linear(x, sd[p + ".fc_2.weight"], sd[p + ".fc_2.bias"])


def embedding(sd: State, p: str, x: Tensor) -> Tensor:
    """embedding.py:15-19 -- LeakyReLU() default slope 0.01."""
    x = _act(_bn(sd, p + ".b1", F.linear(x, sd[p + ".w1.weight"], sd[p + ".w1.bias"])), SLOPE_EMBED)
    return _act(_bn(sd, p + ".b2", F.linear(x, sd[p + ".w2.weight"], sd[p + ".w2.bias"])), SLOPE_EMBED)


def _preamble(spec: NetSpec, x: Tensor):
    """rie.py:289-304 == rie.py:523-538."""
    assert x.dim() == 4 and x.shape[-2] == spec.num_joints and x.shape[-1] == spec.in_features
    cin = spec.in_features
    tc = x.shape[1] // cin
    in_current = x[:, tc:tc + 1].reshape(x.shape[0], -1)
    xc = x.reshape(x.shape[0], x.shape[1], -1).permute(0, 2, 1)           # (B, J*Cin, T)
    diff = xc - xc[:, 0:cin, :].repeat(1, xc.shape[1] // cin, 1)         # root-relative
    diff_t = xc - xc[:, :, tc:tc + 1]                                    # relative to "current" frame
    return in_current, xc, diff, diff_t


def _group_channels(spec: NetSpec, g: str):
    cin = spec.in_features
    return [j * cin + c for j in spec.group_joints(g) for c in range(cin)]


def pos_forward(sd: State, spec: NetSpec, x: Tensor, param: Tensor) -> Tensor:
    """RIEModel.forward, rie.py:284-434.  x (B,T,J,Cin), param (B,extrinsic_dim) -> (B,1,J,3)."""
    B, T = x.shape[0], x.shape[1]
    in_current, xc, diff, diff_t = _preamble(spec, x)
    x_global = fc_block(sd, "GlobalInfo", in_current, 2)                                # rie.py:362
    latents = []
    for g in GROUPS:                                                                    # rie.py:306-369
        idx = _group_channels(spec, g)
        xin = torch.cat((xc[:, idx], diff[:, idx], diff_t[:, idx]), dim=1)
        latents.append(temporal_block(sd, f"LocalLayer_{g}", xin, spec.filter_widths))
    tmp = torch.cat(latents, dim=1)                                                     # (B,5,L)  rie.py:371
    tail = [x_global]
    if spec.camera_embedding:
        tail.append(embedding(sd, "embedder", param))
    if spec.stage == 1:                                                                 # rie.py:373-386
        feats = [torch.cat([tmp[:, i]] + tail, dim=1) for i in range(5)]
    else:                                                                               # rie.py:388-407
        mix = []
        for i in range(5):
            others = torch.cat((tmp[:, :i, :], tmp[:, i + 1:, :]), dim=1).reshape(tmp.shape[0], spec.latent * 4)
            mix.append(fc_block(sd, f"FuseBlocks.{i}", others, 1))
        feats = [torch.cat([tmp[:, i], mix[i]] + tail, dim=1) for i in range(5)]
    heads = {}
    for i, g in enumerate(GROUPS):                                                      # rie.py:410-424
        heads[g] = fc_block(sd, f"Integration_{g}", feats[i], 1).view(tmp.shape[0], -1, 3)
    out = torch.stack([heads[g][:, k] for g, k in spec.output_slots()], dim=1)          # rie.py:426-431
    pad = (spec.receptive_field - 1) // 2
    return out.view(B, T - 2 * pad, spec.num_joints, 3)


def trj_forward(sd: State, spec: NetSpec, x: Tensor, param: Tensor) -> Tensor:
    """RIETrajectoryModel.forward, rie.py:518-559 -> (B,1,1,3)."""
    B, T = x.shape[0], x.shape[1]
    in_current, xc, diff, diff_t = _preamble(spec, x)
    x_local = temporal_block(sd, "LocalLayer", torch.cat((xc, diff, diff_t), dim=1), spec.filter_widths)
    parts = [x_local[:, 0], fc_block(sd, "GlobalInfo", in_current, 2)]
    if spec.camera_embedding:
        parts.append(embedding(sd, "embedder", param))
    out = fc_block(sd, "Integration", torch.cat(parts, dim=1), 1)
    pad = (spec.receptive_field - 1) // 2
    return out.view(B, T - 2 * pad, 1, 3)


def lift(sd_pos: State, sd_trj: State, spec: NetSpec, x: Tensor, param: Tensor):
    """Eval-loop composition trainer.py:337,346,353: returns (pos, trj, pos + trj)."""
    with torch.no_grad():
        pos = pos_forward(sd_pos, spec, x, param)
        trj = trj_forward(sd_trj, spec, x, param)
        return pos, trj, pos + trj


def lift_tta(sd_pos: State, sd_trj: State, spec: NetSpec, x: Tensor, param: Tensor, kps_left, kps_right):
    """Flip test-time augmentation of Trainer.evaluate_core (trainer.py:299-302, 337-353): returns (pos, trj, pos+trj)
    after averaging the direct and the un-mirrored mirrored predictions."""
    kl, kr = list(kps_left), list(kps_right)
    with torch.no_grad():
        xf = x.clone()
        xf[:, :, :, 0] *= -1                                                # trainer.py:301
        xf[:, :, kl + kr, :] = xf[:, :, kr + kl, :]                         # trainer.py:302
        pos = pos_forward(sd_pos, spec, x, param)
        pos_f = pos_forward(sd_pos, spec, xf, param)
        pos_f[:, :, :, 0] *= -1                                             # trainer.py:340
        pos_f[:, :, kl + kr] = pos_f[:, :, kr + kl]                         # trainer.py:341-342
        pos = torch.mean(torch.cat((pos, pos_f), dim=1), dim=1, keepdim=True)   # trainer.py:343-345
        trj = trj_forward(sd_trj, spec, x, param)
        trj_f = trj_forward(sd_trj, spec, xf, param)
        trj_f[:, :, :, 0] *= -1                                             # trainer.py:349
        trj = torch.mean(torch.cat((trj, trj_f), dim=1), dim=1, keepdim=True)   # trainer.py:350-352
        return pos, trj, pos + trj                                          # trainer.py:353


def lift_uv(sd_pos: State, sd_trj: State, spec: NetSpec, uv: np.ndarray, cam: np.ndarray):
    """Full hot path from pixels: ray encode (float64 -> float32, trainer.py:298) + both nets.
    cam (B,6) = [fx, fy, cx, cy, pitch, height]; param = [height, pitch] (trainer.py:297)."""
    dtype = next(iter(sd_pos.values())).dtype
    x = torch.from_numpy(ray_encode_batch(uv, cam)).to(dtype)
    param = torch.from_numpy(np.ascontiguousarray(np.asarray(cam, dtype=np.float32)[:, [5, 4]])).to(dtype)
    return lift(sd_pos, sd_trj, spec, x, param)


def normalized2world(pt: np.ndarray, Rn2w: np.ndarray, Tn2w: np.ndarray) -> np.ndarray:
    """camera.py:401-410 (numpy branch): pt @ Rn2w.T + Tn2w.T."""
    return pt @ Rn2w.T + Tn2w.T


def eval_metrics(pred_world: np.ndarray, target_world: np.ndarray) -> Dict[str, float]:
    """lib/loss/loss.py mpjpe (:12-18), n_mpjpe (:72-81), mean_velocity_error (:95-104) as called at
    trainer.py:386-395 on (F,1,J,3) float64 world coordinates."""
    p, t = torch.from_numpy(np.asarray(pred_world)), torch.from_numpy(np.asarray(target_world))
    mp = torch.mean(torch.norm(p - t, dim=3)).item()
    mr = torch.mean(torch.norm(p[:, :, 0:1] - t[:, :, 0:1], dim=3)).item()
    norm_p = torch.mean(torch.sum(p ** 2, dim=3, keepdim=True), dim=2, keepdim=True)
    norm_t = torch.mean(torch.sum(t * p, dim=3, keepdim=True), dim=2, keepdim=True)
    nm = torch.mean(torch.norm(norm_t / norm_p * p - t, dim=3)).item()
    pn, tn = p.numpy().reshape(-1, p.shape[-2], 3), t.numpy().reshape(-1, p.shape[-2], 3)
    mv = float(np.mean(np.linalg.norm(np.diff(pn, axis=0) - np.diff(tn, axis=0), axis=2)))
    return {"mpjpe": mp, "mrpe": mr, "n_mpjpe": nm, "mpjve": mv, "p_mpjpe": p_mpjpe(pn, tn)}


def p_mpjpe(predicted: np.ndarray, target: np.ndarray) -> float:
    """lib/loss/loss.py:30-69: MPJPE after per-frame similarity alignment (orthogonal Procrustes through the SVD of
    X0^T Y0, reflection fix on the last singular vector, scale = trace * |X0| / |Y0|).  (F, J, 3) float64."""
    muX, muY = target.mean(axis=1, keepdims=True), predicted.mean(axis=1, keepdims=True)
    X0, Y0 = target - muX, predicted - muY
    normX = np.sqrt((X0 ** 2).sum(axis=(1, 2), keepdims=True))
    normY = np.sqrt((Y0 ** 2).sum(axis=(1, 2), keepdims=True))
    U, s, Vt = np.linalg.svd(np.matmul((X0 / normX).transpose(0, 2, 1), Y0 / normY))
    V = Vt.transpose(0, 2, 1).copy()
    sign = np.sign(np.linalg.det(np.matmul(V, U.transpose(0, 2, 1))))
    V[:, :, -1] *= sign[:, None]
    s[:, -1] *= sign
    R = np.matmul(V, U.transpose(0, 2, 1))
    a = s.sum(axis=1)[:, None, None] * normX / normY
    aligned = a * np.matmul(predicted, R) + (muX - a * np.matmul(muY, R))
    return float(np.mean(np.linalg.norm(aligned - target, axis=2)))


def eval_windows(seq: Tensor, receptive_field: int) -> Tensor:
    """trainer.py:47-58 eval_data_prepare: (F+RF-1, J, C) -> (F, RF, J, C) sliding windows."""
    return seq.unfold(0, receptive_field, 1).permute(0, 3, 1, 2).contiguous()


def to_torch_state(sd_np: Mapping[str, np.ndarray], dtype=torch.float32) -> Dict[str, Tensor]:
    out = {}
    for k, v in sd_np.items():
        t = torch.from_numpy(np.asarray(v))
        out[k] = t if t.dtype == torch.int64 else t.to(dtype)
    return out
