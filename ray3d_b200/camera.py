"""Camera-side host interface of the lifting path.

Mirrors the slice of ``CameraInfoPacket`` that is on the hot path:

  normalize_screen_coordinates ... lib/camera/camera.py:11-18
  pitch / height / Rc2n .......... lib/camera/camera.py:245-259, 285-345
  encode_uv_with_intrinsic ....... lib/camera/camera.py:423-441
  undistort_point ................ lib/camera/camera.py:412-421   (cv2.undistortPoints, 5 coefficients)
  get_cam_ray_given_uv ........... lib/camera/camera.py:460-471

The per-camera scalars (pitch, height) are host float64 math, computed once per camera exactly
like the reference; the per-keypoint arithmetic runs on the GPU in float64 (bit-identical to the
reference's numpy results) through the C ABI.  In the fused forward (``Lifter.forward_uv``) the
same encode is done inside the input kernel, so this standalone form is only needed when a caller
wants the encoded rays themselves (lib/dataset/__init__.py:191-203).
"""
from __future__ import annotations

import math
from typing import Optional, Union

import numpy as np
import torch

from . import _capi

ArrayLike = Union[np.ndarray, torch.Tensor]


def _to_cuda_f64(a: ArrayLike, device) -> torch.Tensor:
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: ray3d_b200 has no CPU path")
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)) if isinstance(a, np.ndarray) else a
    return t.to(device=device, dtype=torch.float64).contiguous()


def _back(t: torch.Tensor, like: ArrayLike):
    return t.cpu().numpy() if isinstance(like, np.ndarray) else t


def normalize_screen_coordinates(X: ArrayLike, w: float, h: float, device: Optional[int] = None):
    """camera.py:11-18 on the GPU (float64): X / w * 2 - [1, h / w]."""
    assert X.shape[-1] == 2
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    x = _to_cuda_f64(X, dev)
    out = torch.empty_like(x)
    with torch.cuda.device(dev):
        _capi.check(_capi.lib().r3d_normalize_screen_f64(x.data_ptr(), out.data_ptr(), x.numel() // 2, float(w), float(h),
                                                          torch.cuda.current_stream(dev).cuda_stream))
    return _back(out, X)


class RayCamera:
    """The subset of CameraInfoPacket (camera.py:208-504) the lifting path needs.

    One must supply K, R, t like the reference (P is never used on this path).  With ``undistort=True``
    keypoints and the principal point first go through the lens undistortion of camera.py:412-421
    (cv2.undistortPoints(pts, K, dist_coeff, P=K), 5-coefficient model) -- on the device, bit-identical to
    opencv-python 4.13."""

    def __init__(self, K, R, t, res_w=None, res_h=None, undistort=False, dist_coeff=None):
        if undistort:
            if dist_coeff is None or np.asarray(dist_coeff).size != 5:
                raise ValueError("undistort=True needs the 5 distortion coefficients (k1, k2, p1, p2, k3)")
        self.dist_coeff = None if dist_coeff is None else np.asarray(dist_coeff, dtype=np.float64).reshape(-1)
        K = np.asarray(K, dtype=np.float64)
        R = np.asarray(R, dtype=np.float64)
        t = np.asarray(t, dtype=np.float64).reshape(3, 1)
        assert K.shape == (3, 3) and R.shape == (3, 3)
        self.K, self.Rw2c, self.Tw2c = K, R, t
        self.res_w, self.res_h, self.undistort = res_w, res_h, bool(undistort)
        self.Rc2w = R.T
        # camera.py:273-283,308-316: optical axis in world coordinates vs world up
        # (Rc2w @ e_z is exactly the third column of Rc2w: index it instead of going through BLAS)
        ray_world = self.Rc2w[:, 2]
        norm = math.sqrt(sum(float(c) * float(c) for c in ray_world))
        self.cam_pitch_rad = math.acos(float(ray_world[2]) / (norm * 1.0)) - np.pi / 2
        self.cam_orig_world = -self.Rw2c.T @ self.Tw2c                  # camera.py:265-271
        self.height = float(self.cam_orig_world[2, 0])                   # trainer.py:297
        c, s = math.cos(self.cam_pitch_rad), math.sin(self.cam_pitch_rad)
        self.Rc2n = np.array([[1.0, 0.0, 0.0], [0.0, c, s], [0.0, -s, c]])   # camera.py:333-338
        if self.undistort:                                               # camera.py:253-256
            self.pp_cam = self._undistort_host(K[0, 2], K[1, 2]).reshape(1, 2)
        else:
            self.pp_cam = np.array([[K[0, 2], K[1, 2]]])                 # camera.py:258-259
        # normalised frame <-> world (camera.py:246-256): Tc2n = (0, -height, 0)
        Tc2n = np.array([[0.0], [-self.height], [0.0]])
        self.Rn2w = self.Rc2w @ self.Rc2n.T
        self.Tn2w = -self.Rn2w @ Tc2n - self.Rc2w @ self.Tw2c

    @property
    def param(self) -> np.ndarray:
        """[height, pitch] -- the model's second input (trainer.py:297)."""
        return np.array([self.height, self.cam_pitch_rad], dtype=np.float32)

    def table_row(self) -> np.ndarray:
        """float32 [fx, fy, cx, cy, pitch, height] row for Lifter.forward_uv (pinhole cameras whose calibration is
        float32-exact; real calibration doubles and distorted lenses go through table_row64)."""
        if self.undistort:
            raise ValueError("the float32 camera row has no lens model: use table_row64() for undistort=True")
        return np.array([self.K[0, 0], self.K[1, 1], self.K[0, 2], self.K[1, 2], self.cam_pitch_rad, self.height], dtype=np.float32)

    def table_row64(self) -> np.ndarray:
        """R3D_CAM_F64 row (include/ray3d_b200.h): the float64 intrinsics the reference divides by (camera.py:438-439),
        pp_cam (camera.py:253-259), cos/sin of the pitch from libm exactly as Rc2n holds them (camera.py:333-338),
        [pitch, height] for the embedding (trainer.py:297), the 5 distortion coefficients and K's own principal point."""
        d = self.dist_coeff if (self.undistort and self.dist_coeff is not None) else np.zeros(5)
        return np.array([self.K[0, 0], self.K[1, 1], self.pp_cam[0, 0], self.pp_cam[0, 1], math.cos(self.cam_pitch_rad),
                         math.sin(self.cam_pitch_rad), self.cam_pitch_rad, self.height, d[0], d[1], d[2], d[3], d[4],
                         1.0 if self.undistort else 0.0, self.K[0, 2], self.K[1, 2]], dtype=np.float64)

    def _undistort_host(self, u: float, v: float) -> np.ndarray:
        """One point through the same arithmetic as the device kernel (python floats are IEEE doubles, no FMA):
        used once per camera for the principal point (camera.py:253-256)."""
        fx, fy, cx, cy = (float(self.K[0, 0]), float(self.K[1, 1]), float(self.K[0, 2]), float(self.K[1, 2]))
        k1, k2, p1, p2, k3 = (float(c) for c in self.dist_coeff)
        ifx, ify = 1.0 / fx, 1.0 / fy
        x = (u - cx) * ifx
        y = (v - cy) * ify
        x0, y0 = x, y
        for _ in range(5):
            r2 = x * x + y * y
            icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2)
            dx = 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x)
            dy = p1 * (r2 + 2.0 * y * y) + 2.0 * p2 * x * y
            x = (x0 - dx) * icdist
            y = (y0 - dy) * icdist
        return np.array([fx * x + cx, fy * y + cy], dtype=np.float64)

    def undistort_point(self, points2d: ArrayLike, device: Optional[int] = None):
        """camera.py:412-421 on the device: (F, J, 2) pixels -> undistorted pixels, float64."""
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        x = _to_cuda_f64(points2d, dev)
        assert x.shape[-1] == 2 and self.dist_coeff is not None
        out = torch.empty_like(x)
        d5 = (_capi.C.c_double * 5)(*[float(c) for c in self.dist_coeff])
        with torch.cuda.device(dev):
            _capi.check(_capi.lib().r3d_undistort_points_f64(x.data_ptr(), out.data_ptr(), x.numel() // 2, float(self.K[0, 0]),
                                                             float(self.K[1, 1]), float(self.K[0, 2]), float(self.K[1, 2]), d5,
                                                             torch.cuda.current_stream(dev).cuda_stream))
        return _back(out, points2d)

    def _encode(self, uv: ArrayLike, c: float, s: float, device):
        dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        x = _to_cuda_f64(uv, dev)
        assert x.shape[-1] == 2
        if self.undistort:                                               # camera.py:435-436
            x = self.undistort_point(x, dev.index)
        out = torch.empty(x.shape[:-1] + (3,), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _capi.check(_capi.lib().r3d_ray_encode_f64(x.data_ptr(), out.data_ptr(), x.numel() // 2, float(self.K[0, 0]),
                                                       float(self.K[1, 1]), float(self.pp_cam[0, 0]), float(self.pp_cam[0, 1]),
                                                       c, s, torch.cuda.current_stream(dev).cuda_stream))
        return out

    def get_cam_ray_given_uv(self, uv: ArrayLike, device: Optional[int] = None):
        """camera.py:460-471: (F, J, 2) pixels -> (F, J, 3) rays in the pitch-normalised frame, float64."""
        out = self._encode(uv, math.cos(self.cam_pitch_rad), math.sin(self.cam_pitch_rad), device)
        return _back(out, uv)

    def encode_uv_with_intrinsic(self, uv: ArrayLike, device: Optional[int] = None):
        """camera.py:423-441: ((u - ppx)/fx, (v - ppy)/fy), float64 (the ray encode with zero pitch)."""
        out = self._encode(uv, 1.0, 0.0, device)[..., :2].contiguous()
        return _back(out, uv)
