"""Torch-facing wrapper of one native plan: tensors in, tensors out, current CUDA stream honoured.

``Lifter`` is the fused entry point (ray encode + pose net + trajectory net + pos+trj in one
launch sequence) that stands in for the reference's eval step
(lib/train_val/trainer.py:323-353).  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

from typing import Mapping, Optional

import numpy as np
import torch

from . import _capi
from .spec import NetSpec

DEFAULT_PRECISION = "bf16x3"


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class Pending:
    """An asynchronous device submission: its ticket, its output tensors (valid after ``Lifter.join``) and references
    that keep the input tensors alive while a lane may still read them."""
    __slots__ = ("ticket", "outputs", "inputs")

    def __init__(self, ticket: int, outputs, inputs):
        self.ticket, self.outputs, self.inputs = ticket, outputs, inputs


class Lifter:
    """Both networks behind one plan.

    state_pos / state_trj: state_dicts with the reference's key names (numpy arrays or tensors).
    Pass ``state_trj=None`` for a pose-only plan or ``state_pos=None`` for trajectory-only.
    """

    def __init__(self, spec: NetSpec, state_pos: Optional[Mapping[str, object]], state_trj: Optional[Mapping[str, object]],
                 precision: str = DEFAULT_PRECISION, device: Optional[int] = None, options: Optional[Mapping[str, int]] = None):
        nets = (_capi.NET_POS if state_pos is not None else 0) | (_capi.NET_TRJ if state_trj is not None else 0)
        if nets == 0:
            raise ValueError("need at least one state_dict")
        self.spec = spec
        self.has_pos, self.has_trj = state_pos is not None, state_trj is not None
        self.plan = _capi.Plan(spec, nets, precision)
        if state_pos is not None:
            self.plan.load_state(_capi.NET_POS, state_pos)
        if state_trj is not None:
            self.plan.load_state(_capi.NET_TRJ, state_trj)
        for k, v in (options or {}).items():      # build-time tuning (r3d_plan_set_option), e.g. {"tail_fusion": 0}
            self.plan.set_option(k, v)
        self.plan.finalize()
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: ray3d_b200 computes the lifting path on the GPU only")
        self.device = torch.cuda.current_device() if device is None else int(device)
        with torch.cuda.device(self.device):
            self.plan.upload(self.device)

    # -- helpers -----------------------------------------------------------------------------------
    def _dev(self) -> torch.device:
        return torch.device("cuda", self.device)

    def _on_device(self, t: torch.Tensor, what: str, dtype=torch.float32) -> torch.Tensor:
        """The C side dereferences raw pointers on the plan's device: anything else must fail here, loudly."""
        if not isinstance(t, torch.Tensor):
            raise RuntimeError(f"{what} must be a torch tensor, got {type(t).__name__}")
        if not t.is_cuda:
            raise RuntimeError(f"{what} must be a CUDA tensor: ray3d_b200 has no CPU path (got device {t.device})")
        if t.device.index != self.device:
            raise RuntimeError(f"{what} lives on {t.device} but this plan was uploaded to cuda:{self.device}")
        return t.contiguous().to(dtype)

    def _outputs(self, batch: int, want_pos=True, want_trj=True, want_sum=True):
        J, dev = self.spec.num_joints, self._dev()
        pos = torch.empty((batch, 1, J, 3), dtype=torch.float32, device=dev) if (want_pos and self.has_pos) else None
        trj = torch.empty((batch, 1, 1, 3), dtype=torch.float32, device=dev) if (want_trj and self.has_trj) else None
        both = torch.empty((batch, 1, J, 3), dtype=torch.float32, device=dev) if (want_sum and self.has_pos and self.has_trj) else None
        return pos, trj, both

    def _check_x(self, x: torch.Tensor) -> None:
        # mirrors the asserts at lib/model/rie.py:285-287 (AssertionError like the reference)
        assert len(x.shape) == 4
        assert x.shape[-2] == self.spec.num_joints
        assert x.shape[-1] == self.spec.in_features
        if x.shape[1] != self.spec.receptive_field:
            raise RuntimeError(f"sequence length {x.shape[1]} != receptive field {self.spec.receptive_field}: the reference "
                               "model only evaluates windows of exactly one receptive field (rie.py:362-386)")

    def _stream(self) -> int:
        return torch.cuda.current_stream(self._dev()).cuda_stream

    def _rays_input(self, x: torch.Tensor, param: Optional[torch.Tensor], host: bool = False, flags: int = 0):
        """(B,RF,J,Cin) windows + (B,extrinsic_dim) param rows -> descriptor (+ the tensors it points at)."""
        self._check_x(x)
        x = self._host(x, "x") if host else self._on_device(x, "x")
        if self.spec.camera_embedding:
            if param is None:
                raise RuntimeError("param is required when the camera embedding is enabled")
            param = self._host(param, "param") if host else self._on_device(param, "param")
            assert tuple(param.shape) == (x.shape[0], self.spec.extrinsic_dim)
        else:
            param = None
        T, J, Ci = self.spec.receptive_field, self.spec.num_joints, self.spec.in_features
        inp = self.plan.make_input(_ptr(x), T * J * Ci, _capi.SRC_RAYS, _ptr(param), _capi.CAM_PARAM, self.spec.extrinsic_dim, flags)
        return inp, (x, param)

    def _uv_input(self, uv: torch.Tensor, cam: torch.Tensor, host: bool = False, flags: int = 0):
        """(B,RF,J,2) pixel windows + camera rows: (B,6) float32 [fx,fy,cx,cy,pitch,height] or (B,16) float64
        (RayCamera.table_row64: the reference's float64 calibration, optional lens undistortion)."""
        assert uv.dim() == 4 and uv.shape[2] == self.spec.num_joints and uv.shape[3] == 2
        assert uv.shape[1] == self.spec.receptive_field
        uv = self._host(uv, "uv") if host else self._on_device(uv, "uv")
        f64 = cam.dtype == torch.float64
        cam = (self._host(cam, "cam", torch.float64 if f64 else torch.float32) if host
               else self._on_device(cam, "cam", torch.float64 if f64 else torch.float32))
        assert tuple(cam.shape) == (uv.shape[0], _capi.CAM64_STRIDE if f64 else 6)
        if f64 and bool((cam[:, 13] != 0).any()):
            flags |= _capi.IN_UNDISTORT
        T, J = self.spec.receptive_field, self.spec.num_joints
        inp = self.plan.make_input(_ptr(uv), T * J * 2, _capi.SRC_UV, _ptr(cam), _capi.CAM_F64 if f64 else _capi.CAM_F32,
                                   _capi.CAM64_STRIDE if f64 else 6, flags)
        return inp, (uv, cam)

    def _video_frames(self, seq: torch.Tensor, last: int) -> int:
        assert seq.dim() == 3 and seq.shape[1] == self.spec.num_joints and seq.shape[2] == last
        frames = seq.shape[0] - self.spec.receptive_field + 1
        if frames <= 0:
            raise RuntimeError("video shorter than one receptive field")
        return frames

    def _video_uv_input(self, uv_seq: torch.Tensor, cam_row, host: bool, tta: bool):
        """(F+RF-1,J,2) pixels of one edge-padded video + ONE float64 camera row (RayCamera or its table_row64())."""
        frames = self._video_frames(uv_seq, 2)
        row = cam_row.table_row64() if hasattr(cam_row, "table_row64") else cam_row
        row = torch.as_tensor(np.asarray(row, dtype=np.float64) if not isinstance(row, torch.Tensor) else row).reshape(-1).to(torch.float64)
        assert row.numel() == _capi.CAM64_STRIDE, "camera row must be the 16 float64 of RayCamera.table_row64()"
        flags = (_capi.IN_UNDISTORT if float(row[13]) != 0.0 else 0) | (_capi.IN_FLIP_TTA if tta else 0)
        if host:
            uv_seq, row = self._host(uv_seq, "uv_seq"), row.cpu().contiguous()
        else:
            uv_seq, row = self._on_device(uv_seq, "uv_seq"), row.to(self._dev()).contiguous()
        return frames, flags, uv_seq, row

    @staticmethod
    def _host(t: torch.Tensor, what: str, dtype=torch.float32) -> torch.Tensor:
        if not isinstance(t, torch.Tensor) or t.is_cuda or t.dtype != dtype:
            raise RuntimeError(f"{what} must be a CPU {dtype} tensor for the host-buffer calls")
        return t.contiguous()

    # -- device entry points -------------------------------------------------------------------------
    def forward_rays(self, x: torch.Tensor, param: Optional[torch.Tensor], want_pos=True, want_trj=True, want_sum=True):
        """x (B,RF,J,Cin) f32 cuda, param (B,extrinsic_dim) -> (pos (B,1,J,3), trj (B,1,1,3), pos+trj)."""
        inp, keep = self._rays_input(x, param)
        pos, trj, both = self._outputs(keep[0].shape[0], want_pos, want_trj, want_sum)
        if keep[0].shape[0]:
            with torch.cuda.device(self.device):
                self.plan.forward(inp, _ptr(pos), _ptr(trj), _ptr(both), keep[0].shape[0], self._stream())
        return pos, trj, both

    def forward_uv(self, uv: torch.Tensor, cam: torch.Tensor, want_pos=True, want_trj=True, want_sum=True):
        """uv (B,RF,J,2) pixels f32 cuda; cam (B,6) f32 = [fx,fy,cx,cy,pitch,height], or (B,16) f64 rows of
        RayCamera.table_row64() (float64 calibration, lens undistortion inside the input stage)."""
        inp, keep = self._uv_input(uv, cam)
        B = keep[0].shape[0]
        pos, trj, both = self._outputs(B, want_pos, want_trj, want_sum)
        if B:
            with torch.cuda.device(self.device):
                self.plan.forward(inp, _ptr(pos), _ptr(trj), _ptr(both), B, self._stream())
        return pos, trj, both

    # -- asynchronous device entry points: two batches in flight on the plan's two lanes ---------------------------
    def submit_uv(self, uv: torch.Tensor, cam: torch.Tensor, want_pos=True, want_trj=True, want_sum=True) -> "Pending":
        """forward_uv without blocking the caller's stream: the batch is lifted on one of the plan's two lanes
        (alternating), ordered after the work already enqueued on the current stream.  ``join(pending)`` makes the
        current stream wait for it and returns (pos, trj, pos+trj).  Keep two submissions in flight."""
        inp, keep = self._uv_input(uv, cam)
        outs = self._outputs(keep[0].shape[0], want_pos, want_trj, want_sum)
        with torch.cuda.device(self.device):
            ticket = self.plan.submit(inp, _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), keep[0].shape[0], self._stream())
        return Pending(ticket, outs, keep)

    def submit_rays(self, x: torch.Tensor, param: Optional[torch.Tensor], want_pos=True, want_trj=True, want_sum=True) -> "Pending":
        """Asynchronous forward_rays (see submit_uv)."""
        inp, keep = self._rays_input(x, param)
        outs = self._outputs(keep[0].shape[0], want_pos, want_trj, want_sum)
        with torch.cuda.device(self.device):
            ticket = self.plan.submit(inp, _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), keep[0].shape[0], self._stream())
        return Pending(ticket, outs, keep)

    def join(self, pending: "Pending"):
        """The current stream waits (on the device) for the submission; returns its (pos, trj, pos+trj)."""
        with torch.cuda.device(self.device):
            self.plan.join(pending.ticket, self._stream())
        pending.inputs = None          # the lane has been ordered before anything that could reuse their memory
        return pending.outputs

    def _video(self, seq: torch.Tensor, param: Optional[torch.Tensor], tta: bool):
        frames = self._video_frames(seq, self.spec.in_features)
        seq = self._on_device(seq, "seq")
        if self.spec.camera_embedding:
            if param is None:
                raise RuntimeError("param is required when the camera embedding is enabled")
            param = self._on_device(param, "param").reshape(-1)
            assert param.numel() == self.spec.extrinsic_dim
        else:
            param = None
        inp = self.plan.make_input(_ptr(seq), self.spec.num_joints * self.spec.in_features, _capi.SRC_RAYS, _ptr(param), _capi.CAM_PARAM, 0,
                                   _capi.IN_FLIP_TTA if tta else 0)
        pos, trj, both = self._outputs(frames)
        with torch.cuda.device(self.device):
            self.plan.forward(inp, _ptr(pos), _ptr(trj), _ptr(both), frames, self._stream())
        return pos, trj, both

    def forward_video(self, seq: torch.Tensor, param: Optional[torch.Tensor]):
        """seq (F+RF-1, J, Cin) f32 cuda (edge-padded video), param (extrinsic_dim,) -> F sliding-window outputs.
        Replaces eval_data_prepare + np.tile + forward (trainer.py:47-58, 323-337) without materialising windows."""
        return self._video(seq, param, False)

    def forward_video_uv(self, uv_seq: torch.Tensor, camera, tta: bool = False):
        """uv_seq (F+RF-1, J, 2) pixel keypoints of one edge-padded video (cuda), camera = RayCamera (or its 16-double
        table_row64()): every frame is ray-encoded once on the device, windows are indexed in place; with ``tta`` the
        flip augmentation runs in the same launch sequence.  The whole evaluate_core step from pixels
        (trainer.py:297-353) in one call."""
        frames, flags, uv_seq, row = self._video_uv_input(uv_seq, camera, False, tta)
        pos, trj, both = self._outputs(frames)
        with torch.cuda.device(self.device):
            self.plan.forward_video_uv(_ptr(uv_seq), _ptr(row), flags, _ptr(pos), _ptr(trj), _ptr(both), frames, self._stream())
        return pos, trj, both

    # -- flip test-time augmentation (Trainer.evaluate_core with flip_test=True) ------------------------------
    def set_flip(self, kps_left, kps_right, out_left=None, out_right=None) -> None:
        """Declare the left/right joint pairs.  Inputs are mirrored with (kps_left, kps_right) as in
        trainer.py:302; predictions are un-mirrored with (out_left, out_right), which default to the same lists
        because evaluate_core itself indexes the outputs with kps_left/kps_right (trainer.py:341-342)."""
        J = self.spec.num_joints
        out_left = kps_left if out_left is None else out_left
        out_right = kps_right if out_right is None else out_right

        def swap(left, right):
            perm = list(range(J))
            for l, r in zip(left, right):
                perm[l], perm[r] = r, l
            return perm
        with torch.cuda.device(self.device):
            self.plan.set_flip(swap(list(kps_left), list(kps_right)), swap(list(out_left), list(out_right)))

    def forward_rays_tta(self, x: torch.Tensor, param: Optional[torch.Tensor]):
        """forward_rays with the mirrored copy lifted in the same launch sequence and averaged (trainer.py:338-353)."""
        inp, keep = self._rays_input(x, param, flags=_capi.IN_FLIP_TTA)
        pos, trj, both = self._outputs(keep[0].shape[0])
        if keep[0].shape[0]:
            with torch.cuda.device(self.device):
                self.plan.forward(inp, _ptr(pos), _ptr(trj), _ptr(both), keep[0].shape[0], self._stream())
        return pos, trj, both

    def forward_video_tta(self, seq: torch.Tensor, param: Optional[torch.Tensor]):
        """forward_video + flip augmentation: the whole evaluate_core inner step (trainer.py:299-353) in one call."""
        return self._video(seq, param, True)

    # -- host entry points (end-to-end: H2D + kernels + D2H inside the call) ---------------------------
    def _host_out(self, B: int, out: Optional[torch.Tensor]):
        if out is None:
            out = torch.empty((B, 1, self.spec.num_joints, 3), dtype=torch.float32, pin_memory=True)
        assert tuple(out.shape) == (B, 1, self.spec.num_joints, 3) and out.dtype == torch.float32 and not out.is_cuda and out.is_contiguous()
        # (pos, trj, sum) pointers: pos+trj when both nets are loaded, the pose alone otherwise
        if self.has_pos and self.has_trj:
            return out, (None, None, _ptr(out))
        if self.has_pos:
            return out, (_ptr(out), None, None)
        raise RuntimeError("the host-buffer calls need the pose net")

    @staticmethod
    def _need_pinned(*ts) -> None:
        for t in ts:
            if t is not None and not t.is_pinned():
                raise ValueError("asynchronous host submissions need pinned (page-locked) CPU tensors")

    def forward_uv_host(self, uv: torch.Tensor, cam: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uv/cam are CPU tensors (pinned for full PCIe rate); returns pos+trj (B,1,J,3) on the CPU."""
        inp, keep = self._uv_input(uv, cam, host=True)
        out, ptrs = self._host_out(keep[0].shape[0], out)
        with torch.cuda.device(self.device):
            self.plan.forward_host(inp, *ptrs, keep[0].shape[0])
        return out

    def submit_uv_host(self, uv: torch.Tensor, cam: torch.Tensor, out: torch.Tensor) -> int:
        """Asynchronous forward_uv_host for streaming: uv/cam/out are PINNED, contiguous CPU tensors that stay untouched
        until ``wait(ticket)``; the H2D copy of this submission overlaps the kernels of the previous one."""
        if not (uv.is_contiguous() and cam.is_contiguous()):
            raise ValueError("submit_uv_host needs pinned, contiguous float32 CPU tensors")
        inp, keep = self._uv_input(uv, cam, host=True)
        self._need_pinned(keep[0], keep[1], out)
        out, ptrs = self._host_out(keep[0].shape[0], out)
        with torch.cuda.device(self.device):
            return self.plan.submit_host(inp, *ptrs, keep[0].shape[0])

    def submit_rays_host(self, x: torch.Tensor, param: Optional[torch.Tensor], out: torch.Tensor) -> int:
        """Asynchronous forward_rays_host (see submit_uv_host)."""
        if not x.is_contiguous() or (param is not None and not param.is_contiguous()):
            raise ValueError("submit_rays_host needs pinned, contiguous float32 CPU tensors")
        inp, keep = self._rays_input(x, param, host=True)
        self._need_pinned(keep[0], keep[1], out)
        out, ptrs = self._host_out(keep[0].shape[0], out)
        with torch.cuda.device(self.device):
            return self.plan.submit_host(inp, *ptrs, keep[0].shape[0])

    def wait(self, ticket: int) -> None:
        """Block until the submission named by ``ticket`` has its results in host memory."""
        with torch.cuda.device(self.device):
            self.plan.wait(ticket)

    def forward_rays_host(self, x: torch.Tensor, param: Optional[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        inp, keep = self._rays_input(x, param, host=True)
        out, ptrs = self._host_out(keep[0].shape[0], out)
        with torch.cuda.device(self.device):
            self.plan.forward_host(inp, *ptrs, keep[0].shape[0])
        return out

    def forward_video_uv_host(self, uv_seq: torch.Tensor, camera, out: Optional[torch.Tensor] = None, tta: bool = False) -> torch.Tensor:
        """forward_video_uv on HOST buffers: the padded video (136 bytes per frame) crosses PCIe once, the F results come
        back; stands in for trainer.py:297-356 (eval_data_prepare, np.tile, .cuda(), two forwards, flip, .cpu())."""
        frames, flags, uv_seq, row = self._video_uv_input(uv_seq, camera, True, tta)
        out, ptrs = self._host_out(frames, out)
        with torch.cuda.device(self.device):
            self.plan.forward_video_uv_host(_ptr(uv_seq), _ptr(row), flags, *ptrs, frames)
        return out

    def submit_video_uv_host(self, uv_seq: torch.Tensor, cam_row64: torch.Tensor, out: torch.Tensor, tta: bool = False) -> int:
        """Streaming form of forward_video_uv_host: pinned uv_seq / cam_row64 (16 float64) / out, untouched until wait()."""
        frames, flags, uv_seq, row = self._video_uv_input(uv_seq, cam_row64, True, tta)
        if row.data_ptr() != cam_row64.data_ptr():
            raise ValueError("submit_video_uv_host needs the camera row as a contiguous float64 CPU tensor (it is read asynchronously)")
        self._need_pinned(uv_seq, cam_row64, out)
        out, ptrs = self._host_out(frames, out)
        with torch.cuda.device(self.device):
            return self.plan.submit_video_uv_host(_ptr(uv_seq), _ptr(row), flags, *ptrs, frames)
