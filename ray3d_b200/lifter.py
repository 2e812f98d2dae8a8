"""Torch-facing wrapper of one native plan: tensors in, tensors out, current CUDA stream honoured.

``Lifter`` is the fused entry point (ray encode + pose net + trajectory net + pos+trj in one
launch sequence) that stands in for the reference's eval step
(lib/train_val/trainer.py:323-353).  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

from typing import Mapping, Optional, Tuple

import numpy as np
import torch

from . import _capi
from .spec import NetSpec

DEFAULT_PRECISION = "bf16x3"


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor: ray3d_b200 has no CPU path (got device {t.device})")


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
                 precision: str = DEFAULT_PRECISION, device: Optional[int] = None):
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
        self.plan.finalize()
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device: ray3d_b200 computes the lifting path on the GPU only")
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.plan.upload(self.device)

    # -- helpers -----------------------------------------------------------------------------------
    def _outputs(self, batch: int, dev: torch.device, want_pos=True, want_trj=True, want_sum=True):
        J = self.spec.num_joints
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

    @staticmethod
    def _stream(dev) -> int:
        return torch.cuda.current_stream(dev).cuda_stream

    # -- device entry points -------------------------------------------------------------------------
    def forward_rays(self, x: torch.Tensor, param: Optional[torch.Tensor], want_pos=True, want_trj=True, want_sum=True):
        """x (B,RF,J,Cin) f32 cuda, param (B,extrinsic_dim) -> (pos (B,1,J,3), trj (B,1,1,3), pos+trj)."""
        self._check_x(x)
        _require_cuda(x, "x")
        x = x.contiguous().float()
        if self.spec.camera_embedding:
            if param is None:
                raise RuntimeError("param is required when the camera embedding is enabled")
            _require_cuda(param, "param")
            param = param.contiguous().float()
            assert param.shape == (x.shape[0], self.spec.extrinsic_dim)
        pos, trj, both = self._outputs(x.shape[0], x.device, want_pos, want_trj, want_sum)
        if x.shape[0] == 0:
            return pos, trj, both
        with torch.cuda.device(x.device):
            self.plan.forward_rays(_ptr(x), _ptr(param) if self.spec.camera_embedding else None, _ptr(pos), _ptr(trj),
                                   _ptr(both), x.shape[0], self._stream(x.device))
        return pos, trj, both

    def forward_uv(self, uv: torch.Tensor, cam: torch.Tensor, want_pos=True, want_trj=True, want_sum=True):
        """uv (B,RF,J,2) pixels f32 cuda, cam (B,6)=[fx,fy,cx,cy,pitch,height] f32 cuda."""
        _require_cuda(uv, "uv")
        _require_cuda(cam, "cam")
        assert uv.dim() == 4 and uv.shape[2] == self.spec.num_joints and uv.shape[3] == 2
        assert uv.shape[1] == self.spec.receptive_field and cam.shape == (uv.shape[0], 6)
        uv, cam = uv.contiguous().float(), cam.contiguous().float()
        pos, trj, both = self._outputs(uv.shape[0], uv.device, want_pos, want_trj, want_sum)
        if uv.shape[0] == 0:
            return pos, trj, both
        with torch.cuda.device(uv.device):
            self.plan.forward_uv(_ptr(uv), _ptr(cam), _ptr(pos), _ptr(trj), _ptr(both), uv.shape[0], self._stream(uv.device))
        return pos, trj, both

    # -- asynchronous device entry points: two batches in flight on the plan's two lanes ---------------------------
    def submit_uv(self, uv: torch.Tensor, cam: torch.Tensor, want_pos=True, want_trj=True, want_sum=True) -> "Pending":
        """forward_uv without blocking the caller's stream: the batch is lifted on one of the plan's two lanes
        (alternating), ordered after the work already enqueued on the current stream.  ``join(pending)`` makes the
        current stream wait for it and returns (pos, trj, pos+trj).  Keep two submissions in flight."""
        _require_cuda(uv, "uv")
        _require_cuda(cam, "cam")
        assert uv.dim() == 4 and uv.shape[2] == self.spec.num_joints and uv.shape[3] == 2
        assert uv.shape[1] == self.spec.receptive_field and cam.shape == (uv.shape[0], 6)
        uv, cam = uv.contiguous().float(), cam.contiguous().float()
        outs = self._outputs(uv.shape[0], uv.device, want_pos, want_trj, want_sum)
        with torch.cuda.device(uv.device):
            ticket = self.plan.submit_uv(_ptr(uv), _ptr(cam), _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), uv.shape[0],
                                         self._stream(uv.device))
        return Pending(ticket, outs, (uv, cam))

    def submit_rays(self, x: torch.Tensor, param: Optional[torch.Tensor], want_pos=True, want_trj=True, want_sum=True) -> "Pending":
        """Asynchronous forward_rays (see submit_uv)."""
        self._check_x(x)
        _require_cuda(x, "x")
        x = x.contiguous().float()
        if self.spec.camera_embedding:
            _require_cuda(param, "param")
            param = param.contiguous().float()
        outs = self._outputs(x.shape[0], x.device, want_pos, want_trj, want_sum)
        with torch.cuda.device(x.device):
            ticket = self.plan.submit_rays(_ptr(x), _ptr(param) if self.spec.camera_embedding else None, _ptr(outs[0]), _ptr(outs[1]),
                                           _ptr(outs[2]), x.shape[0], self._stream(x.device))
        return Pending(ticket, outs, (x, param))

    def join(self, pending: "Pending"):
        """The current stream waits (on the device) for the submission; returns its (pos, trj, pos+trj)."""
        dev = next(t for t in pending.outputs if t is not None).device
        with torch.cuda.device(dev):
            self.plan.join(pending.ticket, self._stream(dev))
        pending.inputs = None          # the lane has been ordered before anything that could reuse their memory
        return pending.outputs

    def forward_video(self, seq: torch.Tensor, param: Optional[torch.Tensor]):
        """seq (F+RF-1, J, Cin) f32 cuda (edge-padded video), param (extrinsic_dim,) -> F sliding-window outputs.
        Replaces eval_data_prepare + np.tile + forward (trainer.py:47-58, 323-337) without materialising windows."""
        _require_cuda(seq, "seq")
        assert seq.dim() == 3 and seq.shape[1] == self.spec.num_joints and seq.shape[2] == self.spec.in_features
        frames = seq.shape[0] - self.spec.receptive_field + 1
        if frames <= 0:
            raise RuntimeError("video shorter than one receptive field")
        seq = seq.contiguous().float()
        if self.spec.camera_embedding:
            _require_cuda(param, "param")
            param = param.contiguous().float().reshape(-1)
            assert param.numel() == self.spec.extrinsic_dim
        pos, trj, both = self._outputs(frames, seq.device)
        with torch.cuda.device(seq.device):
            self.plan.forward_video(_ptr(seq), _ptr(param) if self.spec.camera_embedding else None, _ptr(pos), _ptr(trj),
                                    _ptr(both), frames, self._stream(seq.device))
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
        self.plan.set_flip(swap(list(kps_left), list(kps_right)), swap(list(out_left), list(out_right)))

    def forward_rays_tta(self, x: torch.Tensor, param: Optional[torch.Tensor]):
        """forward_rays with the mirrored copy lifted in the same launch sequence and averaged (trainer.py:338-353)."""
        self._check_x(x)
        _require_cuda(x, "x")
        x = x.contiguous().float()
        if self.spec.camera_embedding:
            _require_cuda(param, "param")
            param = param.contiguous().float()
        pos, trj, both = self._outputs(x.shape[0], x.device)
        if x.shape[0] == 0:
            return pos, trj, both
        with torch.cuda.device(x.device):
            self.plan.forward_rays_tta(_ptr(x), _ptr(param) if self.spec.camera_embedding else None, _ptr(pos), _ptr(trj), _ptr(both),
                                       x.shape[0], self._stream(x.device))
        return pos, trj, both

    def forward_video_tta(self, seq: torch.Tensor, param: Optional[torch.Tensor]):
        """forward_video + flip augmentation: the whole evaluate_core inner step (trainer.py:299-353) in one call."""
        _require_cuda(seq, "seq")
        assert seq.dim() == 3 and seq.shape[1] == self.spec.num_joints and seq.shape[2] == self.spec.in_features
        frames = seq.shape[0] - self.spec.receptive_field + 1
        if frames <= 0:
            raise RuntimeError("video shorter than one receptive field")
        seq = seq.contiguous().float()
        if self.spec.camera_embedding:
            _require_cuda(param, "param")
            param = param.contiguous().float().reshape(-1)
        pos, trj, both = self._outputs(frames, seq.device)
        with torch.cuda.device(seq.device):
            self.plan.forward_video_tta(_ptr(seq), _ptr(param) if self.spec.camera_embedding else None, _ptr(pos), _ptr(trj),
                                        _ptr(both), frames, self._stream(seq.device))
        return pos, trj, both

    # -- host entry points (end-to-end: H2D + kernels + D2H inside the call) ---------------------------
    def forward_uv_host(self, uv: torch.Tensor, cam: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uv/cam are CPU tensors (pinned for full PCIe rate); returns pos+trj (B,1,J,3) on the CPU."""
        assert not uv.is_cuda and not cam.is_cuda and uv.dtype == torch.float32 and cam.dtype == torch.float32
        uv, cam = uv.contiguous(), cam.contiguous()
        B = uv.shape[0]
        if out is None:
            out = torch.empty((B, 1, self.spec.num_joints, 3), dtype=torch.float32, pin_memory=True)
        with torch.cuda.device(self.device):
            if self.has_pos and self.has_trj:
                self.plan.forward_uv_host(_ptr(uv), _ptr(cam), None, None, _ptr(out), B)
            elif self.has_pos:
                self.plan.forward_uv_host(_ptr(uv), _ptr(cam), _ptr(out), None, None, B)
            else:
                raise RuntimeError("forward_uv_host needs the pose net")
        return out

    def submit_uv_host(self, uv: torch.Tensor, cam: torch.Tensor, out: torch.Tensor) -> int:
        """Asynchronous forward_uv_host for streaming: uv/cam/out are PINNED, contiguous CPU tensors that stay untouched
        until ``wait(ticket)``; the H2D copy of this submission overlaps the kernels of the previous one."""
        for t in (uv, cam, out):
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or not t.is_pinned():
                raise ValueError("submit_uv_host needs pinned, contiguous float32 CPU tensors")
        B = uv.shape[0]
        assert out.shape == (B, 1, self.spec.num_joints, 3)
        with torch.cuda.device(self.device):
            if self.has_pos and self.has_trj:
                return self.plan.submit_uv_host(_ptr(uv), _ptr(cam), None, None, _ptr(out), B)
            if self.has_pos:
                return self.plan.submit_uv_host(_ptr(uv), _ptr(cam), _ptr(out), None, None, B)
        raise RuntimeError("submit_uv_host needs the pose net")

    def submit_rays_host(self, x: torch.Tensor, param: Optional[torch.Tensor], out: torch.Tensor) -> int:
        """Asynchronous forward_rays_host (see submit_uv_host)."""
        self._check_x(x)
        for t in (x, param, out):
            if t is not None and (t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or not t.is_pinned()):
                raise ValueError("submit_rays_host needs pinned, contiguous float32 CPU tensors")
        B = x.shape[0]
        assert out.shape == (B, 1, self.spec.num_joints, 3)
        with torch.cuda.device(self.device):
            if self.has_pos and self.has_trj:
                return self.plan.submit_rays_host(_ptr(x), _ptr(param), None, None, _ptr(out), B)
            return self.plan.submit_rays_host(_ptr(x), _ptr(param), _ptr(out), None, None, B)

    def wait(self, ticket: int) -> None:
        """Block until the submission named by ``ticket`` has its results in host memory."""
        with torch.cuda.device(self.device):
            self.plan.wait(ticket)

    def forward_rays_host(self, x: torch.Tensor, param: Optional[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
        assert not x.is_cuda and x.dtype == torch.float32
        self._check_x(x)
        x = x.contiguous()
        param = param.contiguous() if param is not None else None
        B = x.shape[0]
        if out is None:
            out = torch.empty((B, 1, self.spec.num_joints, 3), dtype=torch.float32, pin_memory=True)
        with torch.cuda.device(self.device):
            if self.has_pos and self.has_trj:
                self.plan.forward_rays_host(_ptr(x), _ptr(param), None, None, _ptr(out), B)
            else:
                self.plan.forward_rays_host(_ptr(x), _ptr(param), _ptr(out), None, None, B)
        return out
