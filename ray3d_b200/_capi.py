"""ctypes binding of libray3d_b200.so (include/ray3d_b200.h).

There is deliberately no fallback: if the shared library is missing or no sm_100 device is
visible, every compute entry point raises.  Nothing in this package computes the lifting path on
the CPU or through PyTorch ops.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import Dict, Iterable, Mapping, Optional, Tuple

import numpy as np

from .spec import NetSpec

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libray3d_b200.so")
_lib = None
_lock = threading.Lock()

NET_POS, NET_TRJ = 1, 2
PRECISIONS = {"fp32": 0, "bf16x3": 1, "bf16": 2}

OK, ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_MISSING_WEIGHT, ERR_CUDA, ERR_STATE, ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5, -6


class R3DConfig(C.Structure):
    _fields_ = [("num_joints", C.c_int32), ("in_features", C.c_int32), ("n_widths", C.c_int32),
                ("widths", C.c_int32 * 8), ("channels", C.c_int32), ("latent", C.c_int32), ("stage", C.c_int32),
                ("extrinsic_dim", C.c_int32), ("embed_dim", C.c_int32), ("nets", C.c_int32),
                ("precision", C.c_int32)]


EXPORTS = {
    # name: (restype, argtypes)
    "r3d_abi_version": (C.c_int, []),
    "r3d_last_error": (C.c_char_p, []),
    "r3d_plan_create": (C.c_int, [C.POINTER(R3DConfig), C.POINTER(C.c_void_p)]),
    "r3d_plan_set_tensor": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.POINTER(C.c_int64), C.c_int]),
    "r3d_plan_finalize": (C.c_int, [C.c_void_p]),
    "r3d_plan_upload": (C.c_int, [C.c_void_p, C.c_int]),
    "r3d_plan_destroy": (None, [C.c_void_p]),
    "r3d_plan_packed_layer": (C.c_int, [C.c_void_p, C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                        C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]),
    "r3d_plan_describe": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]),
    "r3d_plan_weight_bytes": (C.c_int64, [C.c_void_p]),
    "r3d_plan_workspace_bytes": (C.c_int64, [C.c_void_p]),
    "r3d_plan_receptive_field": (C.c_int, [C.c_void_p]),
    "r3d_plan_kernel_launches": (C.c_int, [C.c_void_p]),
    "r3d_plan_graph_launches": (C.c_int64, [C.c_void_p]),
    "r3d_plan_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "r3d_plan_launch_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "r3d_plan_launch_name": (C.c_char_p, [C.c_void_p, C.c_int32]),
    "r3d_forward_rays": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int32, C.c_void_p]),
    "r3d_forward_uv": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int32, C.c_void_p]),
    "r3d_forward_rays_host": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int32]),
    "r3d_forward_uv_host": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int32]),
    "r3d_submit_rays_host": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int32, C.POINTER(C.c_uint64)]),
    "r3d_submit_uv_host": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int32, C.POINTER(C.c_uint64)]),
    "r3d_wait": (C.c_int, [C.c_void_p, C.c_uint64]),
    "r3d_submit_rays": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int32, C.c_void_p, C.POINTER(C.c_uint64)]),
    "r3d_submit_uv": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int32, C.c_void_p, C.POINTER(C.c_uint64)]),
    "r3d_join": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p]),
    "r3d_forward_video": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int32, C.c_void_p]),
    "r3d_plan_set_flip": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "r3d_forward_rays_tta": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int32, C.c_void_p]),
    "r3d_forward_video_tta": (C.c_int, [C.c_void_p] + [C.c_void_p] * 5 + [C.c_int32, C.c_void_p]),
    "r3d_ray_encode_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64] + [C.c_double] * 6 + [C.c_void_p]),
    "r3d_undistort_points_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64] + [C.c_double] * 4 + [C.POINTER(C.c_double), C.c_void_p]),
    "r3d_normalize_screen_f64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p]),
    "r3d_eval_metrics": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "r3d_selftest_gemm": (C.c_int, [C.c_int32] * 6 + [C.POINTER(C.c_double)] * 3),
    "r3d_forward": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_void_p] * 3 + [C.c_int32, C.c_void_p]),
    "r3d_submit": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_void_p] * 3 + [C.c_int32, C.c_void_p, C.POINTER(C.c_uint64)]),
    "r3d_forward_host": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_void_p] * 3 + [C.c_int32]),
    "r3d_submit_host": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_void_p] * 3 + [C.c_int32, C.POINTER(C.c_uint64)]),
    "r3d_plan_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32]),
    "r3d_forward_uv_cam64": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 3 + [C.c_int32, C.c_void_p]),
    "r3d_forward_video_uv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 3 + [C.c_int32, C.c_void_p]),
    "r3d_forward_video_uv_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 3 + [C.c_int32]),
    "r3d_submit_video_uv_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32] + [C.c_void_p] * 3 + [C.c_int32, C.POINTER(C.c_uint64)]),
}
# only in trace builds (R3D_BUILD_TRACE=1): scripts/tile_trace.py
OPTIONAL_EXPORTS = {
    "r3d_debug_tc_trace": (C.c_int, [C.c_int32, C.POINTER(C.c_int64), C.c_int32]),
    "r3d_debug_tail_stats": (C.c_int, [C.POINTER(C.c_uint64)]),      # R3D_BUILD_EXPERIMENTS=1 builds
}

SRC_RAYS, SRC_UV = 0, 1
CAM_PARAM, CAM_F32, CAM_F64 = 0, 1, 2
CAM64_STRIDE = 16
IN_UNDISTORT, IN_FLIP_TTA = 1, 2


class R3DInput(C.Structure):
    """Mirror of r3d_input (include/ray3d_b200.h)."""
    _fields_ = [("src", C.c_void_p), ("window_stride", C.c_int64), ("src_kind", C.c_int32), ("cam_kind", C.c_int32),
                ("cam", C.c_void_p), ("cam_stride", C.c_int64), ("flags", C.c_int32), ("reserved", C.c_int32)]


class R3DError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ray3d_b200 error {code}: {msg}")
        self.code = code


def library_path() -> str:
    return _LIB_PATH


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(_LIB_PATH):
                raise RuntimeError(
                    f"{_LIB_PATH} is missing: build it with `python -m ray3d_b200.build` (needs nvcc). "
                    "ray3d_b200 has no CPU/PyTorch fallback for the lifting path.")
            handle = C.CDLL(_LIB_PATH)
            for name, (res, args) in EXPORTS.items():
                fn = getattr(handle, name)     # AttributeError if the .so does not export it
                fn.restype = res
                fn.argtypes = args
            for name, (res, args) in OPTIONAL_EXPORTS.items():
                fn = getattr(handle, name, None)
                if fn is not None:
                    fn.restype = res
                    fn.argtypes = args
            if handle.r3d_abi_version() != 2:
                raise RuntimeError("libray3d_b200.so ABI version mismatch; rebuild it")
            _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise R3DError(rc, lib().r3d_last_error().decode("utf-8", "replace"))


def make_config(spec: NetSpec, nets: int, precision: str) -> R3DConfig:
    if precision not in PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
    cfg = R3DConfig()
    cfg.num_joints, cfg.in_features = spec.num_joints, spec.in_features
    cfg.n_widths = len(spec.filter_widths)
    if cfg.n_widths > 8:
        raise ValueError("at most 8 filter widths")
    for i, w in enumerate(spec.filter_widths):
        cfg.widths[i] = w
    cfg.channels, cfg.latent, cfg.stage = spec.channels, spec.latent, spec.stage
    cfg.extrinsic_dim, cfg.embed_dim = spec.extrinsic_dim, spec.embed_dim
    cfg.nets, cfg.precision = nets, PRECISIONS[precision]
    return cfg


class Plan:
    """Owns one native r3d_plan.  Host-side construction needs no GPU; upload()/forward do."""

    def __init__(self, spec: NetSpec, nets: int = NET_POS | NET_TRJ, precision: str = "fp32"):
        self.spec, self.nets, self.precision = spec, nets, precision
        self._h = C.c_void_p()
        cfg = make_config(spec, nets, precision)
        check(lib().r3d_plan_create(C.byref(cfg), C.byref(self._h)))
        self.device: Optional[int] = None

    def close(self) -> None:
        if self._h:
            lib().r3d_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # -- weights -----------------------------------------------------------------------------
    def set_tensor(self, net: int, name: str, arr: np.ndarray) -> None:
        a = np.ascontiguousarray(arr, dtype=np.float32)
        shape = (C.c_int64 * max(1, a.ndim))(*a.shape)
        check(lib().r3d_plan_set_tensor(self._h, net, name.encode(), a.ctypes.data_as(C.c_void_p), shape, a.ndim))

    def load_state(self, net: int, state: Mapping[str, object]) -> None:
        """state: name -> numpy array or torch tensor (any device); integer buffers are skipped."""
        for k, v in state.items():
            if k.endswith("num_batches_tracked"):
                continue
            if hasattr(v, "detach"):
                v = v.detach().to("cpu").float().numpy()
            self.set_tensor(net, k, np.asarray(v))

    def finalize(self) -> None:
        check(lib().r3d_plan_finalize(self._h))

    def upload(self, device: int = 0) -> None:
        check(lib().r3d_plan_upload(self._h, int(device)))
        self.device = int(device)

    def packed_layer(self, net: int, layer: str) -> Tuple[np.ndarray, np.ndarray]:
        n, k = C.c_int32(), C.c_int32()
        check(lib().r3d_plan_packed_layer(self._h, net, layer.encode(), C.byref(n), C.byref(k), None, 0, None, 0))
        w = np.empty((n.value, k.value), np.float32)
        b = np.empty((n.value,), np.float32)
        check(lib().r3d_plan_packed_layer(self._h, net, layer.encode(), C.byref(n), C.byref(k),
                                          w.ctypes.data_as(C.c_void_p), w.size, b.ctypes.data_as(C.c_void_p), b.size))
        return w, b

    def describe(self) -> dict:
        import json
        need = C.c_int64()
        check(lib().r3d_plan_describe(self._h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        check(lib().r3d_plan_describe(self._h, buf, need.value, C.byref(need)))
        return json.loads(buf.value.decode())

    @property
    def weight_bytes(self) -> int:
        return int(lib().r3d_plan_weight_bytes(self._h))

    @property
    def workspace_bytes(self) -> int:
        return int(lib().r3d_plan_workspace_bytes(self._h))

    @property
    def receptive_field(self) -> int:
        return int(lib().r3d_plan_receptive_field(self._h))

    @property
    def kernel_launches(self) -> int:
        return int(lib().r3d_plan_kernel_launches(self._h))

    @property
    def graph_launches(self) -> int:
        """forwards of this plan that replayed a captured CUDA graph (small batches)"""
        return int(lib().r3d_plan_graph_launches(self._h))

    def set_profiling(self, enable: bool) -> None:
        check(lib().r3d_plan_set_profiling(self._h, 1 if enable else 0))

    def launch_times(self):
        """[(launch name, mean ms)] over the forwards recorded since set_profiling(True), and the number of runs."""
        n, runs = C.c_int32(), C.c_int32()
        check(lib().r3d_plan_launch_times(self._h, None, 0, C.byref(n), C.byref(runs)))
        buf = (C.c_float * n.value)()
        check(lib().r3d_plan_launch_times(self._h, buf, n.value, C.byref(n), C.byref(runs)))
        names = [lib().r3d_plan_launch_name(self._h, i).decode() for i in range(n.value)]
        return list(zip(names, [float(v) for v in buf])), runs.value

    # -- forward (raw pointers; torch-facing wrappers live in lifter.py / model.py) ------------------
    def forward_rays(self, x_ptr, param_ptr, pos_ptr, trj_ptr, sum_ptr, batch: int, stream: int) -> None:
        check(lib().r3d_forward_rays(self._h, x_ptr, param_ptr, pos_ptr, trj_ptr, sum_ptr, batch, stream))

    def forward_uv(self, uv_ptr, cam_ptr, pos_ptr, trj_ptr, sum_ptr, batch: int, stream: int) -> None:
        check(lib().r3d_forward_uv(self._h, uv_ptr, cam_ptr, pos_ptr, trj_ptr, sum_ptr, batch, stream))

    def set_option(self, name: str, value: int) -> None:
        check(lib().r3d_plan_set_option(self._h, name.encode(), int(value)))

    @staticmethod
    def make_input(src_ptr, window_stride: int, src_kind: int, cam_ptr, cam_kind: int, cam_stride: int, flags: int = 0) -> R3DInput:
        return R3DInput(src_ptr, window_stride, src_kind, cam_kind, cam_ptr, cam_stride, flags, 0)

    def forward(self, inp: R3DInput, pos_ptr, trj_ptr, sum_ptr, n_windows: int, stream: int) -> None:
        check(lib().r3d_forward(self._h, C.byref(inp), pos_ptr, trj_ptr, sum_ptr, n_windows, stream))

    def forward_host(self, inp: R3DInput, pos_ptr, trj_ptr, sum_ptr, n_windows: int) -> None:
        check(lib().r3d_forward_host(self._h, C.byref(inp), pos_ptr, trj_ptr, sum_ptr, n_windows))

    def submit_host(self, inp: R3DInput, pos_ptr, trj_ptr, sum_ptr, n_windows: int) -> int:
        t = C.c_uint64()
        check(lib().r3d_submit_host(self._h, C.byref(inp), pos_ptr, trj_ptr, sum_ptr, n_windows, C.byref(t)))
        return int(t.value)

    def submit(self, inp: R3DInput, pos_ptr, trj_ptr, sum_ptr, n_windows: int, stream: int) -> int:
        t = C.c_uint64()
        check(lib().r3d_submit(self._h, C.byref(inp), pos_ptr, trj_ptr, sum_ptr, n_windows, stream, C.byref(t)))
        return int(t.value)

    def forward_uv_cam64(self, uv_ptr, cam_ptr, undistort: bool, pos_ptr, trj_ptr, sum_ptr, batch: int, stream: int) -> None:
        check(lib().r3d_forward_uv_cam64(self._h, uv_ptr, cam_ptr, int(bool(undistort)), pos_ptr, trj_ptr, sum_ptr, batch, stream))

    def forward_video_uv(self, uv_ptr, cam_ptr, flags: int, pos_ptr, trj_ptr, sum_ptr, frames_out: int, stream: int) -> None:
        check(lib().r3d_forward_video_uv(self._h, uv_ptr, cam_ptr, flags, pos_ptr, trj_ptr, sum_ptr, frames_out, stream))

    def forward_video_uv_host(self, uv_ptr, cam_ptr, flags: int, pos_ptr, trj_ptr, sum_ptr, frames_out: int) -> None:
        check(lib().r3d_forward_video_uv_host(self._h, uv_ptr, cam_ptr, flags, pos_ptr, trj_ptr, sum_ptr, frames_out))

    def submit_video_uv_host(self, uv_ptr, cam_ptr, flags: int, pos_ptr, trj_ptr, sum_ptr, frames_out: int) -> int:
        t = C.c_uint64()
        check(lib().r3d_submit_video_uv_host(self._h, uv_ptr, cam_ptr, flags, pos_ptr, trj_ptr, sum_ptr, frames_out, C.byref(t)))
        return int(t.value)

    def forward_video(self, seq_ptr, param_ptr, pos_ptr, trj_ptr, sum_ptr, frames_out: int, stream: int) -> None:
        check(lib().r3d_forward_video(self._h, seq_ptr, param_ptr, pos_ptr, trj_ptr, sum_ptr, frames_out, stream))

    def set_flip(self, in_perm, out_perm) -> None:
        n = self.spec.num_joints
        a = (C.c_int32 * n)(*[int(v) for v in in_perm])
        b = (C.c_int32 * n)(*[int(v) for v in out_perm])
        check(lib().r3d_plan_set_flip(self._h, a, b))

    def forward_rays_tta(self, x_ptr, param_ptr, pos_ptr, trj_ptr, sum_ptr, batch: int, stream: int) -> None:
        check(lib().r3d_forward_rays_tta(self._h, x_ptr, param_ptr, pos_ptr, trj_ptr, sum_ptr, batch, stream))

    def forward_video_tta(self, seq_ptr, param_ptr, pos_ptr, trj_ptr, sum_ptr, frames_out: int, stream: int) -> None:
        check(lib().r3d_forward_video_tta(self._h, seq_ptr, param_ptr, pos_ptr, trj_ptr, sum_ptr, frames_out, stream))

    def forward_rays_host(self, x_ptr, param_ptr, pos_ptr, trj_ptr, sum_ptr, batch: int) -> None:
        check(lib().r3d_forward_rays_host(self._h, x_ptr, param_ptr, pos_ptr, trj_ptr, sum_ptr, batch))

    def forward_uv_host(self, uv_ptr, cam_ptr, pos_ptr, trj_ptr, sum_ptr, batch: int) -> None:
        check(lib().r3d_forward_uv_host(self._h, uv_ptr, cam_ptr, pos_ptr, trj_ptr, sum_ptr, batch))

    def submit_uv_host(self, uv_ptr, cam_ptr, pos_ptr, trj_ptr, sum_ptr, batch: int) -> int:
        t = C.c_uint64(0)
        check(lib().r3d_submit_uv_host(self._h, uv_ptr, cam_ptr, pos_ptr, trj_ptr, sum_ptr, batch, C.byref(t)))
        return int(t.value)

    def submit_rays_host(self, x_ptr, param_ptr, pos_ptr, trj_ptr, sum_ptr, batch: int) -> int:
        t = C.c_uint64(0)
        check(lib().r3d_submit_rays_host(self._h, x_ptr, param_ptr, pos_ptr, trj_ptr, sum_ptr, batch, C.byref(t)))
        return int(t.value)

    def wait(self, ticket: int) -> None:
        check(lib().r3d_wait(self._h, ticket))

    def submit_uv(self, uv_ptr, cam_ptr, pos_ptr, trj_ptr, sum_ptr, batch: int, stream: int) -> int:
        t = C.c_uint64(0)
        check(lib().r3d_submit_uv(self._h, uv_ptr, cam_ptr, pos_ptr, trj_ptr, sum_ptr, batch, stream, C.byref(t)))
        return int(t.value)

    def submit_rays(self, x_ptr, param_ptr, pos_ptr, trj_ptr, sum_ptr, batch: int, stream: int) -> int:
        t = C.c_uint64(0)
        check(lib().r3d_submit_rays(self._h, x_ptr, param_ptr, pos_ptr, trj_ptr, sum_ptr, batch, stream, C.byref(t)))
        return int(t.value)

    def join(self, ticket: int, stream: int) -> None:
        check(lib().r3d_join(self._h, ticket, stream))


def selftest_gemm(m: int, n: int, k: int, nprob: int = 1, precision: str = "bf16x3", device: int = 0):
    err, t_tc, t_ff = C.c_double(), C.c_double(), C.c_double()
    check(lib().r3d_selftest_gemm(m, n, k, nprob, PRECISIONS[precision], device, C.byref(err), C.byref(t_tc), C.byref(t_ff)))
    return err.value, t_tc.value, t_ff.value
