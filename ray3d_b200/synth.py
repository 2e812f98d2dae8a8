"""Seeded synthetic weights and inputs (SURVEY.md section 8d "Synthetic inputs").

The reference's checkpoints are unreachable (Google Drive, README.md:75-76), so every parity
test and the benchmark run on *generated* weights.  Generation uses numpy's PCG64 stream, which is
stable across machines, so the same (spec, seed) gives bit-identical state_dicts in the build
container (where goldens are produced from the reference) and on the GPU box.

Distributions follow the reference's default torch initialisers (kaiming-uniform with a=sqrt(5)
== U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for conv/linear weight and bias) and the survey's
randomised BatchNorm statistics (so that BN folding bugs cannot hide behind identity BN).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np

from .spec import NetSpec, pos_state_entries, trj_state_entries

WEIGHT_SEED = 14   # mirrors the reference's default --random_seed (cfg/arguments.py:14)


def _fan_in(shape: Tuple[int, ...]) -> int:
    f = 1
    for d in shape[1:]:
        f *= d
    return f


def _fill(entries, rng: np.random.Generator) -> Dict[str, np.ndarray]:
    sd: Dict[str, np.ndarray] = {}
    last_fan_in = 1
    for name, shape, kind in entries:
        if kind == "weight":
            last_fan_in = _fan_in(shape)
            b = 1.0 / math.sqrt(last_fan_in)
            sd[name] = rng.uniform(-b, b, size=shape).astype(np.float32)
        elif kind == "bias":
            b = 1.0 / math.sqrt(last_fan_in)
            sd[name] = rng.uniform(-b, b, size=shape).astype(np.float32)
        elif kind == "bn_weight":
            sd[name] = rng.uniform(0.75, 1.25, size=shape).astype(np.float32)
        elif kind == "bn_bias":
            sd[name] = (0.1 * rng.standard_normal(size=shape)).astype(np.float32)
        elif kind == "bn_mean":
            sd[name] = (0.1 * rng.standard_normal(size=shape)).astype(np.float32)
        elif kind == "bn_var":
            sd[name] = rng.uniform(0.75, 1.25, size=shape).astype(np.float32)
        elif kind == "bn_count":
            sd[name] = np.asarray(0, dtype=np.int64)
        else:  # pragma: no cover
            raise AssertionError(kind)
    return sd


def make_state_dicts(spec: NetSpec, seed: int = WEIGHT_SEED) -> Tuple[Dict[str, np.ndarray], Dict[str, np.ndarray]]:
    """Return (pos_state_dict, trj_state_dict) as numpy arrays keyed by the reference's names."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pos = _fill(pos_state_entries(spec), rng)
    trj = _fill(trj_state_entries(spec), rng)
    return pos, trj


def state_digest(sd: Dict[str, np.ndarray]) -> float:
    """Order-sensitive float64 digest used to prove two boxes generated the same weights."""
    acc = 0.0
    for i, (k, v) in enumerate(sd.items()):
        a = np.asarray(v, dtype=np.float64).ravel()
        if a.size:
            acc += (i + 1) * float(a.sum()) + float(np.abs(a).sum()) * 1e-3 + float(a[:: max(1, a.size // 7)].sum())
    return acc


def make_cameras(batch: int, rng: np.random.Generator, res: int = 1000) -> np.ndarray:
    """Per-sequence camera table (B, 6) float32: fx, fy, cx, cy, pitch_rad, height_m."""
    cam = np.empty((batch, 6), dtype=np.float64)
    if res >= 2000:   # MPI-INF-3DHP-like pinhole cameras (mpii_3dhp_dataset.py:9-122)
        f = rng.uniform(1490.0, 1502.0, size=batch)
        cam[:, 0] = f
        cam[:, 1] = f
        cam[:, 2] = rng.uniform(975.0, 1053.0, size=batch)
        cam[:, 3] = rng.uniform(975.0, 1053.0, size=batch)
    else:             # Human3.6M-like (h36m_dataset.py:19-60)
        cam[:, 0] = rng.uniform(1095.0, 1200.0, size=batch)
        cam[:, 1] = rng.uniform(1095.0, 1200.0, size=batch)
        cam[:, 2] = rng.uniform(460.0, 570.0, size=batch)
        cam[:, 3] = rng.uniform(460.0, 570.0, size=batch)
    cam[:, 4] = rng.uniform(-0.86, 0.20, size=batch)   # pitch sweep of scripts/synthetic/test_aug.py:79
    cam[:, 5] = rng.uniform(1.0, 4.0, size=batch)      # camera height (m)
    return cam.astype(np.float32)


def make_uv(batch: int, frames: int, joints: int, rng: np.random.Generator, res: int = 1000,
            kind: str = "smooth") -> np.ndarray:
    """Pixel keypoints (B, T, J, 2) float32.  'smooth' = root random walk + fixed skeleton offsets
    + per-frame jitter (exercises the diff/diff_t cancellation); 'uniform' = iid U(0, res)."""
    if kind == "uniform":
        return rng.uniform(0.0, float(res), size=(batch, frames, joints, 2)).astype(np.float32)
    start = rng.uniform(0.3 * res, 0.7 * res, size=(batch, 1, 1, 2))
    walk = np.cumsum(2.0 * rng.standard_normal(size=(batch, frames, 1, 2)), axis=1)
    offs = 60.0 * rng.standard_normal(size=(batch, 1, joints, 2))
    offs[:, :, 0] = 0.0
    jitter = 1.5 * rng.standard_normal(size=(batch, frames, joints, 2))
    return (start + walk + offs + jitter).astype(np.float32)


def make_inputs(spec: NetSpec, batch: int, seed: int, res: int = 1000, kind: str = "smooth"):
    """Return (uv (B,T,J,2) f32, cam (B,6) f32) for one synthetic batch."""
    rng = np.random.Generator(np.random.PCG64(seed))
    cam = make_cameras(batch, rng, res)
    uv = make_uv(batch, spec.receptive_field, spec.num_joints, rng, res, kind)
    return uv, cam
