"""Evaluation tail of the reference's eval loop on the device.

``evaluate`` stands in for the block at lib/train_val/trainer.py:355-395: predictions and targets are taken to
world coordinates with ``cam.normalized2world`` (lib/camera/camera.py:401-410) and scored with ``mpjpe``, the root
``mpjpe`` (MRPE), ``n_mpjpe``, ``p_mpjpe`` (per-frame Procrustes alignment, 3x3 SVD) and ``mean_velocity_error`` (lib/loss/loss.py) --
in float64, without the device->host->numpy->torch round trips and per-batch ``.item()`` syncs of the reference.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _capi


def evaluate(pred: torch.Tensor, target: torch.Tensor, Rn2w: Optional[np.ndarray] = None, Tn2w: Optional[np.ndarray] = None) -> Dict[str, float]:
    """pred/target (F, 1, J, 3) or (F, J, 3) float32 CUDA tensors in the normalised frame.  Returns the five means."""
    if not pred.is_cuda or not target.is_cuda:
        raise RuntimeError("ray3d_b200.metrics.evaluate runs on CUDA tensors only")
    assert pred.shape == target.shape
    J = pred.shape[-2]
    p = pred.reshape(-1, J, 3).contiguous().float()
    t = target.reshape(-1, J, 3).contiguous().float()
    F = p.shape[0]
    rt = None
    if Rn2w is not None:
        rt = torch.from_numpy(np.concatenate([np.asarray(Rn2w, np.float64).reshape(9), np.asarray(Tn2w, np.float64).reshape(3)])).to(p.device)
    sums = torch.empty(5, dtype=torch.float64, device=p.device)
    with torch.cuda.device(p.device):
        _capi.check(_capi.lib().r3d_eval_metrics(p.data_ptr(), t.data_ptr(), F, J, rt.data_ptr() if rt is not None else None,
                                                 sums.data_ptr(), torch.cuda.current_stream(p.device).cuda_stream))
    s = sums.cpu().numpy()
    return {"mpjpe": float(s[0] / (F * J)), "mrpe": float(s[1] / F), "n_mpjpe": float(s[2] / (F * J)),
            "p_mpjpe": float(s[4] / (F * J)),
            "mpjve": float(s[3] / ((F - 1) * J)) if F > 1 else float("nan")}
