"""ray3d_b200 -- B200-native (sm_100a) implementation of Ray3D's 2D->3D lifting forward pass.

Public surface:
  Model, RIEModel, RIETrajectoryModel ... drop-in for lib/model of the reference
  Lifter, fused_lifter .................. fused ray-encode + pose + trajectory entry point
  RayCamera, normalize_screen_coordinates  camera-side encode (lib/camera/camera.py)
  NetSpec ............................... static description of the networks
"""
from .spec import NetSpec  # noqa: F401
from .lifter import Lifter, DEFAULT_PRECISION  # noqa: F401
from .model import Model, RIEModel, RIETrajectoryModel, TemporalBlock, FCBlock, Linear, Embedding, fused_lifter  # noqa: F401
from .camera import RayCamera, normalize_screen_coordinates  # noqa: F401
from . import metrics  # noqa: F401

__version__ = "0.1.0"
