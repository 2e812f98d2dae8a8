"""Drop-in nn.Modules for the reference's lifting networks.

``RIEModel`` / ``RIETrajectoryModel`` keep the reference's constructor signature
(lib/model/rie.py:178-181, 443-446), ``forward(x, param)`` signature (rie.py:284, 518), method
surface (``receptive_field``, ``set_bn_momentum``, ``set_training_status``, ``set_augment``) and,
crucially, the exact parameter/buffer names and registration order, so checkpoints,
``load_state_dict(strict=True)`` (trainer.py:161) and the freeze-by-index loop (main.py:164-168)
keep working.  The modules are *parameter containers*: in eval mode ``forward`` hands the
weights (BatchNorm folded, packed once per weight version and device) to the native sm_100a
library and launches the fused kernels on the caller's CUDA stream.  There is no PyTorch
compute path: CPU tensors or training-mode calls raise.

``Model`` mirrors the factory at lib/model/__init__.py:5-62 (same ``model_config`` keys).
"""
from __future__ import annotations

import os
import threading
import weakref
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _capi
from .lifter import DEFAULT_PRECISION, Lifter
from .spec import GROUPS, NetSpec, FC_WIDTH, EMBED_MID


def _precision() -> str:
    return os.environ.get("RAY3D_B200_PRECISION", DEFAULT_PRECISION)


# --------------------------------------------------------------------------------------------------
# parameter containers (names == reference names)
# --------------------------------------------------------------------------------------------------
class TemporalBlock(nn.Module):
    """Weights of the strided temporal tree (reference: lib/model/rie.py:7-63, Optimize1f=True)."""

    def __init__(self, num_joints_in, in_features, num_joints_out, filter_widths, causal=False, dropout=0.2,
                 channels=1024, latten_features=256, dense=False, is_train=True, Optimize1f=True):
        super().__init__()
        if causal or dense or not Optimize1f:
            raise NotImplementedError("only the strided, non-causal, non-dense variant runs in the reference for "
                                      "T == receptive field (SURVEY 8a); that is the one implemented")
        self.is_train, self.augment = is_train, False
        self.filter_widths = list(filter_widths)
        self.pad = [filter_widths[0] // 2]
        dil = filter_widths[0]
        for w in filter_widths[1:]:
            self.pad.append((w - 1) * dil // 2)
            dil *= w
        self.drop = nn.Dropout(dropout)
        self.relu = nn.LeakyReLU(0.2, inplace=True)
        self.expand_bn = nn.BatchNorm1d(channels, momentum=0.1)
        self.shrink = nn.Conv1d(channels, latten_features, 1)
        self.expand_conv = nn.Conv1d(num_joints_in * in_features, channels, filter_widths[0], stride=filter_widths[0], bias=False)
        convs, bns = [], []
        for w in filter_widths[1:]:
            convs += [nn.Conv1d(channels, channels, w, stride=w, bias=False), nn.Conv1d(channels, channels, 1, bias=False)]
            bns += [nn.BatchNorm1d(channels, momentum=0.1), nn.BatchNorm1d(channels, momentum=0.1)]
        self.layers_conv = nn.ModuleList(convs)
        self.layers_bn = nn.ModuleList(bns)

    def set_bn_momentum(self, momentum):
        for bn in [self.expand_bn, *self.layers_bn]:
            bn.momentum = momentum

    def set_training_status(self, is_train):
        self.is_train = is_train

    def set_augment(self, augment):
        self.augment = augment

    def receptive_field(self):
        return 1 + 2 * sum(self.pad)

    def forward(self, x):
        raise RuntimeError("TemporalBlock is evaluated inside the fused native kernels; call the owning model")


class Linear(nn.Module):
    """Residual pair of 1024-wide linears (reference: rie.py:108-120)."""

    def __init__(self, linear_size, p_dropout=0.25):
        super().__init__()
        self.l_size = linear_size
        self.relu = nn.LeakyReLU(0.2, inplace=True)
        self.dropout = nn.Dropout(p_dropout)
        self.w1 = nn.Linear(linear_size, linear_size)
        self.batch_norm1 = nn.BatchNorm1d(linear_size)
        self.w2 = nn.Linear(linear_size, linear_size)
        self.batch_norm2 = nn.BatchNorm1d(linear_size)


class FCBlock(nn.Module):
    """fc_1 -> BN -> act -> n residual blocks -> fc_2 (reference: rie.py:138-157)."""

    def __init__(self, channel_in, channel_out, linear_size, block_num):
        super().__init__()
        self.linear_size, self.block_num, self.channel_in = linear_size, block_num, channel_in
        self.fc_1 = nn.Linear(channel_in, linear_size)
        self.bn_1 = nn.BatchNorm1d(linear_size)
        self.fc_2 = nn.Linear(linear_size, channel_out)
        self.layers = nn.ModuleList([Linear(linear_size, 0.25) for _ in range(block_num)])
        self.relu = nn.LeakyReLU(0.2, inplace=True)
        self.dropout = nn.Dropout(0.25)


class Embedding(nn.Module):
    """Camera-extrinsic embedding (reference: lib/model/embedding.py:4-13)."""

    def __init__(self, in_channels, out_channels, mid_channels=EMBED_MID, p_dropout=0.25):
        super().__init__()
        self.relu = nn.LeakyReLU(inplace=True)
        self.dropout = nn.Dropout(p_dropout)
        self.w1 = nn.Linear(in_channels, mid_channels)
        self.b1 = nn.BatchNorm1d(mid_channels)
        self.w2 = nn.Linear(mid_channels, out_channels)
        self.b2 = nn.BatchNorm1d(out_channels)


# --------------------------------------------------------------------------------------------------
# plan cache shared between a module and its nn.DataParallel replicas
# --------------------------------------------------------------------------------------------------
class _PlanCache:
    def __init__(self, owner: nn.Module):
        self.owner = weakref.ref(owner)
        self.lock = threading.Lock()
        self.plans: Dict[Tuple, Tuple[object, Lifter]] = {}
        self.tensors = None                     # flat list of the owner's parameters + buffers (rebuilt on invalidate)
        self.sibling = None                     # weakref to the other network of the same Model (pose <-> trajectory)
        self.spec_out: Dict[int, tuple] = {}    # id(x) -> speculative result for the sibling's next call

    def invalidate(self):
        with self.lock:
            self.plans.clear()
            self.tensors = None
            self.spec_out.clear()


class _NativeNet(nn.Module):
    """Shared machinery: weight-version tracking and plan lookup.

    The packed native copy of the weights is rebuilt when ``load_state_dict`` runs (also through a wrapper such as
    nn.DataParallel or the reference's ``load_weight``, lib/utils/utils.py:208-218: a post-hook fires inside the
    recursion), when the module is moved/cast, put in training mode, or when any parameter/buffer was replaced or
    modified in place through autograd-visible ops (per-tensor ``(data_ptr, _version)`` fingerprint).  Writes through
    ``tensor.data`` bypass the version counter: call ``refresh_plan()`` after those."""

    _net_kind = "pos"

    def _init_native(self, spec: NetSpec):
        self._spec = spec
        object.__setattr__(self, "_cache", _PlanCache(self))
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._cache.invalidate())

    # invalidation hooks ---------------------------------------------------------------------------
    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self._cache.invalidate()
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._cache.invalidate()
        return out

    def train(self, mode: bool = True):
        if mode:
            self._cache.invalidate()
        return super().train(mode)

    def refresh_plan(self):
        """Force the packed native weights to be rebuilt on the next forward (needed after ``p.data`` edits)."""
        self._cache.invalidate()

    @staticmethod
    def _fingerprint(mod: "_NativeNet"):
        cache = mod._cache
        ts = cache.tensors
        if ts is None:
            ts = cache.tensors = list(mod.parameters()) + list(mod.buffers())
        return hash(tuple([(t.data_ptr(), t._version) for t in ts]))

    def _owner(self) -> "_NativeNet":
        return self._cache.owner() or self          # replicas read the master's weights

    def _lifter(self, device: torch.device, joint_with: Optional["_NativeNet"] = None) -> Lifter:
        """The plan of this network alone, or -- ``joint_with`` = the sibling network -- ONE plan that evaluates both
        (grouped launches over the 5 joint groups + the trajectory net, one input stage)."""
        cache: _PlanCache = self._cache
        owner = self._owner()
        check = os.environ.get("RAY3D_B200_CHECK_WEIGHTS", "1") != "0"
        fp = self._fingerprint(owner) if check else 0
        if joint_with is not None:
            fp = (fp, self._fingerprint(joint_with) if check else 0)
        key = (device.index if device.index is not None else torch.cuda.current_device(), _precision(), joint_with is not None)
        with cache.lock:
            hit = cache.plans.get(key)
            if hit is not None and hit[0] == fp:
                return hit[1]
            sd = owner.state_dict()
            if joint_with is not None:
                sd2 = joint_with.state_dict()
                pos_sd, trj_sd = (sd, sd2) if self._net_kind == "pos" else (sd2, sd)
                spec = owner._spec if self._net_kind == "pos" else joint_with._spec     # the pose net's spec carries `stage`
                lf = Lifter(spec, pos_sd, trj_sd, precision=key[1], device=key[0])
            elif self._net_kind == "pos":
                lf = Lifter(self._spec, sd, None, precision=key[1], device=key[0])
            else:
                lf = Lifter(self._spec, None, sd, precision=key[1], device=key[0])
            cache.plans[key] = (fp, lf)
            return lf

    def _sibling(self) -> Optional["_NativeNet"]:
        ref = self._owner()._cache.sibling
        sib = ref() if ref is not None else None
        if sib is None or sib.training or os.environ.get("RAY3D_B200_JOINT", "1") == "0":
            return None
        return sib

    def _native_forward(self, x: torch.Tensor, param: Optional[torch.Tensor]) -> torch.Tensor:
        assert len(x.shape) == 4                                   # rie.py:285-287
        assert x.shape[-2] == self.num_joints_in
        assert x.shape[-1] == self.in_features
        if self.training:
            raise RuntimeError("ray3d_b200 implements the eval-mode forward only (BatchNorm running statistics, "
                               "no dropout); call .eval() -- training is out of scope of this drop-in")
        if not x.is_cuda:
            raise RuntimeError("ray3d_b200 has no CPU path: move the input to a CUDA device")
        sib = self._sibling()
        if sib is None:
            pos, trj, _ = self._lifter(x.device).forward_rays(x, param, want_sum=False)
            return pos if self._net_kind == "pos" else trj
        # The reference's callers evaluate BOTH networks on the same tensors back to back (trainer.py:337-348:
        # pos_model(x, param) ... trj_model(x, param)).  The first of the two calls runs one joint plan and parks the
        # sibling's output; the sibling's call on the very same tensor objects (same storage, same version counters,
        # same stream, unchanged weights) takes it instead of launching the whole tree again.  Anything else misses
        # and computes normally; a parked result is handed out once (callers mutate outputs in place).
        owner, stream = self._owner(), torch.cuda.current_stream(x.device).cuda_stream
        pv = None if param is None else (id(param), param.data_ptr(), param._version)
        ident = (id(x), x.data_ptr(), x._version, tuple(x.shape), pv, stream, _precision())
        with owner._cache.lock:
            parked = owner._cache.spec_out.pop(id(x), None)
        if parked is not None and parked[0] == ident and parked[1]() is x and parked[2] == self._fingerprint(owner):
            return parked[3]
        lf = self._lifter(x.device, joint_with=sib._owner())
        pos, trj, _ = lf.forward_rays(x, param, want_sum=False)
        mine, theirs = (pos, trj) if self._net_kind == "pos" else (trj, pos)
        sc = sib._owner()._cache
        with sc.lock:
            if len(sc.spec_out) >= 4:
                sc.spec_out.clear()
            sc.spec_out[id(x)] = (ident, weakref.ref(x), self._fingerprint(sib._owner()), theirs)
        return mine


def link_models(pos_model: nn.Module, trj_model: nn.Module) -> None:
    """Declare the two networks of one Model siblings: their eval forwards then share one native plan (one input
    stage, grouped launches), see _NativeNet._native_forward."""
    p, t = getattr(pos_model, "module", pos_model), getattr(trj_model, "module", trj_model)
    if p._spec.replace_stage(1) != t._spec.replace_stage(1):
        return                                      # different architectures: nothing to share
    p._cache.sibling, t._cache.sibling = weakref.ref(t), weakref.ref(p)


def _spec_from_ctor(num_joints_in, in_features, filter_widths, latten_features, channels, stage, extrinsic_dim, embedd_dim):
    embed = extrinsic_dim > 0 and embedd_dim > 0
    return NetSpec(num_joints=num_joints_in, in_features=in_features, filter_widths=tuple(filter_widths), channels=channels,
                   latent=latten_features, stage=stage, extrinsic_dim=extrinsic_dim if embed else 0,
                   embed_dim=embedd_dim if embed else 0)


class RIEModel(_NativeNet):
    """Pose network.  Same ctor/forward as lib/model/rie.py:172-434."""

    _net_kind = "pos"

    def __init__(self, num_joints_in, in_features, num_joints_out, filter_widths, causal=False, dropout=0.2,
                 latten_features=256, channels=1024, dense=False, is_train=True, Optimize1f=True, stage=1,
                 extrinsic_dim=12, embedd_dim=64):
        super().__init__()
        spec = _spec_from_ctor(num_joints_in, in_features, filter_widths, latten_features, channels, stage, extrinsic_dim, embedd_dim)
        self.augment, self.is_train = False, is_train
        self.num_joints_in, self.num_joints_out, self.in_features = num_joints_in, num_joints_out, in_features
        self.latten_features, self.stage = latten_features, stage
        for g in GROUPS:
            setattr(self, f"LocalLayer_{g}", TemporalBlock(3 * len(spec.group_joints(g)), in_features, num_joints_out, filter_widths,
                                                           causal, dropout, channels, latten_features, dense, is_train, Optimize1f))
        self.pad = (self.receptive_field() - 1) // 2
        self.GlobalInfo = FCBlock(num_joints_in * in_features, latten_features, FC_WIDTH, 2)
        if stage != 1:
            self.FuseBlocks = nn.ModuleList([FCBlock(latten_features * 4, latten_features, FC_WIDTH, 1) for _ in range(5)])
        self.camera_embedding = extrinsic_dim > 0 and embedd_dim > 0
        self.extrinsic_dim, self.embedd_dim = extrinsic_dim, embedd_dim
        if self.camera_embedding:
            self.embedder = Embedding(in_channels=extrinsic_dim, out_channels=embedd_dim)
        self.out_features_dim = latten_features * (2 if stage == 1 else 3) + embedd_dim
        for g in GROUPS:
            setattr(self, f"Integration_{g}", FCBlock(self.out_features_dim, spec.group_out_dim(g), FC_WIDTH, 1))
        self._init_native(spec)

    def _blocks(self):
        return [getattr(self, f"LocalLayer_{g}") for g in GROUPS]

    def set_bn_momentum(self, momentum):
        for b in self._blocks():
            b.set_bn_momentum(momentum)

    def set_training_status(self, is_train):
        self.is_train = is_train
        for b in self._blocks():
            b.set_training_status(is_train)

    def set_augment(self, augment):
        self.augment = augment
        for b in self._blocks():
            b.set_augment(augment)

    def receptive_field(self):
        return self.LocalLayer_Torso.receptive_field()

    def forward(self, x, param):
        """x (B, RF, J, Cin), param (B, extrinsic_dim) -> (B, 1, J, 3), freshly allocated."""
        return self._native_forward(x, param)


class RIETrajectoryModel(_NativeNet):
    """Root-trajectory network.  Same ctor/forward as lib/model/rie.py:437-559."""

    _net_kind = "trj"

    def __init__(self, num_joints_in, in_features, num_joints_out, filter_widths, causal=False, dropout=0.2,
                 latten_features=256, channels=1024, dense=False, is_train=True, Optimize1f=True, stage=1,
                 extrinsic_dim=12, embedd_dim=64):
        super().__init__()
        # the trajectory net never has fuse blocks whatever `stage` says (rie.py:491-494)
        spec = _spec_from_ctor(num_joints_in, in_features, filter_widths, latten_features, channels, 1, extrinsic_dim, embedd_dim)
        self.augment, self.is_train = False, is_train
        self.num_joints_in, self.num_joints_out, self.in_features = num_joints_in, num_joints_out, in_features
        self.latten_features, self.stage = latten_features, stage
        self.LocalLayer = TemporalBlock(num_joints_in * 3, in_features, num_joints_out, filter_widths, causal, dropout,
                                        channels, latten_features, dense, is_train, Optimize1f)
        self.pad = (self.receptive_field() - 1) // 2
        self.GlobalInfo = FCBlock(num_joints_in * in_features, latten_features, FC_WIDTH, 2)
        self.camera_embedding = extrinsic_dim > 0 and embedd_dim > 0
        self.extrinsic_dim, self.embedd_dim = extrinsic_dim, embedd_dim
        if self.camera_embedding:
            self.embedder = Embedding(in_channels=extrinsic_dim, out_channels=embedd_dim)
        self.out_features_dim = latten_features * 2 + embedd_dim
        self.Integration = FCBlock(self.out_features_dim, 3, FC_WIDTH, 1)
        self._init_native(spec)

    def set_bn_momentum(self, momentum):
        self.LocalLayer.set_bn_momentum(momentum)

    def set_training_status(self, is_train):
        self.is_train = is_train
        self.LocalLayer.set_training_status(is_train)

    def set_augment(self, augment):
        self.augment = augment
        self.LocalLayer.set_augment(augment)

    def receptive_field(self):
        return self.LocalLayer.receptive_field()

    def forward(self, x, param):
        """x (B, RF, J, Cin), param (B, extrinsic_dim) -> (B, 1, 1, 3), freshly allocated."""
        return self._native_forward(x, param)


def fused_lifter(pos_model: RIEModel, trj_model: RIETrajectoryModel, precision: Optional[str] = None,
                 device: Optional[int] = None) -> Lifter:
    """One plan for both modules (grouped launches across the 5 joint groups + trajectory net)."""
    pos_model = getattr(pos_model, "module", pos_model)
    trj_model = getattr(trj_model, "module", trj_model)
    return Lifter(pos_model._spec, pos_model.state_dict(), trj_model.state_dict(), precision=precision or _precision(), device=device)


class Model(object):
    """Factory with the reference's surface (lib/model/__init__.py:5-62)."""

    def __init__(self, model_config, data_config=None, is_train=True):
        if model_config['CAMERA_EMBDDING']:
            extrinsic_dim, embedd_dim = model_config['EXTRINSIC_DIM'], model_config['EMBEDD_DIM']
        else:
            extrinsic_dim, embedd_dim = 0, 0
        if model_config['MODEL'] != 'RIE':
            raise ValueError('Unrecognized mdoel {}'.format(model_config['MODEL']))
        widths = [int(x) for x in model_config['ARCHITECTURE'].split(',')]
        kw = dict(filter_widths=widths, causal=model_config['CAUSAL'], dropout=model_config['DROPOUT'],
                  channels=model_config['CHANNELS'], latten_features=model_config['LATENT_FEATURES_DIM'],
                  dense=model_config['DENSE'], is_train=is_train, Optimize1f=not model_config['DISABLE_OPTIMIZATIONS'],
                  stage=model_config['STAGE'], extrinsic_dim=extrinsic_dim, embedd_dim=embedd_dim)
        n = model_config['NUM_KPTS']
        pos_model = RIEModel(n, model_config['INPUT_DIM'], n, **kw)
        trj_model = RIETrajectoryModel(n, model_config['INPUT_DIM'], n, **kw) if model_config['TRAJECTORY_MODEL'] else None
        if torch.cuda.is_available():
            # keep the reference's nn.DataParallel wrapper (checkpoint keys carry "module."), but pin it to one
            # device: multi-GPU runs use one process per GPU (ray3d_b200.dist), not per-call weight broadcast.
            dev = [torch.cuda.current_device()]
            pos_model = nn.DataParallel(pos_model, device_ids=dev).cuda()
            trj_model = nn.DataParallel(trj_model, device_ids=dev).cuda() if trj_model is not None else None
        self.pos_model, self.trj_model = pos_model, trj_model
        if trj_model is not None:
            link_models(pos_model, trj_model)

    def get_pos_model(self):
        return self.pos_model

    def get_trj_model(self):
        return self.trj_model
