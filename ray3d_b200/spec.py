"""Static description of the Ray3D lifting networks (pose + trajectory).

This module is pure data/arith: joint-group tables, layer shapes, state_dict key
order.  It is what the C-ABI plan (``include/ray3d_b200.h``) is configured from and what
the synthetic-weight generator walks.  Nothing here touches torch or CUDA.

Reference behaviour it encodes (all paths relative to the reference checkout):
  * joint groups / channel slices ......... lib/model/rie.py:306-357
  * output joint order .................... lib/model/rie.py:415-432
  * TemporalBlock layer list .............. lib/model/rie.py:13-63
  * FCBlock / Linear layer list ........... lib/model/rie.py:108-157
  * Embedding ............................. lib/model/embedding.py:5-13
  * "current frame" = T // in_features .... lib/model/rie.py:290,304
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

GROUPS = ("Torso", "LArm", "RArm", "LLeg", "RLeg")

# Input joints of each group, in the order their channels are concatenated
# (lib/model/rie.py:306-357; the in_features==2 tables select the same joints).
GROUP_JOINTS: Dict[int, Dict[str, Tuple[int, ...]]] = {
    17: {"Torso": (0, 7, 8, 9, 10), "LArm": (14, 15, 16), "RArm": (11, 12, 13),
         "LLeg": (1, 2, 3), "RLeg": (4, 5, 6)},
    15: {"Torso": (0, 1, 14), "LArm": (2, 3, 4), "RArm": (5, 6, 7),
         "LLeg": (8, 9, 10), "RLeg": (11, 12, 13)},
    14: {"Torso": (0, 7), "LArm": (8, 9, 10), "RArm": (11, 12, 13),
         "LLeg": (4, 5, 6), "RLeg": (1, 2, 3)},
}

# Output assembly (lib/model/rie.py:426-431): list of (group, first_joint_in_head, count)
# in the order they are concatenated into output joint slots 0..J-1.
OUTPUT_ORDER: Dict[int, Tuple[Tuple[str, int, int], ...]] = {
    17: (("Torso", 0, 1), ("LLeg", 0, 3), ("RLeg", 0, 3), ("Torso", 1, 4), ("RArm", 0, 3), ("LArm", 0, 3)),
    15: (("Torso", 0, 2), ("LLeg", 0, 3), ("RLeg", 0, 3), ("RArm", 0, 3), ("LArm", 0, 3), ("Torso", 2, 1)),
    14: (("Torso", 0, 1), ("LLeg", 0, 3), ("RLeg", 0, 3), ("RArm", 0, 3), ("LArm", 0, 3), ("Torso", 1, 1)),
}

FC_WIDTH = 1024        # hard-coded linear_size in the reference (rie.py:226,232,245-253,483,494)
EMBED_MID = 32         # Embedding mid_channels default (embedding.py:5)
BN_EPS = 1e-5          # nn.BatchNorm1d default
SLOPE_NET = 0.2        # nn.LeakyReLU(0.2) in TemporalBlock/Linear/FCBlock
SLOPE_EMBED = 0.01     # nn.LeakyReLU() default in Embedding


@dataclass(frozen=True)
class NetSpec:
    """Hyper-parameters shared by the pose and trajectory networks."""
    num_joints: int = 17
    in_features: int = 3
    filter_widths: Tuple[int, ...] = (3, 3, 3)
    channels: int = 256
    latent: int = 256
    stage: int = 1
    extrinsic_dim: int = 2
    embed_dim: int = 64

    def __post_init__(self):
        if self.num_joints not in GROUP_JOINTS:
            raise ValueError(f"unsupported joint count {self.num_joints} (reference supports 17/15/14)")
        if self.in_features not in (2, 3):
            raise ValueError("in_features must be 2 (pixel/intrinsic encoding) or 3 (ray encoding)")
        if not self.filter_widths or any(w < 1 or w % 2 == 0 for w in self.filter_widths):
            raise ValueError("filter widths must be odd and positive")

    def replace_stage(self, stage: int) -> "NetSpec":
        """Same architecture with another ``stage`` (the trajectory net ignores it, rie.py:491-494)."""
        return dataclasses.replace(self, stage=stage)

    # -- derived ---------------------------------------------------------------------------
    @property
    def receptive_field(self) -> int:
        rf = 1
        for w in self.filter_widths:
            rf *= w
        return rf

    @property
    def camera_embedding(self) -> bool:
        return self.extrinsic_dim > 0 and self.embed_dim > 0

    @property
    def current_frame(self) -> int:
        """The reference's idea of "current frame": T // in_features (rie.py:290,304)."""
        return self.receptive_field // self.in_features

    @property
    def pos_feature_dim(self) -> int:
        d = self.latent * (2 if self.stage == 1 else 3)
        return d + (self.embed_dim if self.camera_embedding else 0)

    @property
    def trj_feature_dim(self) -> int:
        return self.latent * 2 + (self.embed_dim if self.camera_embedding else 0)

    def group_joints(self, g: str) -> Tuple[int, ...]:
        return GROUP_JOINTS[self.num_joints][g]

    def group_in_channels(self, g: str) -> int:
        return 3 * len(self.group_joints(g)) * self.in_features

    def group_out_dim(self, g: str) -> int:
        return 3 * len(self.group_joints(g))

    def output_slots(self) -> List[Tuple[str, int]]:
        """For each output joint slot 0..J-1: (group, joint index inside that group's head)."""
        slots: List[Tuple[str, int]] = []
        for g, first, n in OUTPUT_ORDER[self.num_joints]:
            slots.extend((g, first + i) for i in range(n))
        assert len(slots) == self.num_joints
        return slots

    def level_lengths(self) -> List[int]:
        """Time length after expand_conv and after every further strided level."""
        t = self.receptive_field
        out = []
        for w in self.filter_widths:
            t //= w
            out.append(t)
        return out


# ------------------------------------------------------------------------------------------
# state_dict layout (names, shapes, registration order) -- the drop-in contract with
# lib/utils/utils.py:208-218 (load_weight) and trainer.py:161 (load_state_dict strict).
# ------------------------------------------------------------------------------------------
Shape = Tuple[int, ...]


def _bn(prefix: str, c: int) -> List[Tuple[str, Shape, str]]:
    return [(prefix + ".weight", (c,), "bn_weight"), (prefix + ".bias", (c,), "bn_bias"),
            (prefix + ".running_mean", (c,), "bn_mean"), (prefix + ".running_var", (c,), "bn_var"),
            (prefix + ".num_batches_tracked", (), "bn_count")]


def _linear(prefix: str, cin: int, cout: int) -> List[Tuple[str, Shape, str]]:
    return [(prefix + ".weight", (cout, cin), "weight"), (prefix + ".bias", (cout,), "bias")]


def temporal_block_entries(prefix: str, in_ch: int, spec: NetSpec) -> List[Tuple[str, Shape, str]]:
    """Entries of one TemporalBlock in nn.Module registration order (rie.py:29-63)."""
    c = spec.channels
    w = spec.filter_widths
    e: List[Tuple[str, Shape, str]] = []
    e += _bn(prefix + ".expand_bn", c)
    e += [(prefix + ".shrink.weight", (spec.latent, c, 1), "weight"), (prefix + ".shrink.bias", (spec.latent,), "bias")]
    e += [(prefix + ".expand_conv.weight", (c, in_ch, w[0]), "weight")]
    for i in range(1, len(w)):
        e += [(prefix + f".layers_conv.{2 * (i - 1)}.weight", (c, c, w[i]), "weight")]
        e += [(prefix + f".layers_conv.{2 * (i - 1) + 1}.weight", (c, c, 1), "weight")]
    for i in range(1, len(w)):
        e += _bn(prefix + f".layers_bn.{2 * (i - 1)}", c)
        e += _bn(prefix + f".layers_bn.{2 * (i - 1) + 1}", c)
    return e


def fc_block_entries(prefix: str, cin: int, cout: int, nblocks: int) -> List[Tuple[str, Shape, str]]:
    """Entries of one FCBlock in registration order (rie.py:140-157)."""
    h = FC_WIDTH
    e: List[Tuple[str, Shape, str]] = []
    e += _linear(prefix + ".fc_1", cin, h)
    e += _bn(prefix + ".bn_1", h)
    e += _linear(prefix + ".fc_2", h, cout)
    for i in range(nblocks):
        p = prefix + f".layers.{i}"
        e += _linear(p + ".w1", h, h)
        e += _bn(p + ".batch_norm1", h)
        e += _linear(p + ".w2", h, h)
        e += _bn(p + ".batch_norm2", h)
    return e


def embedding_entries(prefix: str, spec: NetSpec) -> List[Tuple[str, Shape, str]]:
    e: List[Tuple[str, Shape, str]] = []
    e += _linear(prefix + ".w1", spec.extrinsic_dim, EMBED_MID)
    e += _bn(prefix + ".b1", EMBED_MID)
    e += _linear(prefix + ".w2", EMBED_MID, spec.embed_dim)
    e += _bn(prefix + ".b2", spec.embed_dim)
    return e


def pos_state_entries(spec: NetSpec) -> List[Tuple[str, Shape, str]]:
    """(name, shape, kind) for RIEModel.state_dict() in order (rie.py:197-253)."""
    e: List[Tuple[str, Shape, str]] = []
    for g in GROUPS:
        e += temporal_block_entries(f"LocalLayer_{g}", spec.group_in_channels(g), spec)
    e += fc_block_entries("GlobalInfo", spec.num_joints * spec.in_features, spec.latent, 2)
    if spec.stage != 1:
        for i in range(5):
            e += fc_block_entries(f"FuseBlocks.{i}", spec.latent * 4, spec.latent, 1)
    if spec.camera_embedding:
        e += embedding_entries("embedder", spec)
    for g in GROUPS:
        e += fc_block_entries(f"Integration_{g}", spec.pos_feature_dim, spec.group_out_dim(g), 1)
    return e


def trj_state_entries(spec: NetSpec) -> List[Tuple[str, Shape, str]]:
    """(name, shape, kind) for RIETrajectoryModel.state_dict() in order (rie.py:465-494)."""
    e: List[Tuple[str, Shape, str]] = []
    e += temporal_block_entries("LocalLayer", 3 * spec.num_joints * spec.in_features, spec)
    e += fc_block_entries("GlobalInfo", spec.num_joints * spec.in_features, spec.latent, 2)
    if spec.camera_embedding:
        e += embedding_entries("embedder", spec)
    e += fc_block_entries("Integration", spec.trj_feature_dim, 3, 1)
    return e


def flops_per_sequence(spec: NetSpec, with_trj: bool = True) -> float:
    """2 x MACs of every conv/linear for one RF-long window (matches SURVEY section 8a table)."""
    def tblock(in_ch: int) -> int:
        lens = spec.level_lengths()
        c, w = spec.channels, spec.filter_widths
        macs = lens[0] * in_ch * w[0] * c
        for i in range(1, len(w)):
            macs += lens[i] * (c * w[i] * c + c * c)
        macs += lens[-1] * c * spec.latent
        return macs

    def fcb(cin: int, cout: int, n: int) -> int:
        return cin * FC_WIDTH + n * 2 * FC_WIDTH * FC_WIDTH + FC_WIDTH * cout

    def emb() -> int:
        return (spec.extrinsic_dim * EMBED_MID + EMBED_MID * spec.embed_dim) if spec.camera_embedding else 0

    jc = spec.num_joints * spec.in_features
    macs = sum(tblock(spec.group_in_channels(g)) for g in GROUPS)
    macs += fcb(jc, spec.latent, 2) + emb()
    if spec.stage != 1:
        macs += 5 * fcb(4 * spec.latent, spec.latent, 1)
    macs += sum(fcb(spec.pos_feature_dim, spec.group_out_dim(g), 1) for g in GROUPS)
    if with_trj:
        macs += tblock(3 * jc) + fcb(jc, spec.latent, 2) + emb() + fcb(spec.trj_feature_dim, 3, 1)
    return 2.0 * macs


def weight_count(spec: NetSpec, with_trj: bool = True) -> int:
    def n(entries):
        tot = 0
        for _, shp, kind in entries:
            if kind == "bn_count":
                continue
            k = 1
            for d in shp:
                k *= d
            tot += k
        return tot
    return n(pos_state_entries(spec)) + (n(trj_state_entries(spec)) if with_trj else 0)
