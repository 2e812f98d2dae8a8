"""Multi-GPU layout: one process per GPU, sequences sharded, one all-gather of the outputs.

The reference's only multi-GPU mechanism is nn.DataParallel (lib/model/__init__.py:51-53), which
re-broadcasts all weights and scatters/gathers through GPU 0 on every call.  Windows are
independent in eval mode (no cross-sample op), so here each rank keeps a resident copy of the
packed weights, lifts its own contiguous shard of the batch and the (B/world, J+1, 3) results are
combined with a single ``all_gather_into_tensor`` (NCCL over NVLink/NVSwitch on GPUs; gloo in the
CPU tests of the host logic).  No data-path collective exists besides that gather.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced shard [lo, hi) of `batch` sequences for `rank` (first ranks take the remainder)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_outputs(both: torch.Tensor, trj: torch.Tensor) -> torch.Tensor:
    """(b,1,J,3) pos+trj and (b,1,1,3) root -> one (b, J+1, 3) block so a single collective moves both."""
    return torch.cat((both.reshape(both.shape[0], -1, 3), trj.reshape(trj.shape[0], 1, 3)), dim=1).contiguous()


def unpack_outputs(packed: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    J = packed.shape[1] - 1
    return packed[:, :J].reshape(-1, 1, J, 3), packed[:, J:].reshape(-1, 1, 1, 3)


def gather_outputs(local: torch.Tensor, batch: int, group=None) -> torch.Tensor:
    """All-gather rank-local (b_r, J+1, 3) blocks into the full (batch, J+1, 3) tensor on every rank.

    Shards may differ by one row when world does not divide batch: blocks are padded to the largest
    shard for the collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_bounds(batch, r, world)[1] - shard_bounds(batch, r, world)[0] for r in range(world)]
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    mx = max(sizes)
    send = local
    if local.shape[0] != mx:
        send = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        send[: local.shape[0]] = local
    recv = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send.contiguous(), group=group)
    if all(s == mx for s in sizes):
        return recv
    recv = recv.view(world, mx, *local.shape[1:])
    return torch.cat([recv[r, : sizes[r]] for r in range(world)], dim=0)


class OverlappedGather:
    """Streaming form of ``gather_outputs`` for a loop of steps: the all-gather of step i runs on the backend's own
    communication stream while step i+1 is being computed.  ``depth`` send/receive buffer pairs rotate; ``submit``
    returns a ticket, ``result(ticket)`` makes the current stream (NCCL) or the host (gloo) wait for that collective
    and returns the full (batch, J+1, 3) tensor -- valid until ``depth`` further submissions.  All shards must have the
    same size (world divides batch); use ``gather_outputs`` for ragged batches."""

    def __init__(self, local_rows: int, joints: int, device, depth: int = 2, group=None, dtype=torch.float32):
        self.group, self.depth = group, depth
        world = dist.get_world_size(group)
        self.send = [torch.empty((local_rows, joints + 1, 3), dtype=dtype, device=device) for _ in range(depth)]
        self.recv = [torch.empty((world * local_rows, joints + 1, 3), dtype=dtype, device=device) for _ in range(depth)]
        self.work = [None] * depth
        self.seq = 0

    def submit(self, both: torch.Tensor, trj: torch.Tensor) -> int:
        k = self.seq % self.depth
        if self.work[k] is not None:          # the collective that last used this buffer pair
            self.work[k].wait()
        J = self.send[k].shape[1] - 1
        self.send[k][:, :J].copy_(both.reshape(-1, J, 3))
        self.send[k][:, J:].copy_(trj.reshape(-1, 1, 3))
        self.work[k] = dist.all_gather_into_tensor(self.recv[k], self.send[k], group=self.group, async_op=True)
        self.seq += 1
        return self.seq - 1

    def result(self, ticket: int) -> torch.Tensor:
        if not (self.seq - self.depth <= ticket < self.seq):
            raise ValueError(f"ticket {ticket} is not in flight (next {self.seq}, depth {self.depth})")
        k = ticket % self.depth
        if self.work[k] is not None:
            self.work[k].wait()
            self.work[k] = None
        return self.recv[k]

    def drain(self) -> None:
        for k in range(self.depth):
            if self.work[k] is not None:
                self.work[k].wait()
                self.work[k] = None


def lift_sharded(lift_fn: Callable[[torch.Tensor, torch.Tensor], Tuple[torch.Tensor, torch.Tensor]], uv_local: torch.Tensor,
                 cam_local: torch.Tensor, batch: int, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Lift this rank's shard with `lift_fn(uv, cam) -> (pos+trj, trj)` and all-gather the results.
    Returns full-batch (batch,1,J,3), (batch,1,1,3) on every rank."""
    both, trj = lift_fn(uv_local, cam_local)
    return unpack_outputs(gather_outputs(pack_outputs(both, trj), batch, group))
