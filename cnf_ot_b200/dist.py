"""Data-parallel plumbing: row sharding and the step's single all-reduce.

The reference is single-device; the path shards over independent sample rows
(SURVEY.md §8e): rank g takes rows [g*n/G, (g+1)*n/G) of every term, the kernels
produce SUMS with the global 1/B, 1/b factors folded in, and one
`all_reduce(SUM)` of the contiguous [gradient | loss slots] buffer yields the
whole-batch result on every rank.  Two transports:
  * `PeerExchange` (GPUs of one node): the step's final reduction kernel itself exchanges
    the buffer through peer-mapped memory over NVLink / NVSwitch (cnfot_mfc_step_dp) -- one
    kernel does the rank's reduction and the all-reduce, no NCCL launch;
  * `all_reduce_sum`: torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch


def rank_world() -> Tuple[int, int]:
  import torch.distributed as td
  if td.is_available() and td.is_initialized():
    return td.get_rank(), td.get_world_size()
  return 0, 1


def shard(n: int, rank: int, world: int) -> slice:
  """Contiguous, balanced row range of rank `rank` (sizes differ by at most one)."""
  if world < 1 or not (0 <= rank < world):
    raise ValueError("bad rank/world")
  base, rem = divmod(n, world)
  lo = rank * base + min(rank, rem)
  return slice(lo, lo + base + (1 if rank < rem else 0))


def all_reduce_sum(buf: torch.Tensor) -> torch.Tensor:
  import torch.distributed as td
  if td.is_available() and td.is_initialized() and td.get_world_size() > 1:
    td.all_reduce(buf, op=td.ReduceOp.SUM)
  return buf


class PeerExchange:
  """Peer-mapped exchange buffers for the fused step + all-reduce (include/cnfot.h,
  cnfot_peer_desc).  torch's symmetric memory does the cross-process mapping; this class only
  owns the buffers, the epoch counter and the descriptor.  Raises if symmetric memory is
  unavailable -- callers then fall back to `all_reduce_sum`."""

  def __init__(self, shape, device, group=None):
    import torch.distributed as td
    import torch.distributed._symmetric_memory as symm_mem
    from . import _lib
    self._lib = _lib
    group = td.group.WORLD if group is None else group
    self.rank, self.world = td.get_rank(group), td.get_world_size(group)
    if self.world > 8:
      raise RuntimeError("PeerExchange covers the GPUs of one node (world <= 8)")
    lib = _lib.load()
    desc = _lib.flow_desc(shape)
    self.shape = shape
    n_x = lib.cnfot_dp_exchange_floats(desc, self.world)
    n_f = lib.cnfot_dp_flag_count(desc, self.world)
    self.xbuf = symm_mem.empty(n_x, dtype=torch.float32, device=device)
    self.flags = symm_mem.empty(n_f, dtype=torch.int32, device=device)
    self.xbuf.zero_()
    self.flags.zero_()
    hx = symm_mem.rendezvous(self.xbuf, group=group.group_name)
    hf = symm_mem.rendezvous(self.flags, group=group.group_name)
    self._handles = (hx, hf)
    self._xptrs = [int(p) for p in hx.buffer_ptrs]
    self._fptrs = [int(p) for p in hf.buffer_ptrs]
    torch.cuda.synchronize(device)
    td.barrier(group)  # every rank's flags are zeroed before the first epoch
    self.epoch = 0

  def next_desc(self, shape):
    d = self.peek_desc(shape)
    self.epoch += 1
    return d

  def peek_desc(self, shape):
    """The descriptor of the NEXT exchange (epoch + 1) without consuming it."""
    if shape != self.shape:
      raise ValueError("PeerExchange was sized for another flow shape")
    d = self._lib.PeerDesc()
    d.rank, d.world, d.epoch = self.rank, self.world, (self.epoch + 1) & 0xFFFFFFFF or 1
    for k in range(self.world):
      d.xbuf[k] = self._xptrs[k]
      d.flags[k] = self._fptrs[k]
    return d
