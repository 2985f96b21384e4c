"""Data-parallel plumbing: row sharding and the step's single all-reduce.

The reference is single-device; the path shards over independent sample rows
(SURVEY.md §8e): rank g takes rows [g*n/G, (g+1)*n/G) of every term, the kernels
produce SUMS with the global 1/B, 1/b factors folded in, and one
`all_reduce(SUM)` of the contiguous [gradient | loss slots] buffer (NCCL over
NVLink on GPUs, gloo in the CPU tests) yields the whole-batch result on every
rank.
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
