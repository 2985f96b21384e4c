"""Parameter-blob layout: haiku pytree <-> one flat fp32 buffer.

The reference keeps its parameters as a two-level haiku dict (SURVEY.md A.3;
created by tracing `flow.log_prob`, /root/reference/cnf_ot/models/flows.py:215,
/root/reference/cnf_ot/mfc/solvers.py:54):

  params["~"]["first"]                                (1, P)    P = 3*num_bins + 1
  params[f"mlp_layer{l}_d{d}/~/linear_{m}"]["w"|"b"]   (in, H), (H,)
  params[f"linear_out_layer{l}_d{d}"]["w"|"b"]         (H, P), (P,)

The kernels read ONE flat buffer (same layout for the gradient):

  [ first (Pp) ]
  for l in range(L): for d in range(1, D):
    W0 ((d+1) x H) b0 (H) { Wm (H x H) bm (H) }_{m=1..M-1} Wout (H x Pp) bout (Pp)

with Pp = P rounded up to a multiple of 4 (rows stay 16-byte aligned; padding
is zero).  The offsets here must match `make_layout` / `mlp_offset` in
csrc/flow_math.cuh; `tests/test_layout.py` checks them against the C ABI.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Iterator, List, Tuple

import torch

Params = Dict[str, Dict[str, torch.Tensor]]


def mlp_key(layer: int, d: int, m: int) -> str:
  return f"mlp_layer{layer}_d{d}/~/linear_{m}"


def out_key(layer: int, d: int) -> str:
  return f"linear_out_layer{layer}_d{d}"


@dataclass(frozen=True)
class FlowShape:
  """Static arguments of RQSFlow (flows.py:178-199) as the kernels see them."""
  dim: int
  num_layers: int
  mlp_layers: int
  hidden: int
  num_bins: int
  range_min: float = -10.0
  range_max: float = 10.0
  min_bin_size: float = 1e-4
  min_knot_slope: float = 1e-4
  # cond_shape == (1,) (time-conditioned, the MFC solvers) or (0,) (unconditional, cnf_ot/dr/trainers.py:41-68).
  # An unconditional flow keeps the SAME blob layout: row 0 of every input matrix (the weights of t) exists but
  # is zero and is not a leaf, and the kernels are called with t = 0, so its gradient is exactly zero too.
  conditional: bool = True

  @property
  def P(self) -> int:
    return 3 * self.num_bins + 1

  @property
  def Pp(self) -> int:
    return (self.P + 3) // 4 * 4

  @property
  def mlp_const(self) -> int:
    H, M, Pp = self.hidden, self.mlp_layers, self.Pp
    return H + (M - 1) * (H * H + H) + H * Pp + Pp

  @property
  def layer_stride(self) -> int:
    D, H = self.dim, self.hidden
    return (D - 1) * self.mlp_const + H * ((D - 1) * (D + 2) // 2)

  @property
  def blob_size(self) -> int:
    return self.Pp + self.num_layers * self.layer_stride

  def mlp_offset(self, layer: int, d: int) -> int:
    H = self.hidden
    return (self.Pp + layer * self.layer_stride + (d - 1) * self.mlp_const +
            H * ((d - 1) * (d + 2) // 2))

  def linear_offset(self, layer: int, d: int, m: int, bias: bool) -> int:
    """m < mlp_layers: hidden linear m; m == mlp_layers: the output linear."""
    H, M, Pp = self.hidden, self.mlp_layers, self.Pp
    off = self.mlp_offset(layer, d)
    n_in = d + 1
    if m == 0:
      return off + (n_in * H if bias else 0)
    off += n_in * H + H + (m - 1) * (H * H + H)
    if m < M:
      return off + (H * H if bias else 0)
    return off + (H * Pp if bias else 0)

  def param_count(self) -> int:
    """Number of reference parameters (un-padded), e.g. 1200 for mfc.yaml."""
    P, H, M = self.P, self.hidden, self.mlp_layers
    n = P
    c = 1 if self.conditional else 0
    for _ in range(self.num_layers):
      for d in range(1, self.dim):
        n += (d + c) * H + H + (M - 1) * (H * H + H) + H * P + P
    return n

  def leaves(self) -> Iterator[Tuple[str, str, Tuple[int, ...], int, int]]:
    """(module, leaf, shape, blob offset, blob row stride) in haiku order."""
    P, Pp, H, M = self.P, self.Pp, self.hidden, self.mlp_layers
    yield "~", "first", (1, P), 0, Pp
    for l in range(self.num_layers):
      for d in range(1, self.dim):
        fan_in = d + 1 if self.conditional else d
        for m in range(M):
          skip = H if (m == 0 and not self.conditional) else 0   # unconditional: the t row is not a leaf
          yield mlp_key(l, d, m), "w", (fan_in, H), self.linear_offset(l, d, m, False) + skip, H
          yield mlp_key(l, d, m), "b", (H, ), self.linear_offset(l, d, m, True), H
          fan_in = H
        yield out_key(l, d), "w", (H, P), self.linear_offset(l, d, M, False), Pp
        yield out_key(l, d), "b", (P, ), self.linear_offset(l, d, M, True), Pp


def pack(shape: FlowShape, params: Params, dtype=torch.float32) -> torch.Tensor:
  """haiku pytree -> flat blob (CPU tensor)."""
  blob = torch.zeros(shape.blob_size, dtype=dtype)
  for mod, leaf, shp, off, stride in shape.leaves():
    v = torch.as_tensor(params[mod][leaf]).detach().to("cpu", dtype)
    if tuple(v.shape) != shp:
      raise ValueError(f"{mod}/{leaf}: expected shape {shp}, got {tuple(v.shape)}")
    rows = v.reshape(-1, shp[-1])
    for r in range(rows.shape[0]):
      blob[off + r * stride:off + r * stride + shp[-1]] = rows[r]
  return blob


def unpack(shape: FlowShape, blob: torch.Tensor, like: Params = None) -> Params:
  """flat blob -> haiku pytree (drops padding).  dtypes follow `like` if given."""
  blob = blob.detach().to("cpu")
  out: Params = {}
  for mod, leaf, shp, off, stride in shape.leaves():
    nrow = 1
    for s in shp[:-1]:
      nrow *= s
    rows = [blob[off + r * stride:off + r * stride + shp[-1]] for r in range(nrow)]
    v = torch.stack(rows).reshape(shp).clone()
    if like is not None:
      v = v.to(like[mod][leaf].dtype)
    out.setdefault(mod, {})[leaf] = v
  return out
