"""CPU oracle: scalar rational-quadratic spline (RQS) bijector, torch float64.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`cnf_ot_b200/`) may
import this module; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` do, and only as the
checker / the timed CPU baseline.

What it restates
----------------
The reference builds every spline as
`distrax.RationalQuadraticSpline(params, range_min=-10., range_max=10.,
min_knot_slope=1e-4, boundary_slopes='unconstrained')`
(`/root/reference/cnf_ot/models/flows.py:124-132`) and calls its
`forward_and_log_det` / `inverse_and_log_det`
(`/root/reference/cnf_ot/models/autoregressive.py:100,130`).  distrax is a
third-party dependency that is NOT vendored under `/root/reference` and is not
version-pinned anywhere in it (no requirements file, `pyproject.toml:7-10`
omits it).  This file restates the published algorithm of
`distrax/_src/bijectors/rational_quadratic_spline.py` (distrax 0.1.x; the
Durkan et al. 2019 neural-spline-flow formulas, the same rational-quadratic
form the reference's own `cnf_ot/models/nsf_symbol.py:6-10` writes down).

PARITY OF THIS FILE UNPINNED in absolute value (the rest of oracle/ is pinned to the reference's own
code, which calls this file where it would call distrax: `tests/golden/refshim.py`): neither JAX nor
distrax can be installed in this environment, and the reference's only test of this path
(`/root/reference/tests/test_rqs_accuracy.py`) holds no golden vectors, only
invariants (round trips, log-det vs autodiff Jacobian, boundary round trips,
all < 1e-12 in float64).  That test file is executed UNMODIFIED against this restatement
(`tests/golden/run_reference_tests.py`, on stand-ins for jax / distrax) and passes; `tests/test_oracle_rqs.py` re-runs
the same invariants, and `tests/golden/ref_rqs_symbolic_k*.npz` pin the in-bin map and its derivative to the
reference's own symbolic statement (`cnf_ot/models/nsf_symbol.py:3-10`).
"""
from __future__ import annotations

import math
from typing import Tuple

import torch

Tensor = torch.Tensor

# Spline constants used by the reference flow (flows.py:124-132; min_bin_size
# is the distrax constructor default).
RANGE_MIN = -10.0
RANGE_MAX = 10.0
MIN_BIN_SIZE = 1e-4
MIN_KNOT_SLOPE = 1e-4


def num_bins_of(params: Tensor) -> int:
  p = params.shape[-1]
  if p % 3 != 1 or p < 4:
    raise ValueError(f"last dim of spline params must be 3*K+1, got {p}")
  return (p - 1) // 3


def normalize_knots(
  params: Tensor,
  range_min: float = RANGE_MIN,
  range_max: float = RANGE_MAX,
  min_bin_size: float = MIN_BIN_SIZE,
  min_knot_slope: float = MIN_KNOT_SLOPE,
) -> Tuple[Tensor, Tensor, Tensor]:
  """Raw params (..., 3K+1) -> knot x positions, y positions, slopes (..., K+1).

  Restates the distrax RationalQuadraticSpline constructor: softmax bin sizes
  rescaled so every bin is >= min_bin_size, cumulative sums padded with the
  exact range ends, softplus slopes offset so a raw 0 gives slope 1.
  """
  k = num_bins_of(params)
  if k * min_bin_size > range_max - range_min:
    raise ValueError("min_bin_size too large for the range")
  if min_knot_slope >= 1.0:
    raise ValueError("min_knot_slope must be < 1")
  span = range_max - range_min
  uw, uh, us = params[..., :k], params[..., k:2 * k], params[..., 2 * k:]
  widths = torch.softmax(uw, dim=-1) * (span - k * min_bin_size) + min_bin_size
  heights = torch.softmax(uh, dim=-1) * (span - k * min_bin_size) + min_bin_size
  lo = torch.full(params.shape[:-1] + (1, ), range_min, dtype=params.dtype)
  hi = torch.full(params.shape[:-1] + (1, ), range_max, dtype=params.dtype)
  x_pos = torch.cat(
    [lo, range_min + torch.cumsum(widths[..., :-1], dim=-1), hi], dim=-1
  )
  y_pos = torch.cat(
    [lo, range_min + torch.cumsum(heights[..., :-1], dim=-1), hi], dim=-1
  )
  offset = math.log(math.exp(1.0 - min_knot_slope) - 1.0)
  slopes = torch.nn.functional.softplus(us + offset) + min_knot_slope
  return x_pos, y_pos, slopes


def _select_bin(v: Tensor, pos: Tensor):
  """One-hot bin mask with the first-bin fallback, plus the tail flags.

  Half-open bins [pos[k], pos[k+1]); when no bin matches (v outside the range)
  bin 0 is used so the spline branch stays finite.  The reference never forms
  an integer index; `idx` here is argmax of the mask (both tails report 0).
  """
  vv = v.unsqueeze(-1)
  below = v <= pos[..., 0]
  above = v >= pos[..., -1]
  mask = (vv >= pos[..., :-1]) & (vv < pos[..., 1:])
  none = ~mask.any(dim=-1, keepdim=True)
  first = torch.zeros_like(mask)
  first[..., 0] = True
  mask = torch.where(none, first, mask)
  idx = mask.to(torch.int64).argmax(dim=-1)
  return mask.to(pos.dtype), idx, below, above


def _gather(mask: Tensor, arr: Tensor):
  left = (mask * arr[..., :-1]).sum(-1)
  right = (mask * arr[..., 1:]).sum(-1)
  return left, right


def rqs_forward(x: Tensor, params: Tensor, **kw):
  """y = S(x), log|dS/dx|, bin index.  x: (...,), params: (..., 3K+1)."""
  x_pos, y_pos, slopes = normalize_knots(params, **kw)
  x_pos, y_pos, slopes = (a.to(x.dtype) for a in (x_pos, y_pos, slopes))
  mask, idx, below, above = _select_bin(x, x_pos)
  x0, x1 = _gather(mask, x_pos)
  y0, y1 = _gather(mask, y_pos)
  d0, d1 = _gather(mask, slopes)
  bw = x1 - x0
  bh = y1 - y0
  sl = bh / bw
  z = torch.clamp((x - x0) / bw, 0.0, 1.0)
  z2 = z * z
  z1mz = z - z2
  omz2 = (1.0 - z)**2
  st = d1 + d0 - 2.0 * sl
  num = bh * (sl * z2 + d0 * z1mz)
  den = sl + st * z1mz
  y = y0 + num / den
  logdet = 2.0 * torch.log(sl) + torch.log(
    d1 * z2 + 2.0 * sl * z1mz + d0 * omz2
  ) - 2.0 * torch.log(den)
  # linear tails outside the knot range
  y = torch.where(below, (x - x_pos[..., 0]) * slopes[..., 0] + y_pos[..., 0], y)
  y = torch.where(
    above, (x - x_pos[..., -1]) * slopes[..., -1] + y_pos[..., -1], y
  )
  logdet = torch.where(below, torch.log(slopes[..., 0]), logdet)
  logdet = torch.where(above, torch.log(slopes[..., -1]), logdet)
  return y, logdet, idx


def _stable_root(a: Tensor, b: Tensor, c: Tensor) -> Tensor:
  disc = b * b - 4.0 * a * c
  tiny = torch.finfo(disc.dtype).tiny
  root = torch.sqrt(torch.clamp(disc, min=tiny))
  root = torch.where(disc > 0.0, root, torch.zeros_like(root))
  num = torch.where(b >= 0, 2.0 * c, -b + root)
  den = torch.where(b >= 0, -b - root, 2.0 * a)
  return num / den


def rqs_inverse(y: Tensor, params: Tensor, **kw):
  """x = S^{-1}(y), log|dS^{-1}/dy|, bin index."""
  x_pos, y_pos, slopes = normalize_knots(params, **kw)
  x_pos, y_pos, slopes = (a.to(y.dtype) for a in (x_pos, y_pos, slopes))
  mask, idx, below, above = _select_bin(y, y_pos)
  x0, x1 = _gather(mask, x_pos)
  y0, y1 = _gather(mask, y_pos)
  d0, d1 = _gather(mask, slopes)
  bw = x1 - x0
  bh = y1 - y0
  sl = bh / bw
  w = torch.clamp((y - y0) / bh, 0.0, 1.0)
  st = d1 + d0 - 2.0 * sl
  qc = -sl * w
  qb = d0 - st * w
  qa = sl - qb
  z = torch.clamp(_stable_root(qa, qb, qc), 0.0, 1.0)
  x = bw * z + x0
  z2 = z * z
  z1mz = z - z2
  omz2 = (1.0 - z)**2
  den = sl + st * z1mz
  logdet = -2.0 * torch.log(sl) - torch.log(
    d1 * z2 + 2.0 * sl * z1mz + d0 * omz2
  ) + 2.0 * torch.log(den)
  x = torch.where(below, (y - y_pos[..., 0]) / slopes[..., 0] + x_pos[..., 0], x)
  x = torch.where(
    above, (y - y_pos[..., -1]) / slopes[..., -1] + x_pos[..., -1], x
  )
  logdet = torch.where(below, -torch.log(slopes[..., 0]), logdet)
  logdet = torch.where(above, -torch.log(slopes[..., -1]), logdet)
  return x, logdet, idx
