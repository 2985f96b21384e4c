"""CPU oracle: the reference's time-conditioned autoregressive RQS flow, torch f64.

TEST INFRASTRUCTURE ONLY (see `oracle/rqs.py` header for who may import it).
PARITY PINNED to outputs of the reference's own source files run in this container
(`tests/golden/ref_flow_*.npz`: flows.py / autoregressive.py / conditional.py imported unmodified from
/root/reference on torch-f64 stand-ins for jax / haiku / distrax, `tests/golden/make_reference_golden.py`;
agreement 1e-10, `tests/test_reference_golden.py`), plus the invariants in `tests/test_oracle_*.py`.
The spline underneath is third-party (distrax) and stays a restatement: see `oracle/rqs.py`.

Restates, with explicit arrays instead of haiku/JAX tracing:
  * conditioner: `/root/reference/cnf_ot/models/flows.py:46-86`
      None  -> shared float32 parameter "~/first" of shape (1, P), zeros
      input -> hk.nets.MLP([H]*M, relu, activate_final=True) -> hk.Linear(P)
               (output layer zero-initialised => identity flow at init)
  * per-layer algebra: `/root/reference/cnf_ot/models/autoregressive.py:76-136`
  * stacking / direction / signs:
      `/root/reference/cnf_ot/models/flows.py:138-175` (alternating
      permutations, ConditionalInverse(ConditionalChain(layers)), N(0,I) base)
      `/root/reference/cnf_ot/models/conditional.py:147-177,217-243,316-321,376-402`

Parameter pytree (haiku naming):
  params["~"]["first"]                                  (1, P)   float32
  params[f"mlp_layer{l}_d{d}/~/linear_{m}"]["w"|"b"]     (in, H), (H,)
  params[f"linear_out_layer{l}_d{d}"]["w"|"b"]           (H, P), (P,)
for l in range(L), d in range(1, D) (d = position in the layer's permutation).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch

from . import rqs

Tensor = torch.Tensor
Params = Dict[str, Dict[str, Tensor]]


class FlowSpec:
  """Static description of a flow (what `RQSFlow(...)` closes over)."""

  def __init__(self, dim: int, num_layers: int, hidden_sizes: Sequence[int],
               num_bins: int, conditional: bool = True):
    # conditional=False: cond_shape=(0,), the unconditional flows of cnf_ot/dr/trainers.py:41-68
    # (autoregressive.py:94-98: the condition is concatenated only `if self.is_conditional`)
    self.conditional = bool(conditional)
    self.dim = int(dim)
    self.num_layers = int(num_layers)
    self.hidden_sizes = [int(h) for h in hidden_sizes]
    self.num_bins = int(num_bins)
    self.num_spline_params = 3 * self.num_bins + 1

  def permutation(self, layer: int) -> List[int]:
    # minimum_perm=True: identity on even layers, reversed on odd layers
    # (flows.py:141-143 with RQSFlow passing minimum_perm=True, :198)
    p = list(range(self.dim))
    return p if layer % 2 == 0 else p[::-1]

  def param_count(self) -> int:
    p = self.num_spline_params
    n = p
    for _ in range(self.num_layers):
      for d in range(1, self.dim):
        fan_in = d + 1 if self.conditional else d
        for h in self.hidden_sizes:
          n += fan_in * h + h
          fan_in = h
        n += fan_in * p + p
    return n


def mlp_key(layer: int, d: int, m: int) -> str:
  return f"mlp_layer{layer}_d{d}/~/linear_{m}"


def out_key(layer: int, d: int) -> str:
  return f"linear_out_layer{layer}_d{d}"


def _trunc_normal(gen: torch.Generator, shape, std: float) -> Tensor:
  # haiku's hk.Linear default: TruncatedNormal(stddev=1/sqrt(fan_in)), cut at
  # +-2 (in units of the untruncated normal's stddev).
  out = torch.empty(shape, dtype=torch.float64)
  flat = out.view(-1)
  filled = 0
  while filled < flat.numel():
    cand = torch.randn(2 * (flat.numel() - filled) + 16, generator=gen,
                       dtype=torch.float64)
    cand = cand[cand.abs() <= 2.0][:flat.numel() - filled]
    flat[filled:filled + cand.numel()] = cand
    filled += cand.numel()
  return out * std


def init_params(spec: FlowSpec, seed: int = 0) -> Params:
  """Reference initialisation: identity flow (zero `first`, zero out layers)."""
  gen = torch.Generator().manual_seed(seed)
  p = spec.num_spline_params
  params: Params = {"~": {"first": torch.zeros(1, p, dtype=torch.float32)}}
  for l in range(spec.num_layers):
    for d in range(1, spec.dim):
      fan_in = d + 1 if spec.conditional else d
      for m, h in enumerate(spec.hidden_sizes):
        params[mlp_key(l, d, m)] = {
          "w": _trunc_normal(gen, (fan_in, h), 1.0 / math.sqrt(fan_in)),
          "b": torch.zeros(h, dtype=torch.float64),
        }
        fan_in = h
      params[out_key(l, d)] = {
        "w": torch.zeros(fan_in, p, dtype=torch.float64),
        "b": torch.zeros(p, dtype=torch.float64),
      }
  return params


def perturb_params(params: Params, sigma: float, seed: int = 43) -> Params:
  """Benchmark/parity parameter set: hidden weights as
  initialised; biases, output layers and `first` ~ N(0, sigma^2)."""
  gen = torch.Generator().manual_seed(seed)
  out: Params = {}
  for mod in params:  # insertion order is deterministic
    out[mod] = {}
    for name, v in params[mod].items():
      hidden_w = name == "w" and mod.startswith("mlp_")
      if hidden_w:
        out[mod][name] = v.clone()
      else:
        noise = torch.randn(v.shape, generator=gen, dtype=torch.float64)
        out[mod][name] = (v.to(torch.float64) + sigma * noise).to(v.dtype)
  return out


def conditioner(spec: FlowSpec, params: Params, layer: int, d: int,
                inp: Optional[Tensor]) -> Tensor:
  """Raw spline params (..., P) for position d of `layer`."""
  if inp is None:
    return params["~"]["first"][0]  # float32, shared by all layers
  h = inp
  for m in range(len(spec.hidden_sizes)):
    lin = params[mlp_key(layer, d, m)]
    h = torch.relu(h @ lin["w"] + lin["b"])
  lin = params[out_key(layer, d)]
  return h @ lin["w"] + lin["b"]


def _cond_input(y: Tensor, c: Optional[Tensor], idx: List[int]) -> Tensor:
  # [c, y[perm[:d]]] with c broadcast over rows; no c for unconditional flows (autoregressive.py:94-98)
  cols = y[..., idx]
  if c is None:
    return cols
  c_ = c.expand(cols.shape[:-1] + c.shape[-1:])
  return torch.cat([c_, cols], dim=-1)


def layer_forward(spec, params, layer, x, c):
  """Autoregressive.forward_and_log_det: conditioners read the OUTPUT (sequential)."""
  perm = spec.permutation(layer)
  ys = [None] * spec.dim
  lds = []
  for d in range(spec.dim):
    i = perm[d]
    if d == 0:
      theta = conditioner(spec, params, layer, d, None)
    else:
      built = torch.stack([ys[j] for j in perm[:d]], dim=-1)
      theta = conditioner(spec, params, layer, d, _cond_input(built, c, list(range(d))))
    y_i, ld_i, _ = rqs.rqs_forward(x[..., i], theta)
    ys[i] = y_i
    lds.append(ld_i)
  return torch.stack(ys, dim=-1), sum(lds)


def layer_inverse(spec, params, layer, y, c):
  """Autoregressive.inverse_and_log_det: conditioners read the INPUT (parallel)."""
  perm = spec.permutation(layer)
  xs = [None] * spec.dim
  lds = []
  for d in range(spec.dim):
    i = perm[d]
    if d == 0:
      theta = conditioner(spec, params, layer, d, None)
    else:
      theta = conditioner(spec, params, layer, d, _cond_input(y, c, perm[:d]))
    x_i, ld_i, _ = rqs.rqs_inverse(y[..., i], theta)
    xs[i] = x_i
    lds.append(ld_i)
  return torch.stack(xs, dim=-1), sum(lds)


def _as_cond(c, rows_like: Tensor) -> Optional[Tensor]:
  if c is None:   # unconditional flow
    return None
  c = torch.as_tensor(c, dtype=rows_like.dtype)
  if c.dim() == 0:
    c = c.reshape(1)
  return c


def flow_forward_and_log_det(spec, params, x, c=None):
  """flow.bijector.forward: latent -> physical ("sample direction").

  ConditionalInverse.forward = chain.inverse = layers in list order, each
  layer's inverse_and_log_det (conditional.py:153-157,169-177,233-237)."""
  c = _as_cond(c, x) if spec.conditional else None
  ld = torch.zeros(x.shape[:-1], dtype=x.dtype)
  for l in range(spec.num_layers):
    x, ld_l = layer_inverse(spec, params, l, x, c)
    ld = ld + ld_l
  return x, ld


def flow_inverse_and_log_det(spec, params, y, c=None):
  """flow.bijector.inverse: physical -> latent ("log-prob direction").

  ConditionalInverse.inverse = chain.forward = reversed layers, each layer's
  forward_and_log_det (conditional.py:147-151,159-167,239-243)."""
  c = _as_cond(c, y) if spec.conditional else None
  ld = torch.zeros(y.shape[:-1], dtype=y.dtype)
  for l in reversed(range(spec.num_layers)):
    y, ld_l = layer_forward(spec, params, l, y, c)
    ld = ld + ld_l
  return y, ld


def base_log_prob(x: Tensor) -> Tensor:
  # Independent(Normal(0,1)) over the event dim (flows.py:166-173)
  return (-0.5 * x * x - 0.5 * math.log(2.0 * math.pi)).sum(-1)


def log_prob(spec, params, value, cond):
  """ConditionalTransformed.log_prob (conditional.py:316-321)."""
  x, ildj = flow_inverse_and_log_det(spec, params, value, cond)
  return base_log_prob(x) + ildj


def sample(spec, params, latent, cond):
  """ConditionalTransformed.sample with the N(0,I) draw given explicitly
  (`latent` replaces `seed`; conditional.py:376-380).  cond: (rows, 1)."""
  y, _ = flow_forward_and_log_det(spec, params, latent, cond)
  return y


def sample_and_log_prob(spec, params, latent, cond):
  """ConditionalTransformed.sample_and_log_prob (conditional.py:382-402)."""
  y, fldj = flow_forward_and_log_det(spec, params, latent, cond)
  return y, base_log_prob(latent) - fldj


def clone_params(params: Params, requires_grad: bool = False) -> Params:
  out: Params = {}
  for mod, leaves in params.items():
    out[mod] = {
      k: v.detach().clone().requires_grad_(requires_grad)
      for k, v in leaves.items()
    }
  return out


def leaves(params: Params) -> List[Tensor]:
  return [v for mod in params.values() for v in mod.values()]
