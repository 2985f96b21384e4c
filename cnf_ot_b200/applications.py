"""Host-side mirror of /root/reference/cnf_ot/mfc/applications.py.

Same function names, argument order and meaning as the reference.  Two kinds of
use, like in the reference:

  * forward values of a single term (e.g. `density_fit_kl_loss_fn` for the KL
    printed at solvers.py:112-115): flow kernels (`model.apply.*`) plus a short
    device-side reduction of the per-row results;
  * the train step: `value_and_grad(loss_fn)` on a `functools.partial` of
    `ot_loss_fn` / `rwpo_loss_fn` / `fp_loss_fn` -- exactly how
    solvers.py:58-97 builds and differentiates its loss -- runs ONE fused kernel
    (`cnfot_mfc_step`) that evaluates every term and its backward pass.

Random draws: every sampler of one loss call uses the same key (as in the
reference), drawn with `cnf_ot_b200.random` (torch Philox, not jax.random).
"""
from __future__ import annotations

import functools
import math
from typing import Callable, Dict, Optional

import torch

from . import dist as _dist
from . import ops, random
from .flows import ParamTree


# ------------------------------------------------------------------ data draws
def sample_source_fn(seed, sample_shape: int, dim: int, device):
  """8-mode Gaussian mixture of kl_loss_fn (applications.py:34-71), dim == 2;
  for other dims the Gaussian variant the reference keeps commented (:28-32, ot.py:72-80).
  The component noise is the SAME z `sample_target_fn` returns for this key (:81-82)."""
  return random.ot_source(seed, (sample_shape, dim), device=device)


def sample_target_fn(seed, sample_shape: int, dim: int, device):
  """N(0, I) target (applications.py:73-79); same key => same z as the source noise."""
  return random.normal(seed, (sample_shape, dim), device=device)


def _cond_rows(n, value, device):
  return torch.full((n, 1), float(value), dtype=torch.float32, device=device)


# ------------------------------------------------------------------ single terms (forward values)
def kl_loss_fn(model, dim, T, params, cond, rng, batch_size):
  """applications.py:11-86."""
  s1 = sample_source_fn(rng, batch_size, dim, model.device)
  s2 = sample_target_fn(rng, batch_size, dim, model.device)
  samples = s1 * ((T - cond) / T) + s2 * (cond / T)
  lp = model.apply.log_prob(params, samples, cond=torch.tensor([float(cond)]))
  return -lp.double().mean()


def density_fit_kl_loss_fn(model, dim, T, params, rng, batch_size):
  """applications.py:166-173."""
  return kl_loss_fn(model, dim, T, params, 0, rng, batch_size) + \
    kl_loss_fn(model, dim, T, params, T, rng, batch_size)


def reverse_kl_loss_fn(model, dim, T, beta, params, cond, rng, batch_size):
  """applications.py:129-163."""
  samples, lp = model.apply.sample_and_log_prob(
    params, cond=_cond_rows(batch_size, cond, model.device), seed=rng, sample_shape=(batch_size, ))
  r2 = (samples.double()**2).sum(-1)

  def log_pdf(var):
    return -0.5 * r2 / var - 0.5 * dim * math.log(2 * math.pi * var)

  terms = []
  if T - cond > 0:
    terms.append(log_pdf(2.0 / beta * (T + 1)) + math.log((T - cond) / T))
  if cond > 0:
    terms.append(log_pdf(2.0 / beta) + math.log(cond / T))
  logq = torch.logsumexp(torch.stack(terms), dim=0)
  return (lp.double() - logq).mean()


def ot_reverse_kl_loss_fn(model, dim, T, params, rng, batch_size):
  """applications.py:91-126: reverse KL against N(3, I) at t = 0 plus N(0, I) at t = 1 (the reference keeps it beside
  `density_fit_kl_loss_fn`; its only call site, :390, is commented out)."""
  loss = 0.0
  for cond, mean in ((0.0, 3.0), (1.0, 0.0)):
    samples, lp = model.apply.sample_and_log_prob(
      params, cond=_cond_rows(batch_size, cond, model.device), seed=rng, sample_shape=(batch_size, ))
    d2 = ((samples.double() - mean)**2).sum(-1)
    loss = loss + (lp.double() - (-0.5 * d2 - 0.5 * dim * math.log(2 * math.pi))).mean()
  return loss


def potential_loss_fn(model, dim, a, subtype, params, cond, rng, batch_size):
  """applications.py:176-205."""
  r = model.apply.sample(params, cond=_cond_rows(batch_size, cond, model.device), seed=rng,
                         sample_shape=(batch_size, )).double()
  if subtype == "quadratic":
    return ((r**2).sum(1) / 2).mean()
  if subtype == "double_well":
    return ((torch.linalg.norm(r - a, dim=1) * torch.linalg.norm(r + a, dim=1) / 2)**2).mean()
  if subtype == "obstacle":
    return (50 * torch.exp(-(r**2).sum(1) / 2)).mean()
  return None  # the reference falls through silently (:200-205)


def _samples_at(model, params, t, rng, batch_size):
  return model.apply.sample(params, seed=rng, sample_shape=(batch_size, ),
                            cond=_cond_rows(batch_size, t, model.device)).double()


def kinetic_loss_fn(model, dim, dt, params, cond, rng, batch_size):
  """applications.py:220-242."""
  r1 = _samples_at(model, params, cond - dt / 2, rng, batch_size)
  r2 = _samples_at(model, params, cond + dt / 2, rng, batch_size)
  velocity = (r2 - r1) / dt
  return (velocity**2).mean() * dim / 2


def _fd_score(model, params, r3, cond, dx, dim):
  score = torch.zeros_like(r3)
  c = torch.tensor([float(cond)])
  for i in range(dim):
    dr = torch.zeros(1, dim, dtype=torch.float32, device=r3.device)
    dr[0, i] = dx / 2
    lp1 = model.apply.log_prob(params, r3.float() + dr, cond=c).double()
    lp2 = model.apply.log_prob(params, r3.float() - dr, cond=c).double()
    score[:, i] = (lp1 - lp2) / dx
  return score


def kinetic_with_score_loss_fn(model, dim, beta, dt, dx, params, cond, rng, batch_size):
  """applications.py:245-276."""
  r1 = _samples_at(model, params, cond - dt / 2, rng, batch_size)
  r2 = _samples_at(model, params, cond + dt / 2, rng, batch_size)
  r3 = _samples_at(model, params, cond, rng, batch_size)
  velocity = (r2 - r1) / dt + _fd_score(model, params, r3, cond, dx, dim) / beta
  return (velocity**2).mean() * dim / 2


def _truth(r3, dim, a, subtype):
  if subtype == "gradient":
    x, y = r3[:, 0], r3[:, 1]
    q = x**2 + y**2 - 4
    return a * torch.stack([-q * x, -q * y - 2 * (y - 1)], dim=1)
  if subtype == "nongradient":
    if dim % 2 != 0:
      # the reference raises for dim != 2 (:358-360); the block-diagonal extension needs even dim
      raise Exception("nongradient case is only implemented for even dim!")
    rot = torch.empty_like(r3)
    rot[:, 0::2] = -r3[:, 1::2]
    rot[:, 1::2] = r3[:, 0::2]
    return -r3 * a + rot * 0.5
  if subtype == "lorenz":
    if dim != 3:
      raise Exception("Lorenz dynamics is only defined for 3 dim!")
    _r = 9
    return torch.stack([10 * (r3[:, 1] - r3[:, 0]),
                        _r * r3[:, 0] * (28 / _r - r3[:, 2]) - r3[:, 1],
                        _r * r3[:, 0] * r3[:, 1] - r3[:, 2] * 8 / 3], dim=1)
  raise Exception(f"Unknown velocity field: {subtype}")


def flow_matching_loss_fn(model, dim, a, sigma, subtype, dt, dx, params, cond, rng, batch_size):
  """applications.py:279-374 (dt and dx are overridden to 0.01 there, :286,301)."""
  dt = dx = 0.01
  r1 = _samples_at(model, params, cond - dt / 2, rng, batch_size)
  r2 = _samples_at(model, params, cond + dt / 2, rng, batch_size)
  r3 = _samples_at(model, params, cond, rng, batch_size)
  velocity = (r2 - r1) / dt + _fd_score(model, params, r3, cond, dx, dim) * sigma
  return ((velocity - _truth(r3, dim, a, subtype))**2).mean() * dim / 2


# ------------------------------------------------------------------ full losses
def _t_batch(rng, t_batch_size, scale):
  return (random.uniform(rng, (t_batch_size, ), device="cpu").double() * scale).tolist()


def ot_loss_fn(model, dim, T, dt, t_batch_size, subtype, params, rng, _lambda, batch_size):
  """applications.py:377-402 (forward value; see value_and_grad for the train step)."""
  loss = _lambda * density_fit_kl_loss_fn(model, dim, T, params, rng, batch_size)
  for t in _t_batch(rng, t_batch_size, 1.0):
    loss = loss + kinetic_loss_fn(model, dim, dt, params, t, rng, batch_size // 32) / t_batch_size
    if subtype == "obstacle":
      loss = loss + potential_loss_fn(model, dim, 0, subtype, params, t, rng, batch_size // 32)
  return loss


def rwpo_loss_fn(model, dim, T, beta, dt, dx, t_batch_size, subtype, a, params, rng, _lambda,
                 batch_size):
  """applications.py:405-421."""
  loss = _lambda * reverse_kl_loss_fn(model, dim, T, beta, params, 0, rng, batch_size) + \
    potential_loss_fn(model, dim, a, subtype, params, T, rng, batch_size)
  for t in _t_batch(rng, t_batch_size, T):
    loss = loss + kinetic_with_score_loss_fn(model, dim, beta, dt, dx, params, t, rng,
                                             batch_size // 32) / t_batch_size * T
  return loss


def fp_loss_fn(model, dim, T, a, sigma, dt, dx, t_batch_size, subtype, params, rng, _lambda,
               batch_size):
  """applications.py:424-441 (beta = 4 hard-coded at :432)."""
  beta = 4
  loss = _lambda * reverse_kl_loss_fn(model, dim, T, beta, params, 0, rng, batch_size)
  for t in _t_batch(rng, t_batch_size, T):
    loss = loss + flow_matching_loss_fn(model, dim, a, sigma, subtype, dt, dx, params, t, rng,
                                        batch_size // 32) / t_batch_size * T
  return loss


# ------------------------------------------------------------------ the train step
def _step_config(loss_fn) -> Dict:
  """Recover the mfc.yaml-style description from the partial solvers.py:58-88 builds."""
  if not isinstance(loss_fn, functools.partial) or loss_fn.keywords:
    raise TypeError("value_and_grad expects functools.partial(ot_loss_fn | rwpo_loss_fn | fp_loss_fn, model, ...)")
  f, a = loss_fn.func, loss_fn.args
  if f is ot_loss_fn and len(a) == 6:
    model, dim, T, dt, tbs, subtype = a
    cfg = {"general": {"type": "ot", "dim": dim, "dt": dt, "dx": 0.01, "t_batch_size": tbs},
           "ot": {"subtype": subtype}}
    horizon = 1.0
    if T != 1:
      raise ValueError("ot_loss_fn is defined on [0, 1] (solvers.py:81)")
  elif f is rwpo_loss_fn and len(a) == 9:
    model, dim, T, beta, dt, dx, tbs, subtype, aa = a
    cfg = {"general": {"type": "rwpo", "dim": dim, "dt": dt, "dx": dx, "t_batch_size": tbs},
           "rwpo": {"T": T, "beta": beta, "a": aa, "pot_type": subtype}}
    horizon = float(T)
  elif f is fp_loss_fn and len(a) == 9:
    model, dim, T, aa, sigma, dt, dx, tbs, subtype = a
    cfg = {"general": {"type": "fp", "dim": dim, "dt": dt, "dx": dx, "t_batch_size": tbs},
           "fp": {"T": T, "a": aa, "sigma": sigma, "velocity_field_type": subtype}}
    horizon = float(T)
  else:
    raise TypeError("value_and_grad supports the three train losses of applications.py only")
  if dim != model.shape.dim:
    raise ValueError("loss dim and model dim differ")
  return {"model": model, "cfg": cfg, "horizon": horizon, "problem": ops.problem_desc(cfg)}


def draw_step_inputs(model, cfg, horizon, rng, batch_size, shard=None):
  """The draws one loss call makes, all from the same key (applications.py:81-82,392,416) -- the arrays the
  step kernel generates on chip for this key.  shard = (rows of the B-row terms, rows of the b-row terms) as
  slices: only that part is materialised (data-parallel ranks never draw the whole batch)."""
  from . import _lib
  dim, dev = model.shape.dim, model.device
  b = batch_size // 32
  rs, ss = (slice(0, batch_size), slice(0, b)) if shard is None else shard
  typ = cfg["general"]["type"]
  key = random.as_key(rng).value
  inputs = {"t_batch": _t_batch(rng, cfg["general"]["t_batch_size"], horizon),
            "latent_sub": ops.philox_rows(key, 0, _lib.ROWS_NORMAL, b, dim, dev, rows=ss)}
  if typ == "ot":
    inputs["src"] = ops.philox_rows(key, 0, _lib.ROWS_OT_SOURCE, batch_size, dim, dev, rows=rs)
    inputs["tgt"] = ops.philox_rows(key, 0, _lib.ROWS_NORMAL, batch_size, dim, dev, rows=rs)
  else:
    inputs["latent"] = ops.philox_rows(key, 0, _lib.ROWS_NORMAL, batch_size, dim, dev, rows=rs)
  return inputs


_peer_exchanges = {}


def peer_exchange(shape, device):
  """The exchange buffers of the fused step + all-reduce for this flow shape (one per process and shape), or
  None when the all-reduce has to go through torch.distributed: a single process, the gloo backend of the CPU
  tests, the wide-conditioner engine, symmetric memory unavailable, or CNFOT_DP_TRANSPORT=nccl."""
  import os
  import torch.distributed as td
  rank, world = _dist.rank_world()
  if world == 1 or world > 8 or torch.device(device).type != "cuda" or os.environ.get("CNFOT_DP_TRANSPORT") == "nccl":
    return None
  if td.get_backend() != "nccl" or not ops.fused_update_supported(shape):
    return None
  k = (shape, torch.device(device).index)
  if k not in _peer_exchanges:
    try:
      _peer_exchanges[k] = _dist.PeerExchange(shape, torch.device(device))
    except Exception:   # symmetric memory unavailable: NCCL all-reduce
      _peer_exchanges[k] = None
  return _peer_exchanges[k]


def value_and_grad(loss_fn: Callable):
  """jax.value_and_grad(loss_fn) of solvers.py:94 for the three MFC losses.

  Returns f(params, rng, _lambda, batch_size) -> (loss, grads) with grads a ParamTree.  The draws of the call are
  made inside the step kernel from `rng` (the same numbers `draw_step_inputs` returns).  With torch.distributed
  initialised every rank evaluates ITS rows only and the [gradient | loss] buffer is summed with ONE all-reduce
  (SURVEY.md §8e): inside the step kernel over peer-mapped memory when the GPUs share a node, else by
  torch.distributed."""
  sc = _step_config(loss_fn)
  model, cfg, problem = sc["model"], sc["cfg"], sc["problem"]

  def fn(params, rng, _lambda, batch_size, inputs: Optional[Dict] = None):
    if not isinstance(params, ParamTree):
      raise TypeError("params must be the ParamTree returned by model.init / update")
    rank, world = _dist.rank_world()
    B, b = batch_size, batch_size // 32
    rs, ss = _dist.shard(B, rank, world), _dist.shard(b, rank, world)
    n = model.shape.blob_size
    if inputs is None and ops.fused_update_supported(model.shape):
      px = peer_exchange(model.shape, model.device)
      out = ops.mfc_step_rng(model.shape, problem, params.blob, random.as_key(rng).value, 0,
                             cfg["general"]["t_batch_size"], float(_lambda), B, b, rows_B=rs, rows_b=ss, peers=px)
      if px is None:
        _dist.all_reduce_sum(out)
      return out[n], ParamTree(model.shape, out[:n])
    if inputs is None:   # the wide-conditioner engine takes arrays: this rank's shard of the same draws
      inputs = draw_step_inputs(model, cfg, sc["horizon"], rng, batch_size, shard=(rs, ss))
      rs, ss = slice(None), slice(None)
    g = lambda k: None if inputs.get(k) is None else inputs[k]
    out = ops.mfc_step(model.shape, problem, params.blob,
                       None if g("latent") is None else g("latent")[rs],
                       g("latent_sub")[ss],
                       None if g("src") is None else g("src")[rs],
                       None if g("tgt") is None else g("tgt")[rs],
                       inputs["t_batch"], float(_lambda), B, b)
    _dist.all_reduce_sum(out)
    return out[n], ParamTree(model.shape, out[:n])

  fn.step_config = sc
  return fn
