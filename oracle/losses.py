"""CPU oracle: the MFC loss terms and the train step's value-and-grad, torch f64.

TEST INFRASTRUCTURE ONLY (see `oracle/rqs.py` header).  PARITY PINNED to the
reference's own `cnf_ot/mfc/applications.py` run unmodified in this container on torch-f64 stand-ins for
jax / haiku / distrax (`tests/golden/ref_step_*.npz`, generator `tests/golden/make_reference_golden.py`):
all three losses, seven sub-types, loss 1e-15 and gradient 5e-15 apart (`tests/test_reference_golden.py`
asserts 1e-12; the float32 `first` leaf 2e-6).  The spline underneath is third-party: see `oracle/rqs.py`.

Restates `/root/reference/cnf_ot/mfc/applications.py` with the random draws
made explicit inputs (the reference passes one PRNG key `rng` to every sampler
in a loss call, so every flow evaluation of a call sees the SAME latent rows;
here `latent` is that array and the b = B//32 sub-batch terms use its first b
rows; `src`/`tgt` are the data batches `kl_loss_fn` draws; `t_batch` the
uniform times drawn at `applications.py:392,416,435`).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import flow as oflow

Tensor = torch.Tensor


# ---------------------------------------------------------------- data draws
def source_mixture(gen: torch.Generator, n: int, dim: int):
  """8-mode Gaussian mixture on a circle of radius 5 (applications.py:34-71).

  The reference reuses one key for the component noise and for the target
  N(0,I) draw (applications.py:81-82), so source = z + centre[idx] and
  target = z share the same z.  Returns (source, target)."""
  if dim != 2:
    raise ValueError("the mixture source is defined for dim == 2 only")
  R = 5.0
  centres = torch.tensor(
    [[0.0, R], [R, 0.0], [0.0, -R], [-R, 0.0], [0.6 * R, 0.8 * R],
     [0.6 * R, -0.8 * R], [-0.6 * R, -0.8 * R], [-0.6 * R, 0.8 * R]],
    dtype=torch.float64
  )
  z = torch.randn(n, dim, generator=gen, dtype=torch.float64)
  idx = torch.randint(0, 8, (n, ), generator=gen)
  return z + centres[idx], z


def source_gaussian(gen: torch.Generator, n: int, dim: int, shift: float = -3.0):
  """Gaussian -> Gaussian variant (commented source at applications.py:28-32,
  legacy ot.py:72-80): source N(shift*1, I), target N(0, I), same z."""
  z = torch.randn(n, dim, generator=gen, dtype=torch.float64)
  return z + shift, z


# ---------------------------------------------------------------- loss terms
def kl_loss(spec, params, samples: Tensor, cond: float) -> Tensor:
  """kl_loss_fn (applications.py:11-86) on already-mixed samples."""
  c = torch.tensor([cond], dtype=torch.float64)
  return -oflow.log_prob(spec, params, samples, c).mean()


def density_fit_kl_loss(spec, params, src: Tensor, tgt: Tensor, T: float):
  """density_fit_kl_loss_fn (applications.py:166-173): KL at t=0 and t=T."""
  tot = 0.0
  for cond in (0.0, float(T)):
    samples = src * (T - cond) / T + tgt * cond / T
    tot = tot + kl_loss(spec, params, samples, cond)
  return tot


def _iso_normal_pdf(s: Tensor, var: float) -> Tensor:
  d = s.shape[-1]
  return torch.exp(-0.5 * (s * s).sum(-1) / var) / (2.0 * math.pi * var)**(d / 2)


def reverse_kl_loss(spec, params, latent, cond: float, T: float, beta: float):
  """reverse_kl_loss_fn (applications.py:129-163)."""
  c = torch.full((latent.shape[0], 1), float(cond), dtype=torch.float64)
  samples, lp = oflow.sample_and_log_prob(spec, params, latent, c)
  src = _iso_normal_pdf(samples, 2.0 / beta * (T + 1))
  tgt = _iso_normal_pdf(samples, 2.0 / beta)
  return (lp - torch.log(src * (T - cond) / T + tgt * cond / T)).mean()


def potential_value(r: Tensor, subtype: str, a: float) -> Tensor:
  if subtype == "quadratic":
    return (r * r).sum(-1) / 2
  if subtype == "double_well":
    n1 = torch.linalg.norm(r - a, dim=-1)
    n2 = torch.linalg.norm(r + a, dim=-1)
    return (n1 * n2 / 2)**2
  if subtype == "obstacle":
    return 50.0 * torch.exp(-(r * r).sum(-1) / 2)
  raise ValueError(f"unknown potential {subtype}")


def potential_loss(spec, params, latent, cond: float, subtype: str, a: float):
  """potential_loss_fn (applications.py:176-205)."""
  c = torch.full((latent.shape[0], 1), float(cond), dtype=torch.float64)
  return potential_value(oflow.sample(spec, params, latent, c), subtype, a).mean()


def _samples_at(spec, params, latent, t: float):
  c = torch.full((latent.shape[0], 1), float(t), dtype=torch.float64)
  return oflow.sample(spec, params, latent, c)


def kinetic_loss(spec, params, latent, t: float, dt: float):
  """kinetic_loss_fn (applications.py:220-242)."""
  d = latent.shape[-1]
  r1 = _samples_at(spec, params, latent, t - dt / 2)
  r2 = _samples_at(spec, params, latent, t + dt / 2)
  v = (r2 - r1) / dt
  return (v * v).mean() * d / 2


def _fd_score(spec, params, r3: Tensor, t: float, dx: float) -> Tensor:
  d = r3.shape[-1]
  c = torch.tensor([t], dtype=torch.float64)
  cols = []
  for i in range(d):
    dr = torch.zeros(1, d, dtype=torch.float64)
    dr[0, i] = dx / 2
    lp1 = oflow.log_prob(spec, params, r3 + dr, c)
    lp2 = oflow.log_prob(spec, params, r3 - dr, c)
    cols.append((lp1 - lp2) / dx)
  return torch.stack(cols, dim=-1)


def kinetic_with_score_loss(spec, params, latent, t, beta, dt, dx):
  """kinetic_with_score_loss_fn (applications.py:245-276)."""
  d = latent.shape[-1]
  r1 = _samples_at(spec, params, latent, t - dt / 2)
  r2 = _samples_at(spec, params, latent, t + dt / 2)
  r3 = _samples_at(spec, params, latent, t)
  v = (r2 - r1) / dt + _fd_score(spec, params, r3, t, dx) / beta
  return (v * v).mean() * d / 2


def drift(r3: Tensor, subtype: str, a: float) -> Tensor:
  """`truth` of flow_matching_loss_fn (applications.py:308-372).

  gradient: the 2-D "smiling" drift; lorenz: 3-D; nongradient: the reference
  raises for dim != 2 -- here extended block-diagonally for even dim
  (J_D = I_{D/2} (x) [[0,1],[-1,0]]), identical at dim == 2."""
  d = r3.shape[-1]
  if subtype == "gradient":
    if d != 2:
      raise ValueError("gradient drift is defined for dim == 2 only")
    x, y = r3[:, 0], r3[:, 1]
    q = x * x + y * y - 4.0
    return a * torch.stack([-q * x, -q * y - 2.0 * (y - 1.0)], dim=-1)
  if subtype == "nongradient":
    if d % 2 != 0:
      raise ValueError("nongradient drift needs an even dim")
    rot = torch.empty_like(r3)
    rot[:, 0::2] = -r3[:, 1::2]
    rot[:, 1::2] = r3[:, 0::2]
    return -a * r3 + 0.5 * rot
  if subtype == "lorenz":
    if d != 3:
      raise ValueError("Lorenz dynamics is only defined for 3 dim")
    s = 9.0
    x, y, z = r3[:, 0], r3[:, 1], r3[:, 2]
    return torch.stack(
      [10.0 * (y - x), s * x * (28.0 / s - z) - y, s * x * y - z * 8.0 / 3.0],
      dim=-1
    )
  raise ValueError(f"unknown velocity field {subtype}")


def flow_matching_loss(spec, params, latent, t, a, sigma, subtype):
  """flow_matching_loss_fn (applications.py:279-374); dt = dx = 0.01 are
  hard-coded there (:286,301) regardless of the arguments."""
  dt = dx = 0.01
  d = latent.shape[-1]
  r1 = _samples_at(spec, params, latent, t - dt / 2)
  r2 = _samples_at(spec, params, latent, t + dt / 2)
  r3 = _samples_at(spec, params, latent, t)
  v = (r2 - r1) / dt + _fd_score(spec, params, r3, t, dx) * sigma
  return ((v - drift(r3, subtype, a))**2).mean() * d / 2


# ---------------------------------------------------------------- full losses
def _sub_batch(inputs):
  """Latent rows of the B//32 sub-batch terms: an explicit `latent_sub` draw (the
  reference draws a fresh (b, D) normal from the same key) or the first b rows."""
  if inputs.get("latent_sub") is not None:
    return inputs["latent_sub"]
  lat = inputs["latent"]
  return lat[:lat.shape[0] // 32]


def ot_loss(spec, params, inputs, lam, *, T, dt, subtype):
  """ot_loss_fn (applications.py:377-402)."""
  tb = inputs["t_batch"]
  sub = _sub_batch(inputs)
  loss = lam * density_fit_kl_loss(spec, params, inputs["src"], inputs["tgt"], T)
  for t in tb.tolist():
    loss = loss + kinetic_loss(spec, params, sub, t, dt) / len(tb)
    if subtype == "obstacle":
      loss = loss + potential_loss(spec, params, sub, t, "obstacle", 0.0)
  return loss


def rwpo_loss(spec, params, inputs, lam, *, T, beta, dt, dx, subtype, a):
  """rwpo_loss_fn (applications.py:405-421)."""
  tb = inputs["t_batch"]
  lat = inputs["latent"]
  sub = _sub_batch(inputs)
  loss = lam * reverse_kl_loss(spec, params, lat, 0.0, T, beta)
  loss = loss + potential_loss(spec, params, lat, T, subtype, a)
  for t in tb.tolist():
    loss = loss + kinetic_with_score_loss(
      spec, params, sub, t, beta, dt, dx
    ) / len(tb) * T
  return loss


def fp_loss(spec, params, inputs, lam, *, T, a, sigma, subtype):
  """fp_loss_fn (applications.py:424-441); beta = 4 is hard-coded (:432)."""
  tb = inputs["t_batch"]
  lat = inputs["latent"]
  sub = _sub_batch(inputs)
  loss = lam * reverse_kl_loss(spec, params, lat, 0.0, T, 4.0)
  for t in tb.tolist():
    loss = loss + flow_matching_loss(
      spec, params, sub, t, a, sigma, subtype
    ) / len(tb) * T
  return loss


def loss_from_config(cfg: Dict, spec, params, inputs, lam=None) -> Tensor:
  """Dispatch on `general.type` exactly like solvers.py:58-88."""
  g = cfg["general"]
  lam = cfg["train"]["_lambda"] if lam is None else lam
  if g["type"] == "ot":
    return ot_loss(spec, params, inputs, lam, T=1.0, dt=g["dt"],
                   subtype=cfg["ot"]["subtype"])
  if g["type"] == "rwpo":
    r = cfg["rwpo"]
    return rwpo_loss(spec, params, inputs, lam, T=float(r["T"]),
                     beta=float(r["beta"]), dt=g["dt"], dx=g["dx"],
                     subtype=r["pot_type"], a=float(r["a"]))
  if g["type"] == "fp":
    f = cfg["fp"]
    return fp_loss(spec, params, inputs, lam, T=float(f["T"]), a=float(f["a"]),
                   sigma=float(f["sigma"]), subtype=f["velocity_field_type"])
  raise ValueError(f"Unknown problem type: {g['type']}...")


def spec_from_config(cfg: Dict) -> oflow.FlowSpec:
  c = cfg["cnf"]
  return oflow.FlowSpec(cfg["general"]["dim"], c["flow_num_layers"],
                        [c["hidden_size"]] * c["mlp_num_layers"], c["num_bins"])


def value_and_grad(cfg: Dict, spec, params, inputs, lam=None):
  """update()'s `jax.value_and_grad(loss_fn)` (solvers.py:94), optimiser excluded."""
  p = oflow.clone_params(params, requires_grad=True)
  loss = loss_from_config(cfg, spec, p, inputs, lam)
  loss.backward()
  grads = {
    mod: {k: (v.grad if v.grad is not None else torch.zeros_like(v))
          for k, v in lv.items()}
    for mod, lv in p.items()
  }
  return loss.detach(), grads
