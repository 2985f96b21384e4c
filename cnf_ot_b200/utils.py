"""Evaluation energies of the MFC solvers -- the callers right after the train step.

Mirrors `calc_kinetic_energy` / `calc_score_kinetic_energy` of
/root/reference/cnf_ot/utils.py:311-389 (called from cnf_ot/mfc/solvers.py:141-164): same
names, arguments and return values, but each is ONE forward-only CUDA kernel over a chunk of the
time grid (cnfot_kinetic_energy) instead of 2-(3+2 dim) flow evaluations per time in Python.
As in the reference every time of the grid gets a fresh N(0, I) batch (`rng` is split per time);
equal `rng` => equal result.
"""
from __future__ import annotations

import torch

from . import ops, random
from .flows import _blob_of

_T_CHUNK = 128  # times per launch: bounds the latent buffer at _T_CHUNK * batch_size rows


def _model_of(sample_fn):
  model = getattr(sample_fn, "__self__", None)
  if model is None or not hasattr(model, "shape"):
    raise TypeError("sample_fn must be `model.apply.sample` of a cnf_ot_b200 flow model")
  return model


def _energy(model, params, rng, batch_size, t_array, with_score, kappa):
  W = _blob_of(model.shape, params, model.device)
  total = torch.zeros((), dtype=torch.float64, device=model.device)
  n_t = len(t_array)
  rng = random.as_key(rng)
  for lo in range(0, n_t, _T_CHUNK):
    ts = t_array[lo:lo + _T_CHUNK]
    _rng, rng = random.split(rng)
    latent = random.normal(_rng, (len(ts) * batch_size, model.shape.dim), device=model.device)
    e = ops.kinetic_energy(model.shape, W, latent, ts, dt=0.01, with_score=with_score, kappa=kappa, dx=0.01,
                           latent_blocks=len(ts))
    total += e * (len(ts) / n_t)
  return total


def calc_kinetic_energy(sample_fn, params, rng, batch_size: int = 65536, t_size: int = 10000, dim: int = 1):
  """Monte-Carlo kinetic energy, utils.py:311-340: mean over t in linspace(0, 1, t_size) of
  mean(((r(t+dt/2) - r(t-dt/2)) / dt)^2) / 2 * dim, dt = 0.01."""
  model = _model_of(sample_fn)
  if dim != model.shape.dim:
    raise ValueError(f"dim={dim} does not match the flow's dimension {model.shape.dim}")
  t_array = torch.linspace(0, 1, t_size, dtype=torch.float64).tolist()
  return _energy(model, params, rng, batch_size, t_array, False, 0.0)


def calc_score_kinetic_energy(sample_fn, log_prob_fn, params, T: float = 1, beta: float = 1, dim: int = 1,
                              rng=random.PRNGKey(0), batch_size: int = 65536, t_size: int = 10000):
  """Kinetic energy with the score-corrected velocity, utils.py:343-389 (dx = dt = 0.01,
  velocity += score / beta); t in linspace(0, T, t_size).  `log_prob_fn` is accepted for signature
  parity; the kernel evaluates the same model's log-density."""
  model = _model_of(sample_fn)
  if getattr(log_prob_fn, "__self__", model) is not model:
    raise TypeError("sample_fn and log_prob_fn must belong to the same model")
  if dim != model.shape.dim:
    raise ValueError(f"dim={dim} does not match the flow's dimension {model.shape.dim}")
  t_array = torch.linspace(0, T, t_size, dtype=torch.float64).tolist()
  return _energy(model, params, rng, batch_size, t_array, True, 1.0 / beta)
