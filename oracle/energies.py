"""CPU oracle: the evaluation energies, torch f64.

TEST INFRASTRUCTURE ONLY (see `oracle/rqs.py` header).  PARITY PINNED to the reference's
own `cnf_ot/utils.py` run unmodified here (`tests/golden/ref_energy_*.npz`, 1e-12; see `oracle/flow.py`).

Restates `calc_kinetic_energy` / `calc_score_kinetic_energy` of
`/root/reference/cnf_ot/utils.py:311-389` with the per-time latent batches made an explicit input
(`latent[i]` is the N(0, I) batch the reference draws for time `t_array[i]`).
"""
import torch

from . import flow as oflow


def _samples(spec, params, latent, t):
  return oflow.sample(spec, params, latent, torch.full((latent.shape[0], 1), float(t), dtype=torch.float64))


def kinetic_energy(spec, params, latent, t_array, dt=0.01):
  """utils.py:311-340.  latent: (n_t, batch, D)."""
  dim = latent.shape[-1]
  e = 0.0
  for i, t in enumerate(t_array):
    r1 = _samples(spec, params, latent[i], t - dt / 2)
    r2 = _samples(spec, params, latent[i], t + dt / 2)
    v = (r2 - r1) / dt
    e = e + (v**2).mean() / 2
  return e / len(t_array) * dim


def score_kinetic_energy(spec, params, latent, t_array, beta, dt=0.01, dx=0.01):
  """utils.py:343-389.  latent: (n_t, batch, D)."""
  dim = latent.shape[-1]
  e = 0.0
  for i, t in enumerate(t_array):
    r1 = _samples(spec, params, latent[i], t - dt / 2)
    r2 = _samples(spec, params, latent[i], t + dt / 2)
    r3 = _samples(spec, params, latent[i], t)
    v = (r2 - r1) / dt
    c = torch.tensor([float(t)], dtype=torch.float64)
    score = torch.zeros_like(r3)
    for j in range(dim):
      dr = torch.zeros(1, dim, dtype=torch.float64)
      dr[0, j] = dx / 2
      score[:, j] = (oflow.log_prob(spec, params, r3 + dr, c) - oflow.log_prob(spec, params, r3 - dr, c)) / dx
    v = v + score / beta
    e = e + (v**2).mean() / 2
  return e / len(t_array) * dim
