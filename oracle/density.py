"""CPU oracle: densities on grids / at samples, torch f64.

TEST INFRASTRUCTURE ONLY.  Restates what the reference computes behind its plotting helpers and in the fp evaluation
tail: `utils.plot_density_snapshot` / `plot_density_and_trajectory` (`/root/reference/cnf_ot/utils.py:572-642`: the
arrays handed to `imshow`, the trajectories handed to `scatter`), and `rmse_mc_loss_fn` / `rmse_grid_loss_fn`
(`/root/reference/cnf_ot/mfc/solvers.py:238-301`).  PARITY PINNED: `tests/golden/ref_density_d2.npz` holds what the
reference's own `utils.py` functions passed to a recording matplotlib stand-in, and the two RMSE values computed with the
reference's model API line by line (`tests/golden/make_reference_golden.py:density_case`).
"""
import math

import torch

from . import flow as oflow


def grid_points(domain, nx, ny):
  """XY = hstack(meshgrid(linspace(x0, x1, nx), linspace(y0, y1, ny))): row index iy * nx + ix (utils.py:584-588)."""
  x0, x1, y0, y1 = domain
  x = torch.linspace(x0, x1, nx, dtype=torch.float64)
  y = torch.linspace(y0, y1, ny, dtype=torch.float64)
  Y, X = torch.meshgrid(y, x, indexing="ij")
  return torch.stack([X.reshape(-1), Y.reshape(-1)], dim=1)


def density_on_grid(spec, params, t_array, domain, n):
  XY = grid_points(domain, n, n)
  out = []
  for t in t_array:
    lp = oflow.log_prob(spec, params, XY, torch.tensor([float(t)], dtype=torch.float64))
    out.append(torch.exp(lp).reshape(n, n))
  return torch.stack(out)


def trajectories(spec, params, r_, t_array):
  xi, _ = oflow.flow_inverse_and_log_det(spec, params, r_, torch.zeros(1, dtype=torch.float64))
  return torch.stack([oflow.flow_forward_and_log_det(spec, params, xi, torch.tensor([float(t)], dtype=torch.float64))[0]
                      for t in t_array])


def _iso_pdf(s, var):
  d = s.shape[-1]
  return torch.exp(-0.5 * (s * s).sum(-1) / var) / (2.0 * math.pi * var)**(d / 2)


def fp_reference(s, cond, a, T):
  """source_prob (1 - cond) + target_prob cond, solvers.py:238-252,270-276."""
  v1 = math.exp(-2 * a * T) * (4 - 1 / 2 / a) + 1 / 2 / a
  return _iso_pdf(s, 4.0) * (1 - cond) + _iso_pdf(s, v1) * cond


def rmse_grid(spec, params, cond, grid_size, a, T):
  XY = grid_points([-5, 5, -5, 5], grid_size, grid_size)
  p = torch.exp(oflow.log_prob(spec, params, XY, torch.tensor([float(cond)], dtype=torch.float64)))
  return torch.sqrt(((p - fp_reference(XY, cond, a, T))**2).mean())


def rmse_mc(spec, params, cond, latent, a, T):
  c = torch.full((latent.shape[0], 1), float(cond), dtype=torch.float64)
  samples, lp = oflow.sample_and_log_prob(spec, params, latent, c)
  return torch.sqrt(((torch.exp(lp) - fp_reference(samples, cond, a, T))**2).mean())
