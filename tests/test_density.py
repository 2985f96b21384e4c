"""Densities on grids / at Monte-Carlo samples.

CPU: oracle/density.py reproduces what the reference's own `utils.plot_density_snapshot` /
`plot_density_and_trajectory` (cnf_ot/utils.py:572-642) handed to a recording matplotlib stand-in, and the fp L2
errors of solvers.py:254-301 (tests/golden/ref_density_d2.npz, generator make_reference_golden.py:density_case).
GPU: cnfot_density_grid / cnfot_density_mc (one launch each, grid and latent generated on chip) against the same
fixture and against the oracle on the exported draws."""
import os

import numpy as np
import pytest
import torch

from cnf_ot_b200.layout import FlowShape, unpack
from oracle import density as odens
from oracle import flow as oflow

HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
  z = np.load(os.path.join(HERE, "golden", "ref_density_d2.npz"))
  g = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}
  shape = FlowShape(*(int(v) for v in g["shape"]))
  spec = oflow.FlowSpec(shape.dim, shape.num_layers, [shape.hidden] * shape.mlp_layers, shape.num_bins)
  params = unpack(shape, g["blob"], like=oflow.init_params(spec, seed=0))   # dtypes of the reference: float64, `first` float32
  return g, shape, spec, params


def test_oracle_matches_the_reference_plots_and_rmse():
  g, shape, spec, params = _load()
  k = int(g["stride"])
  with torch.no_grad():
    snap = odens.density_on_grid(spec, params, np.linspace(0, 1, 10), [-6, 6, -6, 6], 100)
    assert float((snap[:, ::k, ::k] - g["snap_sub"]).abs().max()) < 1e-12
    assert float((snap.sum((1, 2)) - g["snap_sum"]).abs().max()) < 1e-9
    dens2 = odens.density_on_grid(spec, params, g["t_traj"].tolist(), g["domain2"].tolist(), 100)
    assert float((dens2[:, ::k, ::k] - g["dens2_sub"]).abs().max()) < 1e-12
    traj = odens.trajectories(spec, params, g["r"], g["t_traj"].tolist())
    assert float((traj - g["traj"]).abs().max()) < 1e-10
    a, T = float(g["fp_a"]), float(g["fp_T"])
    assert abs(float(odens.rmse_mc(spec, params, 1.0, g["latent_mc"], a, T)) - float(g["rmse_mc"])) < 1e-12
    assert abs(float(odens.rmse_grid(spec, params, 1.0, int(g["grid_size"]), a, T)) - float(g["rmse_grid"])) < 1e-12


@pytest.mark.gpu
def test_gpu_density_grid_and_mc_match_the_reference(engine):
  from cnf_ot_b200 import _lib, ops, random, utils
  from cnf_ot_b200.flows import FlowModel, ParamTree
  g, shape, spec, params = _load()
  k = int(g["stride"])
  model = FlowModel(shape, "cuda")
  P = ParamTree(shape, g["blob"].float().cuda())
  close = lambda a, b, tol: float(((a.double().cpu() - b).abs() / (b.abs() + 1e-2)).max()) < tol
  # the ten snapshots of plot_density_snapshot, one launch
  snap = utils.plot_density_snapshot(model.apply.log_prob, P)
  assert _lib.last_launch_info()["engine"] == engine
  assert snap.shape == (10, 100, 100) and close(snap[:, ::k, ::k], g["snap_sub"], 2e-5)
  assert float(((snap.double().sum((1, 2)).cpu() - g["snap_sum"]) / g["snap_sum"]).abs().max()) < 1e-5
  # plot_density_and_trajectory: densities on a rectangular domain + trajectories
  dens2 = utils.density_on_grid(model.apply.log_prob, P, g["t_traj"].tolist(), g["domain2"].tolist(), 100)
  assert close(dens2[:, ::k, ::k], g["dens2_sub"], 2e-5)
  traj = utils.trajectories(model.apply.forward, model.apply.inverse, P, g["r"].float().cuda(), g["t_traj"].tolist())
  assert float(((traj.double().cpu() - g["traj"]).abs() / (g["traj"].abs() + 1)).max()) < 2e-5
  # the grid densities equal log_prob on an explicit XY array (the kernel generates the grid itself)
  XY = odens.grid_points([-6, 6, -6, 6], 100, 100).float().cuda()
  lp = model.apply.log_prob(P, XY, cond=torch.tensor([float(np.linspace(0, 1, 10)[3])]))
  assert float((torch.exp(lp).reshape(100, 100) - snap[3]).abs().max()) < 2e-6
  # rmse_grid_loss_fn (solvers.py:282-301)
  a, T = float(g["fp_a"]), float(g["fp_T"])
  rg = float(utils.rmse_grid_loss_fn(model.apply.log_prob, P, 1.0, int(g["grid_size"]), a=a, T=T))
  assert abs(rg - float(g["rmse_grid"])) < 2e-5 * float(g["rmse_grid"])
  # rmse_mc_loss_fn (solvers.py:254-278): latent drawn on chip; the oracle gets the exported rows
  key = random.PRNGKey(11)
  n = 50000
  rm = float(utils.rmse_mc_loss_fn(model, P, 1.0, key, n, a=a, T=T))
  lat = ops.philox_rows(key.value, 0, _lib.ROWS_NORMAL, n, 2, "cuda").double().cpu()
  with torch.no_grad():
    want = float(odens.rmse_mc(spec, params, 1.0, lat, a, T))
  assert abs(rm - want) < 2e-5 * want
  smp, dens, _ = ops.density_mc(shape, P.blob, 1.0, key.value, 4096, want_samples=True, want_density=True)
  y, lp = model.apply.sample_and_log_prob(P, cond=torch.ones(4096, 1, device="cuda"), seed=key, sample_shape=(4096, ))
  assert float((smp - y).abs().max()) < 1e-5 and float((dens - torch.exp(lp)).abs().max()) < 1e-6
