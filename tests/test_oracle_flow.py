"""Pins the flow/loss oracle by the invariants the reference documents
(SURVEY.md §4): parameter pytree shape/count, identity at init, exact
inverse pair, log-det vs autodiff Jacobian, autograd vs finite differences."""
import math

import pytest
import torch

from oracle import flow as oflow
from oracle import losses as olosses

CFG = {
  "general": {"type": "ot", "dim": 2, "dx": 0.01, "dt": 0.01, "t_batch_size": 1, "seed": 42},
  "ot": {"subtype": "obstacle"},
  "rwpo": {"T": 1, "beta": 1, "a": 1, "pot_type": "double_well"},
  "fp": {"T": 1, "a": 1, "sigma": 0.5, "velocity_field_type": "nongradient"},
  "cnf": {"flow_num_layers": 2, "mlp_num_layers": 2, "hidden_size": 16, "num_bins": 5},
  "train": {"epochs": 1, "lr": 1e-3, "_lambda": 5000.0, "batch_size": 256, "eval_frequency": 100},
}


def _spec(dim=2, L=2, H=16, M=2, K=5):
  return oflow.FlowSpec(dim, L, [H] * M, K)


def test_param_tree_matches_reference_count():
  spec = _spec()
  params = oflow.init_params(spec)
  n = sum(v.numel() for v in oflow.leaves(params))
  assert n == 1200 == spec.param_count()  # mfc.yaml defaults, SURVEY A.3
  assert params["~"]["first"].dtype == torch.float32
  assert params["mlp_layer0_d1/~/linear_0"]["w"].shape == (2, 16)
  assert params["mlp_layer1_d1/~/linear_1"]["w"].shape == (16, 16)
  assert params["linear_out_layer1_d1"]["w"].shape == (16, 16)
  assert _spec(dim=10).param_count() == 11824


def test_identity_at_init():
  spec = _spec(dim=3)
  params = oflow.init_params(spec, seed=1)
  gen = torch.Generator().manual_seed(0)
  x = torch.randn(64, 3, generator=gen, dtype=torch.float64)
  c = torch.full((64, 1), 0.3, dtype=torch.float64)
  y = oflow.sample(spec, params, x, c)
  assert (y - x).abs().max() < 1e-6  # `first` is float32 in the reference
  lp = oflow.log_prob(spec, params, x, torch.tensor([0.3], dtype=torch.float64))
  assert (lp - oflow.base_log_prob(x)).abs().max() < 1e-6


@pytest.mark.parametrize("dim,L,sigma", [(2, 2, 0.3), (3, 3, 0.2), (10, 2, 0.05)])
def test_inverse_pair_and_logdet(dim, L, sigma):
  spec = _spec(dim=dim, L=L)
  params = oflow.perturb_params(oflow.init_params(spec, seed=2), sigma)
  gen = torch.Generator().manual_seed(5)
  x = torch.randn(128, dim, generator=gen, dtype=torch.float64)
  c = torch.tensor([0.7], dtype=torch.float64)
  y, fld = oflow.flow_forward_and_log_det(spec, params, x, c)
  xr, ild = oflow.flow_inverse_and_log_det(spec, params, y, c)
  assert (xr - x).abs().max() < 1e-9
  assert (fld + ild).abs().max() < 1e-9
  # log-det equals log|det J| of the sample-direction map (one row)
  jac = torch.autograd.functional.jacobian(
    lambda v: oflow.flow_forward_and_log_det(spec, params, v[None], c)[0][0], x[0]
  )
  assert abs(fld[0] - torch.log(torch.abs(torch.linalg.det(jac)))) < 1e-8


def _inputs(cfg, gen):
  B, D = cfg["train"]["batch_size"], cfg["general"]["dim"]
  if D == 2:
    src, tgt = olosses.source_mixture(gen, B, D)
  else:
    src, tgt = olosses.source_gaussian(gen, B, D)
  return {
    "latent": torch.randn(B, D, generator=gen, dtype=torch.float64),
    "src": src, "tgt": tgt,
    "t_batch": torch.rand(cfg["general"]["t_batch_size"], generator=gen, dtype=torch.float64),
  }


@pytest.mark.parametrize("typ", ["ot", "rwpo", "fp"])
def test_grad_matches_finite_differences(typ):
  cfg = {k: dict(v) for k, v in CFG.items()}
  cfg["general"]["type"] = typ
  cfg["train"]["_lambda"] = 50.0
  spec = olosses.spec_from_config(cfg)
  params = oflow.perturb_params(oflow.init_params(spec, seed=3), 0.3)
  params["~"]["first"] = params["~"]["first"].to(torch.float64)  # FD needs f64 leaves
  inputs = _inputs(cfg, torch.Generator().manual_seed(42))
  loss, grads = olosses.value_and_grad(cfg, spec, params, inputs)
  assert torch.isfinite(loss)
  gen = torch.Generator().manual_seed(9)
  for mod in ["~", "mlp_layer0_d1/~/linear_0", "mlp_layer1_d1/~/linear_1", "linear_out_layer0_d1"]:
    for leaf in params[mod]:
      v = params[mod][leaf]
      flat = int(torch.randint(0, v.numel(), (1, ), generator=gen))
      eps = 1e-6
      vals = []
      for sgn in (+1, -1):
        p2 = oflow.clone_params(params)
        p2[mod][leaf].view(-1)[flat] += sgn * eps
        vals.append(olosses.loss_from_config(cfg, spec, p2, inputs))
      fd = (vals[0] - vals[1]) / (2 * eps)
      an = grads[mod][leaf].reshape(-1)[flat]
      assert abs(fd - an) <= 1e-5 * max(1.0, abs(an)), (mod, leaf, float(fd), float(an))


def test_rwpo_quadratic_reference_point():
  """At the identity flow the rKL term at cond=0 is E[log N(x;0,I) - log N(x;0,2(T+1)/beta I)]
  (closed form), an anchor for the sign/normalisation conventions."""
  spec = _spec()
  params = oflow.init_params(spec)
  gen = torch.Generator().manual_seed(0)
  lat = torch.randn(20000, 2, generator=gen, dtype=torch.float64)
  T, beta = 1.0, 1.0
  var = 2.0 / beta * (T + 1)
  val = olosses.reverse_kl_loss(spec, params, lat, 0.0, T, beta)
  m2 = (lat * lat).sum(-1).mean()
  expect = -0.5 * m2 + 0.5 * m2 / var + math.log(var)  # D=2: (D/2) log var
  assert abs(val - expect) < 1e-5


def test_drift_extension_reduces_to_reference_at_2d():
  r = torch.tensor([[1.0, 2.0], [-0.5, 0.25]], dtype=torch.float64)
  J = torch.tensor([[0.0, 1.0], [-1.0, 0.0]], dtype=torch.float64)
  expect = -r * 1.5 + (r @ J) * 0.5  # applications.py:361-363
  assert torch.allclose(olosses.drift(r, "nongradient", 1.5), expect)
