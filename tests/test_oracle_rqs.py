"""Pins the oracle's spline against the invariants of the reference's own test
(/root/reference/tests/test_rqs_accuracy.py): same three configurations, same
parameter distribution (:174-210), same four checks and the same 1e-12 bound."""
import math

import pytest
import torch

from oracle import rqs

CONFIGS = [  # test_rqs_accuracy.py:28-53
  dict(K=10, batch=100, feat=2, lo=-5.0, hi=5.0, tr=(-4.0, 4.0)),
  dict(K=5, batch=50, feat=1, lo=-3.0, hi=3.0, tr=(-2.5, 2.5)),
  dict(K=20, batch=200, feat=3, lo=-4.0, hi=4.0, tr=(-3.5, 3.5)),
]


def _gen_params(gen, batch, feat, K):
  w = torch.rand(batch, feat, K, generator=gen, dtype=torch.float64) * 1.9 + 0.1
  w = w / w.sum(-1, keepdim=True)
  h = torch.rand(batch, feat, K, generator=gen, dtype=torch.float64) * 1.9 + 0.1
  h = h / h.sum(-1, keepdim=True)
  s = torch.rand(batch, feat, K + 1, generator=gen, dtype=torch.float64) * 1.5 + 0.5
  return torch.cat([w, h, s], dim=-1)


def _uniform(gen, shape, lo, hi):
  return torch.rand(shape, generator=gen, dtype=torch.float64) * (hi - lo) + lo


@pytest.mark.parametrize("cfg", CONFIGS)
def test_reference_invariants(cfg):
  gen = torch.Generator().manual_seed(42)
  kw = dict(range_min=cfg["lo"], range_max=cfg["hi"], min_knot_slope=1e-3)
  params = _gen_params(gen, cfg["batch"], cfg["feat"], cfg["K"])
  shape = (cfg["batch"], cfg["feat"])
  # 1: inverse(forward(x)) == x
  x = _uniform(gen, shape, *cfg["tr"])
  y, _, _ = rqs.rqs_forward(x, params, **kw)
  xr, _, _ = rqs.rqs_inverse(y, params, **kw)
  assert (xr - x).abs().max() < 1e-12
  # 2: forward(inverse(y)) == y
  yt = _uniform(gen, shape, *cfg["tr"])
  xi, _, _ = rqs.rqs_inverse(yt, params, **kw)
  yr, _, _ = rqs.rqs_forward(xi, params, **kw)
  assert (yr - yt).abs().max() < 1e-12
  # 3: sum(logdet) == log|det jacobian(forward)| (diagonal map)
  xj = _uniform(gen, (cfg["feat"], ), cfg["tr"][0] * 0.5, cfg["tr"][1] * 0.5)
  p0 = params[0]
  _, ld, _ = rqs.rqs_forward(xj, p0, **kw)
  jac = torch.autograd.functional.jacobian(
    lambda v: rqs.rqs_forward(v, p0, **kw)[0], xj
  )
  assert abs(ld.sum() - torch.log(torch.abs(torch.linalg.det(jac)))) < 1e-12
  # inverse log-det is the negative of the forward one at the image point
  yj, ldf, _ = rqs.rqs_forward(xj, p0, **kw)
  _, ldi, _ = rqs.rqs_inverse(yj, p0, **kw)
  assert (ldf + ldi).abs().max() < 1e-12
  # 4: boundary behaviour
  eps = 1e-6
  pts = torch.tensor(
    [[cfg["lo"] + eps] * cfg["feat"], [cfg["hi"] - eps] * cfg["feat"],
     [0.0] * cfg["feat"], [cfg["tr"][0] * 0.5] * cfg["feat"],
     [cfg["tr"][1] * 0.5] * cfg["feat"]], dtype=torch.float64
  )
  yb, _, _ = rqs.rqs_forward(pts, params[:5], **kw)
  xb, _, _ = rqs.rqs_inverse(yb, params[:5], **kw)
  assert (xb - pts).abs().max() < 1e-12


def test_identity_at_zero_params():
  """Zero raw params => knots equally spaced on [-10, 10], slopes exactly 1,
  spline == identity (flows.py:71-76 relies on this)."""
  p = torch.zeros(16, dtype=torch.float64)
  xp, yp, s = rqs.normalize_knots(p)
  assert torch.allclose(xp, torch.tensor([-10., -6., -2., 2., 6., 10.], dtype=torch.float64), atol=1e-14)
  assert torch.equal(xp, yp)
  assert (s - 1.0).abs().max() < 1e-15
  x = torch.linspace(-12, 12, 97, dtype=torch.float64)
  y, ld, _ = rqs.rqs_forward(x, p)
  assert (y - x).abs().max() < 1e-14 and ld.abs().max() < 1e-14
  xi, ldi, _ = rqs.rqs_inverse(x, p)
  assert (xi - x).abs().max() < 1e-14 and ldi.abs().max() < 1e-14


def test_tails_and_bin_convention():
  gen = torch.Generator().manual_seed(1)
  p = torch.randn(16, generator=gen, dtype=torch.float64)
  xp, yp, s = rqs.normalize_knots(p)
  x = torch.tensor([-15.0, -10.0, 10.0, 12.5], dtype=torch.float64)
  y, ld, idx = rqs.rqs_forward(x, p)
  assert torch.allclose(y[0], (x[0] + 10.0) * s[0] - 10.0)
  assert torch.allclose(y[3], (x[3] - 10.0) * s[-1] + 10.0)
  assert torch.allclose(ld[0], torch.log(s[0])) and torch.allclose(ld[3], torch.log(s[-1]))
  assert idx.tolist() == [0, 0, 0, 0]  # both tails report bin 0
  # an x exactly on an interior knot belongs to the right-hand bin
  _, _, idk = rqs.rqs_forward(xp[2:3].clone(), p)
  assert idk.item() == 2


def test_monotone_and_continuous_derivative():
  gen = torch.Generator().manual_seed(3)
  p = torch.randn(25, generator=gen, dtype=torch.float64) * 1.5
  x = torch.linspace(-11, 11, 4001, dtype=torch.float64, requires_grad=True)
  y, ld, _ = rqs.rqs_forward(x, p)
  (g, ) = torch.autograd.grad(y.sum(), x)
  assert (y[1:] > y[:-1]).all()
  assert (torch.log(g) - ld).abs().max() < 1e-10
