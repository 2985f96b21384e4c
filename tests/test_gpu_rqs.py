"""GPU parity: cnfot_rqs_* (C ABI) vs the oracle's restatement of
distrax.RationalQuadraticSpline, plus the invariants of the reference's own
tests/test_rqs_accuracy.py re-run on the kernels at float32 tolerances."""
import pytest
import torch

from cnf_ot_b200 import ops
from oracle import rqs
from util import rel_err

pytestmark = pytest.mark.gpu
# |a-b| / (|b|+1), float32 kernels vs float64 oracle on float32 inputs: 99.9 % of the rows
# within TOL, all rows within TOL_MAX (narrow bins amplify the rounding of x - x_k)
TOL, TOL_MAX = 2e-5, 2e-4


def close(a, b, tol=TOL, tol_max=TOL_MAX):
  a = a.detach().cpu().double().reshape(-1)
  b = b.detach().cpu().double().reshape(-1)
  e = (a - b).abs() / (b.abs() + 1.0)
  return float(e.quantile(0.999)) < tol and float(e.max()) < tol_max


def _case(K, n, seed, scale=0.5, spread=5.0):
  g = torch.Generator().manual_seed(seed)
  theta = (torch.randn(n, 3 * K + 1, generator=g, dtype=torch.float64) * scale).float()
  v = (torch.randn(n, generator=g, dtype=torch.float64) * spread).float()
  return theta, v


@pytest.mark.parametrize("K", [3, 4, 5, 8, 10, 16])
@pytest.mark.parametrize("inverse", [False, True])
def test_values_bins_and_vjp(K, inverse):
  n = 5000 + 37  # ragged last tile
  theta, v = _case(K, n, K + 100 * inverse)
  g = torch.Generator().manual_seed(9)
  gout = torch.randn(n, generator=g).float()
  gld = torch.randn(n, generator=g).float()
  f = rqs.rqs_inverse if inverse else rqs.rqs_forward
  vv, th = v.double().requires_grad_(True), theta.double().requires_grad_(True)
  o, l, idx = f(vv, th)
  (o * gout.double() + l * gld.double()).sum().backward()
  fn = ops.rqs_inverse if inverse else ops.rqs_forward
  out, ld, bins = fn(v.cuda(), theta.cuda(), K, want_bins=True)
  assert close(out, o)
  assert close(ld, l)
  # bin indices must match exactly except where the input sits on a knot (fp32 vs fp64 tie)
  xp, yp, _ = rqs.normalize_knots(theta.double())
  pos = yp if inverse else xp
  near = ((v.double().unsqueeze(-1) - pos).abs().min(-1).values < 1e-4)
  mism = (bins.cpu().long() != idx) & ~near
  assert int(mism.sum()) == 0
  gin, gth = ops.rqs_vjp(inverse, v.cuda(), theta.cuda(), gout.cuda(), gld.cuda(), K)
  ok = ~near  # adjoints jump across knots
  assert close(gin.cpu()[ok], vv.grad[ok], 1e-4, 2e-3)
  assert close(gth.cpu()[ok], th.grad[ok], 1e-4, 2e-3)


@pytest.mark.parametrize("K", [5, 8])
def test_reference_invariants_on_gpu(K):
  """test_rqs_accuracy.py's four checks (round trips, log-det, boundary points), float32."""
  n = 1 << 20
  theta, v = _case(K, n, 7, scale=0.5, spread=3.0)
  theta, v = theta.cuda(), v.cuda()
  y, ld, _ = ops.rqs_forward(v, theta, K)
  xr, ldi, _ = ops.rqs_inverse(y, theta, K)
  err = (xr - v).abs() / (v.abs() + 1)
  assert float(err.float().quantile(0.999)) < 2e-5 and float(err.max()) < 1e-3, float(err.max())
  assert float((ld + ldi).abs().float().quantile(0.999)) < 5e-5
  assert bool(torch.isfinite(y).all()) and bool(torch.isfinite(ld).all())
  # boundary points of the reference test: range_min+eps, range_max-eps, 0
  pts = torch.tensor([-10 + 1e-4, 10 - 1e-4, 0.0, -10.0, 10.0, -12.0, 15.0], device="cuda")
  th = theta[:pts.numel()].contiguous()
  yb, _, _ = ops.rqs_forward(pts, th, K)
  xb, _, _ = ops.rqs_inverse(yb, th, K)
  assert float((xb - pts).abs().max()) < 1e-4


def test_identity_at_zero_params():
  x = torch.linspace(-12, 12, 1001, device="cuda")
  theta = torch.zeros(1001, 16, device="cuda")
  y, ld, bins = ops.rqs_forward(x, theta, 5, want_bins=True)
  assert float((y - x).abs().max()) < 2e-6 and float(ld.abs().max()) < 2e-6
  xi, ldi, _ = ops.rqs_inverse(x, theta, 5)
  assert float((xi - x).abs().max()) < 2e-6 and float(ldi.abs().max()) < 2e-6


def test_empty_and_single_row():
  e = torch.empty(0, device="cuda")
  y, ld, _ = ops.rqs_forward(e, torch.empty(0, 16, device="cuda"), 5)
  assert y.numel() == 0 and ld.numel() == 0
  y, ld, _ = ops.rqs_forward(torch.tensor([0.3], device="cuda"), torch.zeros(1, 16, device="cuda"), 5)
  assert abs(float(y) - 0.3) < 1e-6


def test_unsupported_bins_fail_loudly():
  from cnf_ot_b200._lib import CnfotError
  with pytest.raises(CnfotError):
    ops.rqs_forward(torch.zeros(4, device="cuda"), torch.zeros(4, 22, device="cuda"), 7)
  with pytest.raises(CnfotError):
    ops.rqs_forward(torch.zeros(4), torch.zeros(4, 16), 5)  # CPU tensors: no CPU path
